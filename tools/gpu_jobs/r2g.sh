# round-2 GPU job g: where the giant sort spends its clocks; seed_chain timer breakdown; 8-host-thread ranks
mkdir -p gpurun_out
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 1 2>&1 | tail -3
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 200 2>&1 | tail -3
WM_TIMING=1 WM_BENCH_NO_CPU=1 WM_CHUNK_BASES=32000000 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
grep -A30 "timers over" gpurun_out/r2g_bench.err | grep -E "seed\.|wave|round.run|dp\." 
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2g_bench_$name.json 2> gpurun_out/r2g_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2g_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d['config']['host_threads'], d['config']['lanes'], {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run thr8_lanes2 WM_CHUNK_BASES=32000000 WM_THREADS_PER_RANK=8
run thr8_lanes8 WM_CHUNK_BASES=32000000 WM_THREADS_PER_RANK=8 WM_LANES=8
run thr16_lanes8 WM_CHUNK_BASES=32000000 WM_THREADS_PER_RANK=16 WM_LANES=8
run thr64_lanes12 WM_CHUNK_BASES=24000000 WM_LANES=12
