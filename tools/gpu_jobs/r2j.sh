# round-2 GPU job j: bulk-copy (TMA) staging in the fill kernel, walker loop without LDC; submission size / chunk sweep at the driver's step counts
mkdir -p gpurun_out
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 1 --check 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_$name.json 2> gpurun_out/r2j_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2j_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d['config']['host_threads'], d['config']['lanes'], {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run g32_c32
run g32_c16 WM_CHUNK_BASES=16000000
run g32_c24 WM_CHUNK_BASES=24000000
run g8_c32 WM_BENCH_GROUP=8
run g32_c16_l12 WM_CHUNK_BASES=16000000 WM_LANES=12
