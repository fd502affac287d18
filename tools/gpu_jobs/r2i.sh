# round-2 GPU job i: walker without the per-step S2UR, medium arrays on the light walker kernel, sort #3 in one launch set;
# then the driver's own commands at N = 1 (both arms)
mkdir -p gpurun_out
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 1 --check 2>&1 | tail -3
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 8000 --arrays 1500 --check 2>&1 | tail -2
WM_SORT_MEDIUM=0 timeout 300 python tools/bench_sort.py --n 8000 --arrays 1500 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "sort or chain or e2e" > gpurun_out/r2i_pytest_a.log 2>&1; tail -3 gpurun_out/r2i_pytest_a.log
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2i_bench_$name.json 2> gpurun_out/r2i_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2i_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d['config']['host_threads'], d['config']['lanes'], {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run base
run nomedium WM_SORT_MEDIUM=0
echo "--- driver commands, N=1 ---"
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_driver_ref.json 2> gpurun_out/r2i_driver_ref.err ) 2>&1 | grep real
tail -c 1200 gpurun_out/r2i_driver_ref.json
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_driver_b200.json 2> gpurun_out/r2i_driver_b200.err ) 2>&1 | grep real
tail -c 3000 gpurun_out/r2i_driver_b200.json
