# round-2 GPU job h: throttled polling (giant sort feeders, chaining tile kernel)
mkdir -p gpurun_out
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 1 --check 2>&1 | tail -3
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 200 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "sort or chain" > gpurun_out/r2h_pytest_a.log 2>&1; tail -3 gpurun_out/r2h_pytest_a.log
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2h_bench_$name.json 2> gpurun_out/r2h_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2h_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d['config']['host_threads'], d['config']['lanes'], {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run base
run sortmin2048 WM_SORT_GIANT_MIN=2048
run sortmin6144 WM_SORT_GIANT_MIN=6144
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2h_bench_ncu.json 2> gpurun_out/r2h_bench_ncu.err
python tools/ncu_launch_summary.py gpurun_out/r2h_launches.csv 12
