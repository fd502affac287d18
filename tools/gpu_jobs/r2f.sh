# round-2 GPU job f: sort microbenchmark (giant kernel vs single warp), coop DP fill tests + bench
mkdir -p gpurun_out
for g in 1 0; do
  WM_SORT_GIANT=$g timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_sort_g$g.csv python tools/bench_sort.py --n 30000 --arrays 200 --check > gpurun_out/r2f_sort_g$g.log 2>&1
  echo "giant=$g"; tail -2 gpurun_out/r2f_sort_g$g.log; python tools/ncu_launch_summary.py gpurun_out/r2f_sort_g$g.csv 4
done
WM_SORT_GIANT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_sort_one.csv python tools/bench_sort.py --n 30000 --arrays 1 > gpurun_out/r2f_sort_one.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r2f_sort_one.csv 4
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -k "extd2 or e2e" > gpurun_out/r2f_pytest_a.log 2>&1; tail -4 gpurun_out/r2f_pytest_a.log
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2f_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run coop WM_CHUNK_BASES=32000000
run nocoop WM_CHUNK_BASES=32000000 WM_DP_COOP=0
