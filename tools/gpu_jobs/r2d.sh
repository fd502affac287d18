# round-2 GPU job d: giant sort + parallel backtrack: tests, tandem compare, bench (+ two chunk settings), ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "sort or chain" > gpurun_out/r2d_pytest_a.log 2>&1; tail -4 gpurun_out/r2d_pytest_a.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1; tail -6 gpurun_out/r2d_pytest.log
WM_TIMING=1 timeout 300 python tools/run_compare.py --len 3400000 --tandem --reads 600 --n50 12000 --repeat > gpurun_out/r2d_tandem.log 2>&1
grep -E "identical|warm" gpurun_out/r2d_tandem.log
WM_TIMING=1 timeout 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 2600 gpurun_out/r2d_bench.json
for cfg in "8 32000000" "4 32000000"; do
  set -- $cfg
  WM_BENCH_NO_CPU=1 WM_LANES=$1 WM_CHUNK_BASES=$2 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2d_bench_$1_$2.json 2> gpurun_out/r2d_bench_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2d_bench_{sys.argv[1]}_{sys.argv[2]}.json"))
print(sys.argv[1], sys.argv[2], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d["breakdown_s"], f"fill {d['roofline']['kernel_ms']:.0f}/{d['roofline_other']['kernel_ms']:.0f} ms")
PY
done
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2d_bench_ncu.json 2> gpurun_out/r2d_bench_ncu.err
python tools/ncu_launch_summary.py gpurun_out/r2d_launches.csv 12
