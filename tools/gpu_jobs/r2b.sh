# round-2 GPU job b: new tests (boundary, giant chaining tasks, ksw_ll, mid-size), tandem compare, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
WM_TIMING=1 timeout 300 python tools/run_compare.py --len 3400000 --tandem --reads 600 --n50 12000 --repeat > gpurun_out/r2b_tandem.log 2>&1
grep -E "identical|warm" gpurun_out/r2b_tandem.log
WM_TIMING=1 timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2500 gpurun_out/r2b_bench.json
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2b_bench_ncu.json 2> gpurun_out/r2b_bench_ncu.err
wc -l gpurun_out/r2b_launches.csv
