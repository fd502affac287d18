# round-2 first GPU job: tests, bench on the TR workload, tandem compare per chaining formulation, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_env.log; nproc >> gpurun_out/r2a_env.log; free -g >> gpurun_out/r2a_env.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
for mode in 0 1 2; do
  WM_CHAIN_MODE=$mode timeout 300 python tools/run_compare.py --len 3400000 --tandem --reads 600 --n50 12000 --repeat > gpurun_out/r2a_tandem_m$mode.log 2>&1
  echo "mode $mode"; grep -E "identical|warm" gpurun_out/r2a_tandem_m$mode.log
done
for ring in 512 2048; do
  WM_CHAIN_RING=$ring timeout 300 python tools/run_compare.py --len 3400000 --tandem --reads 600 --n50 12000 --repeat > gpurun_out/r2a_tandem_ring$ring.log 2>&1
  echo "ring $ring"; grep -E "identical|warm" gpurun_out/r2a_tandem_ring$ring.log
done
WM_TIMING=1 timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 3500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2a_bench_ncu.json 2> gpurun_out/r2a_bench_ncu.err
wc -l gpurun_out/r2a_launches.csv
