# round-2 GPU job o: device index build + new defaults: full GPU suite, driver-style N=1 run, ncu full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2o_pytest.log 2>&1; tail -4 gpurun_out/r2o_pytest.log
WM_TIMING=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 2600 gpurun_out/r2o_bench.json; grep -E "index ready|workload" gpurun_out/r2o_bench.err
for k in wm_extd2_fill_kernel wm_extd2_fill_coop_kernel wm_chain_fill_tile_kernel wm_anchor_sort_giant_kernel; do
  WM_BENCH_NO_CPU=1 WM_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/r2o_$k -f python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2o_$k.log 2>&1
  ls -la gpurun_out/r2o_$k.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
