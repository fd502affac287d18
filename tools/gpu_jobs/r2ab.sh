# round-2 GPU job ab: launch list of the final build and ncu --set full captures of the packed-pool kernels
mkdir -p gpurun_out
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_launches.csv python bench.py --steps 4 --warmup 2 > gpurun_out/r2ab_bench_ncu.json 2> gpurun_out/r2ab_bench_ncu.err
python tools/ncu_launch_summary.py gpurun_out/r2ab_launches.csv 24
for k in wm_sketch_order_kernel wm_gather2_kernel wm_pack_gather_kernel; do
  WM_BENCH_NO_CPU=1 WM_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/r2ab_$k -f python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2ab_$k.log 2>&1
  ls -la gpurun_out/r2ab_$k.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
gzip -f gpurun_out/r2ab_launches.csv
