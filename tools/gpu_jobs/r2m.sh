# round-2 GPU job m: ncu --set full captures of the dominant kernels on the bench workload (one GPU)
mkdir -p gpurun_out
for k in wm_extd2_fill_kernel wm_extd2_fill_coop_kernel wm_chain_fill_tile_kernel wm_anchor_sort_giant_kernel; do
  WM_BENCH_NO_CPU=1 WM_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/r2m_$k -f python bench.py --steps 1 --warmup 1 --reads 600 > gpurun_out/r2m_$k.log 2>&1
  ls -la gpurun_out/r2m_$k.ncu-rep 2>/dev/null
done
