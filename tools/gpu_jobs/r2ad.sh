# round-2 GPU job ad: two-bucket radix passes in closed form (walker kernels): parity, sort timing with / without, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "sort" --timeout 600 2>&1 | tail -3
for tm in 0 512; do
  echo "== WM_SORT_TWO_MIN=$tm"
  WM_SORT_TWO_MIN=$tm timeout 300 python tools/bench_sort.py --n 30000 --arrays 100 --strand-frac 0.02 --check 2>&1 | tail -2
  WM_SORT_TWO_MIN=$tm WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 150000 --arrays 16 --strand-frac 0.02 --check 2>&1 | tail -4
done
timeout 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q --timeout 900 2>&1 | tail -2
for tm in 512 0; do
WM_SORT_TWO_MIN=$tm WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ad_$tm.json 2> gpurun_out/r2ad_$tm.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2ad_$tm.json"))
print(f"two_min $tm: value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
awk '/timers over/{f=1} f' gpurun_out/r2ad_$tm.err | grep -E "lookup_sort|concat_sort3"
done
