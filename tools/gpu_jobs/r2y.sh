# round-2 GPU job y: result pools in page-locked host memory (D2H of chains and CIGARs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_boundary.py -m gpu -x -q --timeout 900 > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y.json 2> gpurun_out/r2y.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2y.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
awk '/timers over/{f=1} f' gpurun_out/r2y.err | grep -E "d2h|seed.chain|fill_bt|lookup_sort"
