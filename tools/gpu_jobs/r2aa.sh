# round-2 GPU job aa: CIGAR offsets scanned and compacted on the device behind the traceback
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_boundary.py -m gpu -x -q --timeout 900 > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
for i in 1 2; do
WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aa_$i.json 2> gpurun_out/r2aa_$i.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2aa_$i.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
done
awk '/timers over/{f=1} f' gpurun_out/r2aa_2.err | grep -E "d2h|seed.chain|fill_bt|lookup_sort"
