# round-2 GPU job v: packed 2-bit + N read pool (sketch from 64-bit windows, window minimum by doubling, DP gather from the
# pool): the full GPU suite, then the driver-style bench with its parity check
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2v_pytest.log 2>&1; tail -5 gpurun_out/r2v_pytest.log
WM_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2v_bench.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB parity {d.get('parity_checked')} cpu {d['cpu_baseline']['value']}")
print(d['breakdown_s'], d['e2e'])
PY
grep -E "seed.sketch|seed.a_mask|dp.gpu_fill_bt|seed.lookup_sort" gpurun_out/r2v_bench.err | head
