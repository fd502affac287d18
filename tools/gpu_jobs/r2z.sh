# round-2 GPU job z: splice-aware extension kernel (ksw_exts2) parity; full GPU suite; bench with page-locked result pools (2x growth)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2z_pytest.log 2>&1; tail -5 gpurun_out/r2z_pytest.log
WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z.json 2> gpurun_out/r2z.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2z.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']}")
PY
awk '/timers over/{f=1} f' gpurun_out/r2z.err | grep -E "d2h|seed.chain|fill_bt|lookup_sort"
