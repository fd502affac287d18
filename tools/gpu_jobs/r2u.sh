# round-2 GPU job u: bulk-copy staging in the sort kernels (tests), is the first resident pass slower than a repeat?
mkdir -p gpurun_out
timeout 300 python tools/bench_sort.py --n 1500 --arrays 3000 --check 2>&1 | tail -2
timeout 300 python tools/bench_sort.py --n 30000 --arrays 100 --check 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
WM_BENCH_NO_CPU=1 WM_BENCH_VALUE_TWICE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
grep "resident pass again" gpurun_out/r2u_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2u_bench.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB")
PY
WM_BENCH_NO_CPU=1 timeout 600 python bench.py > gpurun_out/r2u_bench_default.json 2> gpurun_out/r2u_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2u_bench_default.json"))
print(f"defaults (steps {d['steps']}): value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
