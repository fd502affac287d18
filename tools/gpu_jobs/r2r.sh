# round-2 GPU job r: guided chunk sizes; full GPU suite on the final kernels; the driver's N = 1 command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
WM_SORT_DEBUG=1 timeout 300 python tools/bench_sort.py --n 30000 --arrays 1 --check 2>&1 | tail -3
for i in 1 2; do
WM_TIMING=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2r_bench_$i.json 2> gpurun_out/r2r_bench_$i.err
python - $i <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2r_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(f"run {sys.argv[1]}: value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} cpu {d['cpu_baseline']['value']/1e6:.1f} Mbase/s parity {d['parity_checked']}", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms ({r['launches']}) / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
done
