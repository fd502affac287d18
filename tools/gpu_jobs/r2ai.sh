# round-2 GPU job ai: result assembly of a DP round in parallel; golden tests and the driver's command on the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q --timeout 600 2>&1 | tail -2
WM_TIMING=1 timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ai_n1.json 2> gpurun_out/r2ai_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2ai_n1.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s parity {d.get('parity_checked')} cpu {d['cpu_baseline']['value']/1e6:.1f}")
PY
awk '/timers over/{f=1} f' gpurun_out/r2ai_n1.err | grep -E "dp\.|round.run_dp"
