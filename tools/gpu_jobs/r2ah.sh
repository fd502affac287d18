# round-2 GPU job ah: where the rest of the lane time goes (upload, host encoding, idle tail)
mkdir -p gpurun_out
WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ah.json 2> gpurun_out/r2ah.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2ah.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
awk '/timers over/{f=1} f' gpurun_out/r2ah.err | grep -E "batch\.|lane\.|wave\.seed_chain|round.run_dp|stage"
