# round-2 GPU job ac: steps of 2000 reads (a pass of four rounds of the lanes instead of two)
mkdir -p gpurun_out
WM_BENCH_NO_CPU=1 timeout 1200 python bench.py --steps 20 --warmup 5 --reads 2000 > gpurun_out/r2ac.json 2> gpurun_out/r2ac.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2ac.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} ms/step {d['ms_per_step']:.0f}")
PY
tail -3 gpurun_out/r2ac.err
