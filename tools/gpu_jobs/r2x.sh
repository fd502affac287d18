# round-2 GPU job x: the lanes' phases on one time axis (WM_TIMELINE), driver-style run without the CPU arm
mkdir -p gpurun_out; rm -f gpurun_out/r2x_timeline.tsv
WM_TIMELINE=gpurun_out/r2x_timeline.tsv WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x.json 2> gpurun_out/r2x.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2x.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
gzip -f gpurun_out/r2x_timeline.tsv; ls -la gpurun_out/r2x_timeline.tsv.gz
