# round-2 GPU job w: how much do the phases of a chunk stretch when lanes run side by side?  Same chunks (32 Mbase), 1 / 4 / 8 lanes,
# host timers summed over the chunks of the timed pass.
mkdir -p gpurun_out
for L in 1 4 8; do
  WM_LANES=$L WM_TIMING=1 WM_BENCH_NO_CPU=1 timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2w_l$L.json 2> gpurun_out/r2w_l$L.err
  echo "== lanes $L"; python - <<PY
import json
d = json.load(open("gpurun_out/r2w_l$L.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s")
PY
  awk '/timers over/{f=1} f' gpurun_out/r2w_l$L.err | grep -E "seed\.|dp\.|wave\.seed|round\.run" | head -20
done
