# round-2 GPU job n: every array above 2048 anchors on the light walker kernel (4 CTAs per SM) vs the two-class split; coop DP threshold
mkdir -p gpurun_out
timeout 300 python tools/bench_sort.py --n 30000 --arrays 400 --check 2>&1 | tail -2
WM_SORT_GIANT_MIN=100000000 timeout 300 python tools/bench_sort.py --n 30000 --arrays 400 --check 2>&1 | tail -2
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench_$name.json 2> gpurun_out/r2n_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2n_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run allmedium WM_SORT_GIANT_MIN=100000000
run coop600k WM_DP_COOP_MIN_CELLS=600000
run coop300k_allmedium WM_DP_COOP_MIN_CELLS=300000 WM_SORT_GIANT_MIN=100000000
