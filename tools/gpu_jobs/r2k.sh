# round-2 GPU job k: how many SMs the CTA-per-task kernels may hold (they block the DP fill CTAs), launch list
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_$name.json 2> gpurun_out/r2k_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2k_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run base
run tile74 WM_CHAIN_TILE_CTAS=74
run tile37 WM_CHAIN_TILE_CTAS=37
run med296 WM_SORT_MEDIUM_CTAS=296
run med296_tile74_g74 WM_SORT_MEDIUM_CTAS=296 WM_CHAIN_TILE_CTAS=74 WM_SORT_GIANT_CTAS=74
WM_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 4 --warmup 2 > gpurun_out/r2k_bench_ncu.json 2> gpurun_out/r2k_bench_ncu.err
python tools/ncu_launch_summary.py gpurun_out/r2k_launches.csv 14
