# round-2 GPU job t: footprint of 64 Mbase chunks; DP fill stream priority
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench_$name.json 2> gpurun_out/r2t_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2t_bench_{sys.argv[1]}.json"))
    r, o = d['roofline'], d['roofline_other']
    print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms ({r['launches']} launches) / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run c32
run c64 WM_CHUNK_BASES=64000000
run c32_prio WM_FILL_PRIO=high
run c64_prio WM_CHUNK_BASES=64000000 WM_FILL_PRIO=high
