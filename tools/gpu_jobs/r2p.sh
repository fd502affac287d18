# round-2 GPU job p: backtrack budget (fewer, larger fill launches), then the full-size configuration 5 (3 Gbp reference)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2p_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms ({r['launches']} launches) / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run bt32
run bt64 WM_BT_BUDGET_GB=64
run bt100 WM_BT_BUDGET_GB=100
nvidia-smi --query-gpu=memory.used --format=csv
echo "--- configuration 5 at full size (3 Gbp) ---"
( time WM_BENCH_REF_LEN=3000000000 WM_BT_BUDGET_GB=64 timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench_3gbp.json 2> gpurun_out/r2p_bench_3gbp.err ) 2>&1 | grep real
tail -c 2700 gpurun_out/r2p_bench_3gbp.json; grep -E "bench\]" gpurun_out/r2p_bench_3gbp.err | tail -6
