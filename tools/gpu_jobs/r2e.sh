# round-2 GPU job e: barrier-free tile chaining kernel, extz2, sort threshold variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -k "chain or extz2 or single_gap or sort" > gpurun_out/r2e_pytest_a.log 2>&1; tail -4 gpurun_out/r2e_pytest_a.log
run() { # name, env...
  name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2e_bench_{sys.argv[1]}.json"))
r, o = d['roofline'], d['roofline_other']
print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", {k: round(v, 1) for k, v in d["breakdown_s"].items()}, f"{r['kernel'][3:14]} {r['kernel_ms']:.0f} ms / {o['kernel'][3:14]} {o['kernel_ms']:.0f} ms")
PY
}
run base WM_CHUNK_BASES=32000000
run sortmin2048 WM_CHUNK_BASES=32000000 WM_SORT_GIANT_MIN=2048
run sortmin4096 WM_CHUNK_BASES=32000000 WM_SORT_GIANT_MIN=4096
run tilemin512 WM_CHUNK_BASES=32000000 WM_CHAIN_TILE_MIN=512
run tilemin8192 WM_CHUNK_BASES=32000000 WM_CHAIN_TILE_MIN=8192
WM_BENCH_NO_CPU=1 WM_DP_STATS=1 WM_CHUNK_BASES=32000000 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_stats.json 2> gpurun_out/r2e_stats.err
grep "sort-stats" gpurun_out/r2e_stats.err | sort -t= -k6 -n | tail -5
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
WM_BENCH_NO_CPU=1 WM_CHUNK_BASES=32000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_bench_ncu.json 2> gpurun_out/r2e_bench_ncu.err
python tools/ncu_launch_summary.py gpurun_out/r2e_launches.csv 10
