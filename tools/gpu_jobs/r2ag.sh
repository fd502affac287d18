# round-2 GPU job ag: lanes x chunk size again, now that host copies no longer serialise the lanes
mkdir -p gpurun_out
run() { name=$1; shift
  env WM_BENCH_NO_CPU=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag_$name.json 2> gpurun_out/r2ag_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2ag_{sys.argv[1]}.json"))
    print(sys.argv[1], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run l8_c24 WM_LANES=8 WM_CHUNK_BASES=24000000
run l12_c24 WM_LANES=12 WM_CHUNK_BASES=24000000
run l12_c16 WM_LANES=12 WM_CHUNK_BASES=16000000
run l16_c16 WM_LANES=16 WM_CHUNK_BASES=16000000
