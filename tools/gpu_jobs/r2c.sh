# round-2 GPU job c: full gpu test suite + lanes x chunk sweep on the bench workload
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; tail -8 gpurun_out/r2c_pytest.log
for cfg in "8 16000000" "8 4000000" "16 4000000" "32 4000000" "16 8000000" "32 2000000"; do
  set -- $cfg
  WM_BENCH_NO_CPU=1 WM_LANES=$1 WM_CHUNK_BASES=$2 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2c_bench_$1_$2.json 2> gpurun_out/r2c_bench_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c_bench_{sys.argv[1]}_{sys.argv[2]}.json"))
print(sys.argv[1], sys.argv[2], f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s", d["breakdown_s"], f"fill {d['roofline']['kernel_ms']:.0f}/{d['roofline_other']['kernel_ms']:.0f} ms")
PY
done
