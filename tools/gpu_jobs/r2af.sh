# round-2 GPU job af: the driver's command at N = 2 on the final build
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2af_n2.json 2> gpurun_out/r2af_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2af_n2.json").read().strip().splitlines()[-1])
print(f"N=2 value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB")
PY
tail -3 gpurun_out/r2af_n2.err
