# round-2 GPU job q (8 GPUs): the driver's scaling command at N = 8, our arm only
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_b200_n8.json 2> gpurun_out/r2q_b200_n8.err ) 2>&1 | grep real
tail -c 2600 gpurun_out/r2q_b200_n8.json; grep -E "bench\]" gpurun_out/r2q_b200_n8.err | tail -20
