# round-2 GPU job l (2 GPUs): the driver's multi-GPU commands, both arms
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_ref_n2.json 2> gpurun_out/r2l_ref_n2.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2l_ref_n2.json
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_b200_n2.json 2> gpurun_out/r2l_b200_n2.err ) 2>&1 | grep real
tail -c 2500 gpurun_out/r2l_b200_n2.json; tail -5 gpurun_out/r2l_b200_n2.err
