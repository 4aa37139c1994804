#!/bin/bash
mkdir -p gpurun_out
N=8
nvidia-smi topo -m > gpurun_out/r2t_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)" > gpurun_out/r2t_lscpu.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2t_k4_n8.json 2> gpurun_out/r2t_k4_n8.err
echo done
