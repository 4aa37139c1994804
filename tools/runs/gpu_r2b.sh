#!/bin/bash
# 2-GPU pass: tests, sharded parity, K4 bench at N=1 and N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2b_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_k4_n1.json 2> gpurun_out/r2b_k4_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_k4_n2.json 2> gpurun_out/r2b_k4_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --workload k5 > gpurun_out/r2b_k5_n2.json 2> gpurun_out/r2b_k5_n2.err
echo done
