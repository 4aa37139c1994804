#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2v_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_k4_n8.json 2> gpurun_out/r2v_k4_n8.err
echo done
