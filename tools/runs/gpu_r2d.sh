#!/bin/bash
mkdir -p gpurun_out
LF_MID_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2d_trace.json 2> gpurun_out/r2d_trace.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/parity_multigpu.py --workload k4 > gpurun_out/r2d_par_k4.json 2> gpurun_out/r2d_par_k4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/parity_multigpu.py --workload k4 --batch 16384 > gpurun_out/r2d_par_k4b.json 2> gpurun_out/r2d_par_k4b.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2d_k4_n2.json 2> gpurun_out/r2d_k4_n2.err
echo done
