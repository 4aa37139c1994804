#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err; }
run k4_n8 --steps 50
LF_MID_TRACE=1 LF_DW_TRACE=1 run k4_n8_trace --steps 20 --no-graph --no-parity-check
echo done
