#!/bin/bash
# N = 8: K4 (weak, with the parity check) and K5
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; }
run k4_n8
run k5_n8 --workload k5 --no-parity-check
echo done
