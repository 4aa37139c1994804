#!/bin/bash
# final scaling points: N = 4 (never run before) with the parity check, K4 and K5
mkdir -p gpurun_out
N=$1
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2s_${name}_n$N.json 2> gpurun_out/r2s_${name}_n$N.err; }
run k4
run k5 --workload k5
run k4_strong --scaling strong
echo done
