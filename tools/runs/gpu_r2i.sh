#!/bin/bash
# N = 2: multi-GPU tests, K4 / K5 / K3 / K2 benches (weak), K4 strong
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2i_tests.log
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; }
run k4_n2
run k5_n2 --workload k5
run k3_n2 --workload k3
run k2_n2 --workload k2
run k4_n2_strong --scaling strong
echo done
