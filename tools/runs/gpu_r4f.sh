#!/bin/bash
# two GPUs: sharded parity of every workload (mid_light_kernel with the peer exchange, two-stream mean fusion with the fused
# all-reduce), K5 / K4 bench lines with parity_check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r4f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 2 --workload k5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r4f_k5_n2.json 2> gpurun_out/r4f_k5_n2.err
LF_NO_CAL_OVERLAP=1 timeout 300 $TR bench.py --gpus 2 --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r4f_k5_n2_serial.json 2> gpurun_out/r4f_k5_n2_serial.err
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r4f_k4_n2.json 2> gpurun_out/r4f_k4_n2.err
echo done
