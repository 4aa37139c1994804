#!/bin/bash
# two GPUs after the fix (an unsharded engine inside a process group no longer all-gathers its payload): sharded parity of every
# workload incl. the DDP layout, repeated K5 trials, K5 bench line with parity_check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r4i_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
: > gpurun_out/r4i_trials.log
for t in 1 2 3; do
  timeout 100 $TR tools/parity_multigpu.py --workload k5 --batch 2048 2>&1 | grep -E "parity_check" | cut -c1-1500 >> gpurun_out/r4i_trials.log
done
timeout 300 $TR bench.py --gpus 2 --workload k5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r4i_k5_n2.json 2> gpurun_out/r4i_k5_n2.err
echo done
