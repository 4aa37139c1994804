#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3s_tests.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r3s_k4_bf16.json 2> gpurun_out/r3s_k4_bf16.err
timeout 200 python bench.py --precision tf32 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r3s_k4_tf32.json 2> gpurun_out/r3s_k4_tf32.err
timeout 200 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r3s_k5_bf16.json 2> gpurun_out/r3s_k5_bf16.err
echo done
