#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3e_tests.log
timeout 200 python bench.py --precision fp32 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r3e_k4_fp32.json 2> gpurun_out/r3e_k4_fp32.err
LF_NO_X3=1 timeout 200 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_k4_fp32_fma.json 2> gpurun_out/r3e_k4_fp32_fma.err
timeout 200 python bench.py --precision tf32 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r3e_k4_tf32.json 2> gpurun_out/r3e_k4_tf32.err
timeout 200 python bench.py --workload k5 --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3e_k5_fp32.json 2> gpurun_out/r3e_k5_fp32.err
echo done
