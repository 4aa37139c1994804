#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_tc_gemm_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3k_tests.log
timeout 200 python bench.py --precision fp32 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r3k_k4_fp32.json 2> gpurun_out/r3k_k4_fp32.err
timeout 200 python bench.py --workload k5 --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3k_k5_fp32.json 2> gpurun_out/r3k_k5_fp32.err
echo done
