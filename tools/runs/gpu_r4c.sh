#!/bin/bash
# calibrated-count pass at <= 64 registers (two CTAs per SM beside the dfeat GEMM): tests + K5 A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_overlap_gpu.py tests/test_fullsize_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r4c_tests.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check"
$B --workload k5 > gpurun_out/r4c_k5_overlap.json 2> gpurun_out/r4c_k5_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 > gpurun_out/r4c_k5_serial.json 2> gpurun_out/r4c_k5_serial.err
$B --workload k5 --precision tf32 > gpurun_out/r4c_k5_tf32_overlap.json 2> gpurun_out/r4c_k5_tf32_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 --precision tf32 > gpurun_out/r4c_k5_tf32_serial.json 2> gpurun_out/r4c_k5_tf32_serial.err
echo done
