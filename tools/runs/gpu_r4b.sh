#!/bin/bash
# mid_light_kernel (no cooperative launch, 128 B of shared memory) so that lf_step_mid runs BESIDE the dfeat GEMM: full GPU suite, K5 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r4b_tests.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check"
$B --workload k5 > gpurun_out/r4b_k5_overlap.json 2> gpurun_out/r4b_k5_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 > gpurun_out/r4b_k5_serial.json 2> gpurun_out/r4b_k5_serial.err
$B --workload k5 --precision tf32 > gpurun_out/r4b_k5_tf32_overlap.json 2> gpurun_out/r4b_k5_tf32_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 --precision tf32 > gpurun_out/r4b_k5_tf32_serial.json 2> gpurun_out/r4b_k5_tf32_serial.err
$B --workload k5 --precision fp32 > gpurun_out/r4b_k5_fp32_overlap.json 2> gpurun_out/r4b_k5_fp32_overlap.err
$B --workload k3 > gpurun_out/r4b_k3.json 2> gpurun_out/r4b_k3.err
$B --workload k1 > gpurun_out/r4b_k1.json 2> gpurun_out/r4b_k1.err
echo done
