#!/bin/bash
# round-2 late: mean fusion with lf_step_mid + calibrated counts beside the dfeat GEMM -- tests, K5 / K4-shaped A/B, sanitizer
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_overlap_gpu.py "tests/test_fullsize_gpu.py::test_k5_full_size_matches_fp64_oracle" -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r4a_tests.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check"
$B --workload k5 > gpurun_out/r4a_k5_overlap.json 2> gpurun_out/r4a_k5_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 > gpurun_out/r4a_k5_serial.json 2> gpurun_out/r4a_k5_serial.err
$B --workload k5 --classes 101 --dim 768 --batch 32768 > gpurun_out/r4a_mean101_overlap.json 2> gpurun_out/r4a_mean101_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 --classes 101 --dim 768 --batch 32768 > gpurun_out/r4a_mean101_serial.json 2> gpurun_out/r4a_mean101_serial.err
$B --workload k5 --precision tf32 > gpurun_out/r4a_k5_tf32_overlap.json 2> gpurun_out/r4a_k5_tf32_overlap.err
LF_NO_CAL_OVERLAP=1 $B --workload k5 --precision tf32 > gpurun_out/r4a_k5_tf32_serial.json 2> gpurun_out/r4a_k5_tf32_serial.err
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_step.py > gpurun_out/r4a_sanitize.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r4a_sanitize.log
echo done
