#!/bin/bash
# deep-prefetch calibrated-count kernel (R rows of both heads in registers before any reduction) beside the dfeat GEMM
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check"
for v in 4 8; do
  LF_CAL_VARIANT=$v $B --workload k5 > gpurun_out/r4e_k5_v$v.json 2> gpurun_out/r4e_k5_v$v.err
  LF_NO_CAL_OVERLAP=1 LF_CAL_VARIANT=$v $B --workload k5 > gpurun_out/r4e_k5_serial_v$v.json 2> gpurun_out/r4e_k5_serial_v$v.err
done
LF_CAL_VARIANT=4 $B --workload k5 --precision tf32 > gpurun_out/r4e_k5_tf32_v4.json 2> gpurun_out/r4e_k5_tf32_v4.err
LF_CAL_VARIANT=4 timeout 300 python -m pytest tests/test_overlap_gpu.py "tests/test_fullsize_gpu.py::test_k5_full_size_matches_fp64_oracle" -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r4e_tests.log
echo done
