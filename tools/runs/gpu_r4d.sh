#!/bin/bash
# why does the calibrated-count pass crawl beside the dfeat GEMM?  L2-only loads (bit 0) / the <= 64-register kernel (bit 1)
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check"
for v in 0 1 2 3; do
  LF_CAL_VARIANT=$v $B --workload k5 > gpurun_out/r4d_k5_v$v.json 2> gpurun_out/r4d_k5_v$v.err
  LF_NO_CAL_OVERLAP=1 LF_CAL_VARIANT=$v $B --workload k5 > gpurun_out/r4d_k5_serial_v$v.json 2> gpurun_out/r4d_k5_serial_v$v.err
done
echo done
