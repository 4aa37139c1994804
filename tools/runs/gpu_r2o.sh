#!/bin/bash
mkdir -p gpurun_out
for m in 15 0 9 11 13 8 1; do
LF_L2_HINTS=$m timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2o_k4_m$m.json 2> gpurun_out/r2o_k4_m$m.err
done
echo done
