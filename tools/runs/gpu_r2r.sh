#!/bin/bash
mkdir -p gpurun_out
LF_BWD_TRACE=1 LF_FWD_TRACE=1 LF_DW_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2r_trace.json 2> gpurun_out/r2r_trace.err
echo done
