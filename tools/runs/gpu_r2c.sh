#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2c_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_k4_n1.json 2> gpurun_out/r2c_k4_n1.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload k2 > gpurun_out/r2c_k2_n1.json 2> gpurun_out/r2c_k2_n1.err
echo done
