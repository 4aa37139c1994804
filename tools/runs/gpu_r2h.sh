#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_step_gpu.py tests/test_parity_gpu.py tests/test_boundary_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2h_tests.log
LF_MID_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2h_trace.json 2> gpurun_out/r2h_trace.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2h_k4_a.json 2> gpurun_out/r2h_k4_a.err
LF_MID_NOCOOP=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2h_k4_b.json 2> gpurun_out/r2h_k4_b.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_k4_c.json 2> gpurun_out/r2h_k4_c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2h_ncu.log 2>&1
echo done
