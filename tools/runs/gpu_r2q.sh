#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_step.py > gpurun_out/r2q_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2q_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_step.py > gpurun_out/r2q_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2q_racecheck.log
timeout 300 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2q_k5.json 2> gpurun_out/r2q_k5.err
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2q_k4.json 2> gpurun_out/r2q_k4.err
echo done
