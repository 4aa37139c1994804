#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2p_tests.log
timeout 300 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_k5_msub.json 2> gpurun_out/r2p_k5_msub.err
LF_NO_MSUB=1 timeout 300 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2p_k5_nomsub.json 2> gpurun_out/r2p_k5_nomsub.err
timeout 300 python bench.py --workload k5 --precision tf32 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_k5_tf32.json 2> gpurun_out/r2p_k5_tf32.err
echo done
