#!/bin/bash
# K5 bf16: bench line with the dominant kernel chosen among the main-stream kernels, then the ncu launch list of the same command
mkdir -p gpurun_out
python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r4k_k5.json 2> gpurun_out/r4k_k5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r4k_k5_bf16_launches.csv python bench.py --workload k5 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r4k_ncu.log 2>&1
echo done
