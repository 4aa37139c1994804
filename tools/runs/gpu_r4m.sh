#!/bin/bash
# final tree: ncu launch list of the default bench command (K4 bf16), the pass B200_PROFILING.md prescribes
mkdir -p gpurun_out
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r4m_k4_bf16_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r4m_ncu.log 2>&1
echo done
