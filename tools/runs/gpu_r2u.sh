#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2u_k4_bf16_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2u_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'tc_|mid_kernel' -s 24 -c 8 -f -o /tmp/r2u_k4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2u_ncu_full.log 2>&1
TC_ORDER=tc_dweight python tools/ncu_summary.py /tmp/r2u_k4.ncu-rep gpurun_out/r2u_k4_bf16 k4/bf16 > gpurun_out/r2u_sum.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none -k regex:'tc_|mid_kernel' -s 24 -c 8 -f -o /tmp/r2u_k4_warm python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2u_ncu_full_warm.log 2>&1
TC_ORDER=tc_dweight python tools/ncu_summary.py /tmp/r2u_k4_warm.ncu-rep gpurun_out/r2u_k4_bf16_warmcache k4/bf16-warm-l2 > gpurun_out/r2u_sum_warm.log 2>&1
echo done
