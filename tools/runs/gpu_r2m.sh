#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_boundary_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2m_tests.log
for wl in k5 k3 k2; do
  prec=fp32; [ $wl = k5 ] && prec=bf16
  timeout 600 ncu --set full --clock-control none -k regex:'_kernel' -s 16 -c 12 -f -o /tmp/r2l_${wl} \
     python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2l_ncu_${wl}.log 2>&1
  TC_ORDER=tc_logits,tc_dfeat,tc_dweight python tools/ncu_summary.py /tmp/r2l_${wl}.ncu-rep gpurun_out/r2l_${wl}_${prec} ${wl}/${prec} > gpurun_out/r2l_sum_${wl}.log 2>&1
done
echo done
