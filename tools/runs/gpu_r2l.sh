#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2l_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_k4.json 2> gpurun_out/r2l_k4.err
for wl in k5 k3 k2; do
  prec=fp32; [ $wl = k5 ] && prec=bf16
  timeout 600 ncu --set full --clock-control none -k regex:'lf::' -s 16 -c 10 -f -o /tmp/r2l_${wl} \
     python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2l_ncu_${wl}.log 2>&1
  TC_ORDER=tc_logits,tc_dfeat,tc_dweight python tools/ncu_summary.py /tmp/r2l_${wl}.ncu-rep gpurun_out/r2l_${wl}_${prec} ${wl}/${prec} > gpurun_out/r2l_sum_${wl}.log 2>&1
done
ls -la /tmp/*.ncu-rep > gpurun_out/r2l_reps.txt
echo done
