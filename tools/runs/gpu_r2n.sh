#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2n_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2n_k4_hints_$i.json 2> gpurun_out/r2n_k4_hints_$i.err
LF_NO_L2_HINTS=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2n_k4_nohints_$i.json 2> gpurun_out/r2n_k4_nohints_$i.err
done
timeout 300 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2n_k5_hints.json 2> gpurun_out/r2n_k5_hints.err
LF_NO_L2_HINTS=1 timeout 300 python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r2n_k5_nohints.json 2> gpurun_out/r2n_k5_nohints.err
timeout 600 ncu --set full --clock-control none -k regex:'narrow_kernel|modulate|mid_kernel|finalize|reduce_splits|rows_' -s 20 -c 12 -f -o /tmp/r2l_k3 \
     python bench.py --workload k3 --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity-check > gpurun_out/r2l_ncu_k3.log 2>&1
python tools/ncu_summary.py /tmp/r2l_k3.ncu-rep gpurun_out/r2l_k3_fp32 k3/fp32 > gpurun_out/r2l_sum_k3.log 2>&1
echo done
