#!/bin/bash
# round-2 measurement pass: GPU tests, benches of every workload, ncu launch list and one --set full capture of K4
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_k4_bf16.json 2> gpurun_out/r2a_k4_bf16.err
python bench.py --workload k5 --no-cpu-baseline > gpurun_out/r2a_k5_bf16.json 2> gpurun_out/r2a_k5_bf16.err
python bench.py --workload k3 --no-cpu-baseline > gpurun_out/r2a_k3.json 2> gpurun_out/r2a_k3.err
python bench.py --workload k2 --no-cpu-baseline > gpurun_out/r2a_k2.json 2> gpurun_out/r2a_k2.err
python bench.py --workload k4 --precision tf32 --no-cpu-baseline > gpurun_out/r2a_k4_tf32.json 2> gpurun_out/r2a_k4_tf32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_k4_bf16_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tc_|mid_kernel' -s 40 -c 8 -f -o gpurun_out/r2a_k4_bf16 \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2a_ncu_full.log 2>&1
echo done
