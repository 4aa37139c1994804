#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_step_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2f_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/parity_multigpu.py --workload k4 > gpurun_out/r2f_par_k4.json 2> gpurun_out/r2f_par_k4.err
LF_MID_TRACE=1 LF_DW_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check --no-graph > gpurun_out/r2f_trace_n2.json 2> gpurun_out/r2f_trace_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_k4_n2.json 2> gpurun_out/r2f_k4_n2.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_k4_n1.json 2> gpurun_out/r2f_k4_n1.err
echo done
