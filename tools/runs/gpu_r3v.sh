#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_heads_gpu.py tests/test_hidden_gpu.py tests/test_tc_gemm_gpu.py tests/test_round2_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3v_tests.log
timeout 200 python tools/prof_round3.py > gpurun_out/r3v_event_times.txt 2> gpurun_out/r3v_plain.err || { echo plain-failed; exit 1; }
REPS=1 timeout 600 ncu --set full --clock-control none -k regex:'tc_gemm_kernel|multi_heads_kernel|multi_finalize_kernel|hidden_dpre_kernel|hidden_db_kernel' -s 13 -c 13 -f -o /tmp/r3v python tools/prof_round3.py > gpurun_out/r3v_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r3v.ncu-rep gpurun_out/r3v_exact_multi_hidden r3v > gpurun_out/r3v_sum.log 2>&1
echo done
