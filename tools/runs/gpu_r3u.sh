#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/prof_round3.py > gpurun_out/r3u_event_times.txt 2> gpurun_out/r3u_plain.err || { echo plain-failed; exit 1; }
REPS=1 timeout 600 ncu --set full --clock-control none -k regex:'lf::' -s 24 -c 24 -f -o /tmp/r3u python tools/prof_round3.py > gpurun_out/r3u_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r3u.ncu-rep gpurun_out/r3u_exact_multi_hidden r3u > gpurun_out/r3u_sum.log 2>&1
echo done
