#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
: > gpurun_out/r4h_trials.log
for t in 1 2 3 4 5; do
  for v in default nooverlap; do
    case $v in default) E="LF_PARITY_VERBOSE=1";; nooverlap) E="LF_PARITY_VERBOSE=1 LF_NO_CAL_OVERLAP=1";; esac
    echo "== trial $t $v" >> gpurun_out/r4h_trials.log
    env $E timeout 100 $TR tools/parity_multigpu.py --workload k5 --batch 2048 2>&1 | grep -E "^\[rank|parity_check" | cut -c1-420 >> gpurun_out/r4h_trials.log
  done
done
echo done
