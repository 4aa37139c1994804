#!/bin/bash
run() { echo "== $*"; env $1 python bench.py --workload k4 --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print({n:k[n]['avg_us'] for n in ('tc_heads_forward','tc_dfeat','tc_dweight','rows_backward_qmf') if n in k}, 'ms/step', round(d['ms_per_step'],3))"; }
run LF_FWD_DBG=15
run LF_FWD_DBG=3
run LF_FWD_DBG=0
