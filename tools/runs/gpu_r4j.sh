#!/bin/bash
# end-of-round state on one GPU: whole GPU suite, smoke, the driver's default bench line (K4 bf16) and the K5 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r4j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4j_smoke.log 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/r4j_k4.json 2> gpurun_out/r4j_k4.err
python bench.py --workload k5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r4j_k5.json 2> gpurun_out/r4j_k5.err
echo done
