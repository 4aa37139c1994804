"""Sharded (multi-GPU) parity, run by pytest when the box has at least two GPUs: N ranks under torchrun must reproduce one
GPU running the concatenated batch (replicated EMA / History / heads bit-identical across ranks, gradients to rounding)
and the CPU oracle on a small problem; and the DDP layout -- unsharded engines inside the process group, a different batch on
every rank -- must behave like lone GPUs (each against the oracle on its own data).  The checker is tools/parity_multigpu.py -- the same code bench.py runs before it
times anything (``parity_check`` in its JSON line), so the driver's scaling runs carry the same evidence."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, *args, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "parity_multigpu.py"), *args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith('{"parity_check"')]
    assert lines, res.stdout[-2000:] + res.stderr[-2000:]
    return res.returncode, json.loads(lines[-1])["parity_check"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("workload,batch", [("k4", 4096), ("k2", 64), ("k3", 1024), ("k5", 2048)])
def test_two_ranks_match_one_gpu_and_the_oracle(workload, batch):
    rc, res = _torchrun(2, "--workload", workload, "--batch", str(batch))
    assert rc == 0 and res["ok"], res
    assert res["vs_single_gpu_on_concatenated_batch"]["ranks_bit_identical"]
    tol = res["oracle_tolerance"]
    assert all(v <= tol for v in res["unsharded_engines_in_the_group_vs_cpu_oracle"].values()), res


def test_parity_checker_on_one_gpu():
    """world = 1: the checker degenerates to engine == engine and engine == oracle; keeps the code path alive on 1-GPU boxes."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_multigpu.py"), "--workload", "k4", "--batch", "2048"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith('{"parity_check"')]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    assert json.loads(lines[-1])["parity_check"]["ok"]
