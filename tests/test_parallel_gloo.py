"""world_size-2 gloo test (CPU) of the batch-sharded protocol in multimodal_clinical_b200/parallel.py.

Each rank evaluates the ORACLE on its shard with global-batch denominators and exchanges exactly what the
CUDA engine exchanges (all-reduce of the packed statistics, all-gather of idx/conf, one all-reduce of the
flat gradient buffer with the calibrated counts in its tail).  The combined result must equal the
single-process oracle on the concatenated batch — the definition of multi-GPU semantics.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, D, C, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_clinical_b200 import parallel
        from oracle import late_fusion as O
        Bg = B * world
        full = O.make_inputs(Bg, D, C, seed=3, n_data=N)
        lo, cnt = parallel.shard_range(rank, B)
        sl = slice(lo, lo + cnt)
        W = [full["W1"], full["W2"]]; b = [full["b1"], full["b2"]]
        f = [full["f1"][sl].clone().requires_grad_(True), full["f2"][sl].clone().requires_grad_(True)]
        Wl = [w.clone().requires_grad_(True) for w in W]
        bl = [x.clone().requires_grad_(True) for x in b]
        y, idx = full["y"][sl], full["idx"][sl]

        # ---- forward half on the shard; packed stats like LF_STAT_*
        z = O.heads_forward(f, Wl, bl)
        zs = torch.stack(z)
        zdf, conf = O.qmf_df(zs)
        stats = torch.zeros(16 + 2 * C, dtype=torch.float64)
        stats[0] = torch.nn.functional.cross_entropy(zdf, y, reduction="sum").item()
        stats[1] = torch.nn.functional.cross_entropy(z[0], y, reduction="sum").item()
        stats[2] = torch.nn.functional.cross_entropy(z[1], y, reduction="sum").item()
        s1, s2 = O.ogm_scores(z[0].detach(), z[1].detach(), y)
        stats[3], stats[4] = float(s1), float(s2)
        stats[16:16 + C] = z[0].detach().sum(0).double()
        stats[16 + C:] = z[1].detach().sum(0).double()
        # ---- exchange 1: ONE all-gather of the byte payload [stats f64 | idx i64 | conf (2,B) f32], then the
        # rank-major decode lf_step_mid performs on the device (sum of the partial statistics in rank order)
        n_stats = 16 + 2 * C
        pay = torch.empty(8 * n_stats + 8 * B + 8 * B, dtype=torch.uint8)
        pay[:8 * n_stats].view(torch.float64).copy_(stats)
        pay[8 * n_stats:8 * n_stats + 8 * B].view(torch.int64).copy_(idx)
        pay[8 * n_stats + 8 * B:].view(torch.float32).view(2, B).copy_(conf.detach())
        gathered = parallel.gather_payload(pay).view(world, -1)
        stats = torch.zeros_like(stats)
        for r in range(world):
            stats += gathered[r, :8 * n_stats].view(torch.float64)
        idx_g = torch.cat([gathered[r, 8 * n_stats:8 * n_stats + 8 * B].view(torch.int64) for r in range(world)])
        conf_g = torch.cat([gathered[r, 8 * n_stats + 8 * B:].view(torch.float32).view(2, B) for r in range(world)], dim=1)
        assert idx_g.shape == (Bg,) and conf_g.shape == (2, Bg)
        assert torch.equal(idx_g, full["idx"])

        # ---- History + regulariser on the GLOBAL batch (identical on every rank)
        hist = O.HistoryState(N)
        for m in range(2):
            O.history_update(hist, m, idx_g.numpy(), float(np.float32(stats[1 + m].item() / Bg)), conf_g[m].numpy())
        cg = conf_g.clone().requires_grad_(True)
        reg = O.qmf_reg_loss_closed(cg, idx_g.numpy(), hist)
        reg.backward()
        g_local = cg.grad[:, sl]                                         # this shard's dL_reg/dconf rows

        # ---- backward half on the shard with global denominators
        loss_local = (torch.nn.functional.cross_entropy(zdf, y, reduction="sum")
                      + torch.nn.functional.cross_entropy(z[0], y, reduction="sum")
                      + torch.nn.functional.cross_entropy(z[1], y, reduction="sum")) / Bg + (conf * g_local).sum()
        loss_local.backward()
        n = C * D
        flat = torch.zeros(2 * (n + C) + 2)
        flat[0:n] = Wl[0].grad.flatten(); flat[n:n + C] = bl[0].grad
        flat[n + C:2 * n + C] = Wl[1].grad.flatten(); flat[2 * n + C:2 * n + 2 * C] = bl[1].grad
        stats[9], stats[10] = 3.0 + rank, 5.0 + rank                     # stand-in calibrated counts
        parallel.pack_grad_exchange(flat, 2 * (n + C), stats, 9, 11)     # exchange 2
        assert stats[9].item() == 3.0 * world + sum(range(world))
        assert stats[10].item() == 5.0 * world + sum(range(world))
        total_loss = (stats[0] + stats[1] + stats[2]).item() / Bg + reg.item()
        q.put((rank, total_loss, flat[:2 * (n + C)].clone().numpy(), f[0].grad.numpy(), hist.correctness.copy(),
               stats[3].item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_qmf_equals_single_process_oracle():
    from oracle import late_fusion as O
    world, B, D, C, N = 2, 24, 32, 5, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, D, C, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0

    full = O.make_inputs(B * world, D, C, seed=3, n_data=N)
    hist = O.HistoryState(N)
    ref = O.qmf_step([full["f1"], full["f2"]], [full["W1"], full["W2"]], [full["b1"], full["b2"]], full["y"],
                     full["idx"], hist)
    n = C * D
    ref_flat = np.concatenate([ref["dW"][0].flatten().numpy(), ref["db"][0].numpy(),
                               ref["dW"][1].flatten().numpy(), ref["db"][1].numpy()])
    for rank, loss, flat, df1, corr, score1 in res:
        assert abs(loss - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
        assert np.linalg.norm(flat - ref_flat) / np.linalg.norm(ref_flat) < 1e-5
        assert np.allclose(corr, hist.correctness, rtol=1e-6, atol=1e-12)
        assert abs(score1 - ref["score1"]) < 1e-4
        want = ref["dfeat"][0][rank * B:(rank + 1) * B].numpy()
        assert np.linalg.norm(df1 - want) / np.linalg.norm(want) < 1e-5
    # both ranks hold the same (all-reduced) gradients
    assert np.array_equal(res[0][2], res[1][2])


def test_single_process_helpers_are_identity():
    from multimodal_clinical_b200 import parallel
    assert parallel.world() == (0, 1)
    t = torch.arange(4.0)
    assert parallel.allreduce_sum_(t) is t
    idx, conf = torch.arange(3), torch.zeros(2, 3)
    a, b = parallel.gather_batch(idx, conf)
    assert a is idx and b is conf
    assert parallel.shard_range(3, 128) == (384, 128)


def _ddp_layout_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_clinical_b200 import parallel
        pay = torch.full((64,), rank + 1, dtype=torch.uint8)
        ok = True
        if rank == 1:
            # only ONE rank calls: a collective would block until the timeout.  The engine's own world size decides.
            got = parallel.gather_payload(pay, None, engine_world=1)
            ok = got is pay
        dist.barrier()
        # a sharded engine (engine_world == group size, or not given) still gathers rank-major
        g = parallel.gather_payload(pay, None, engine_world=world).view(world, -1)
        ok = ok and bool((g[0] == 1).all()) and bool((g[1] == 2).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_unsharded_engine_inside_a_process_group_keeps_its_own_payload():
    """The DDP layout (LateFusionStep(sharded=False) while torch.distributed is initialised): lf_step_mid must consume THIS
    rank's statistics.  parallel.gather_payload used to key on the process group alone and handed back rank 0's row."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_layout_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
