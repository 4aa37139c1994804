"""Structure of the fused step on the GPU: CUDA-graph replay == eager, lf_step_mid == the separate entry
points it supersedes, fused narrow-head kernel == generic kernels, and size-independent properties at the
BASELINE.json full sizes (where the O(B^2) reference / oracle cannot run in seconds)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import late_fusion as O
from tests.util import TOL_FP32, assert_close

pytestmark = pytest.mark.gpu


def _eng(**kw):
    from multimodal_clinical_b200.step import LateFusionStep
    return LateFusionStep(device="cuda:0", **kw)


def _batch(B, D, Cn, seed, N=None):
    inp = O.make_inputs(B, D, Cn, seed=seed, n_data=N)
    return {k: v.cuda() for k, v in inp.items() if torch.is_tensor(v)}


@pytest.mark.parametrize("mode,B,D,Cn,N,prec", [("qmf", 96, 512, 6, 300, "fp32"), ("jlogits", 200, 256, 20, None, "fp32"),
                                               ("qmf", 300, 768, 101, 1000, "tf32")])
def test_cuda_graph_replay_matches_eager(mode, B, D, Cn, N, prec):
    batches = [_batch(B, D, Cn, 10 + s, N) for s in range(3)]
    W = [batches[0]["W1"], batches[0]["W2"]]; b = [batches[0]["b1"], batches[0]["b2"]]
    alpha = 0.8 if mode == "jlogits" else None
    # eager: the capture below runs two warm-up steps on batch 0, so the eager engine does the same
    ea = _eng(num_classes=Cn, mode=mode, n_data=N, precision=prec)
    seq = [batches[0], batches[0], batches[1], batches[2]]
    for s in seq:
        ref = ea.step([s["f1"], s["f2"]], W, b, s["y"], idx=s.get("idx"), ogm_alpha=alpha)
    ref = {"loss": ref.loss.clone(), "dW": ref.dweight[0].clone(), "df": ref.dfeat[1].clone(), "stats": ref.stats.clone()}
    # graph: static input buffers, replay after refreshing their contents
    eg = _eng(num_classes=Cn, mode=mode, n_data=N, precision=prec)
    st = {k: v.clone() for k, v in batches[0].items()}
    g, out = eg.capture([st["f1"], st["f2"]], W, b, st["y"], idx=st.get("idx"), ogm_alpha=alpha, warmup=2)
    for s in (batches[1], batches[2]):
        for k in ("f1", "f2", "y") + (("idx",) if N else ()):
            st[k].copy_(s[k])
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.loss, ref["loss"]) and torch.equal(out.dweight[0], ref["dW"]) and torch.equal(out.dfeat[1], ref["df"])
    assert torch.equal(out.stats, ref["stats"]) and torch.equal(eg.ema_x, ea.ema_x)
    if N:
        assert torch.equal(eg.correctness, ea.correctness) and torch.equal(eg.confidence, ea.confidence)
    else:
        assert torch.equal(eg.coeff, ea.coeff)


def test_step_mid_matches_the_separate_entry_points():
    """lf_step_mid (one cluster launch) vs lf_ema_update + lf_ogm_coeff + lf_qmf_history_step + lf_loss_finalize."""
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    Cn, B, N = 11, 257, 400
    g = torch.Generator().manual_seed(4)
    n_stats = _lib.LF_STATS_HEADER + 2 * Cn
    stats_part = torch.zeros(n_stats, dtype=torch.float64)
    stats_part[:5] = torch.tensor([620.0, 580.3, 601.7, 30.2, 25.9], dtype=torch.float64)
    stats_part[16:] = torch.randn(2 * Cn, generator=g, dtype=torch.float64) * B
    idx = torch.randint(0, N, (B,), generator=g)
    conf = torch.rand(2, B, generator=g)
    corr0 = torch.rand(2, N, generator=g, dtype=torch.float64) * (torch.rand(2, N, generator=g) > 0.3)

    def fresh():
        return dict(corr=corr0.clone().cuda(), confid=torch.zeros(2, N, dtype=torch.float64, device="cuda"),
                    lw=torch.zeros(N + 1, dtype=torch.int64, device="cuda"), ema_x=torch.full((2, Cn), 0.3, device="cuda"),
                    ema_off=torch.zeros(2, Cn, device="cuda"), coeff=torch.ones(2, device="cuda"),
                    loss=torch.zeros(1, device="cuda"), qmf_g=torch.zeros(2, B, device="cuda"),
                    stats=stats_part.clone().cuda())
    a, b = fresh(), fresh()
    idx_d, conf_d, part_d = idx.cuda(), conf.cuda(), stats_part.cuda()
    for rep in range(2):                                   # twice: the ticket counter must advance consistently
        # --- separate entry points
        _lib.check(lib.lf_ema_update(a["ema_x"].data_ptr(), a["ema_off"].data_ptr(), a["stats"].data_ptr(), Cn, B, 0.05, st), "ema")
        _lib.check(lib.lf_ogm_coeff(a["stats"].data_ptr(), 0.8, a["coeff"].data_ptr(), st), "coeff")
        q = _lib.LfQmfArgs()
        q.batch_global, q.n_data = B, N
        q.idx, q.conf = idx_d.data_ptr(), conf_d.data_ptr()
        q.correctness, q.confidence, q.last_writer, q.step_base = a["corr"].data_ptr(), a["confid"].data_ptr(), a["lw"].data_ptr(), 0
        q.stats, q.qmf_g, q.target_out = a["stats"].data_ptr(), a["qmf_g"].data_ptr(), None
        q.g_begin, q.g_count = 0, B
        ws = torch.empty(lib.lf_qmf_workspace_bytes(N), dtype=torch.uint8, device="cuda")
        q.workspace, q.workspace_bytes, q.flags = ws.data_ptr(), ws.numel(), _lib.LF_QMF_ALL
        _lib.check(lib.lf_qmf_history_step(C.byref(q), st), "qmf")
        _lib.check(lib.lf_loss_finalize(a["stats"].data_ptr(), _lib.LF_MODE_QMF, B, a["loss"].data_ptr(), st), "loss")
        # --- one launch
        m = _lib.LfMidArgs()
        m.mode, m.classes, m.batch_global, m.n_ranks, m.batch_local, m.rank, m.n_data, m.update_ema = _lib.LF_MODE_QMF, Cn, B, 1, B, 0, N, 1
        m.stats_parts, m.stats_stride = part_d.data_ptr(), n_stats
        m.idx_parts, m.idx_stride, m.conf_parts, m.conf_stride = idx_d.data_ptr(), B, conf_d.data_ptr(), 2 * B
        m.stats, m.ema_x, m.ema_offset, m.smoothing = b["stats"].data_ptr(), b["ema_x"].data_ptr(), b["ema_off"].data_ptr(), 0.05
        m.alpha, m.coeff_out = 0.8, b["coeff"].data_ptr()
        m.correctness, m.confidence, m.last_writer, m.step_base = b["corr"].data_ptr(), b["confid"].data_ptr(), b["lw"].data_ptr(), 0
        m.qmf_g, m.loss_out = b["qmf_g"].data_ptr(), b["loss"].data_ptr()
        ws2 = torch.zeros(lib.lf_mid_workspace_bytes(B), dtype=torch.uint8, device="cuda")     # zero-initialised once (header contract)
        m.workspace, m.workspace_bytes = ws2.data_ptr(), ws2.numel()
        _lib.check(lib.lf_step_mid(C.byref(m), st), "mid")
        torch.cuda.synchronize()
        for k in ("confid", "lw", "ema_x", "ema_off", "coeff"):
            assert torch.equal(a[k], b[k]), (rep, k)
        assert_close(b["corr"], a["corr"], 1e-14, "correctness")          # fp64 FMA contraction may differ by 1 ulp
        assert_close(b["qmf_g"], a["qmf_g"], 1e-6, "dL_reg/dconf")
        assert_close(b["loss"], a["loss"], 1e-6, "loss")
        assert_close(b["stats"][_lib.STAT["REG_SUM"]], a["stats"][_lib.STAT["REG_SUM"]], 1e-6, "reg sum")
    assert int(b["lw"][N]) == 2 * B


@pytest.mark.parametrize("mode,B,D,Cn,N", [("jlogits", 777, 512, 6, None), ("qmf", 515, 768, 20, 900), ("jlogits", 64, 36, 1, None),
                                           ("qmf", 130, 1024, 32, 300)])
def test_fused_narrow_kernel_matches_generic_kernels(mode, B, D, Cn, N):
    s = _batch(B, D, Cn, 21, N)
    outs = []
    for generic in (False, True):
        if generic:
            os.environ["LF_NO_NARROW"] = "1"
        try:
            e = _eng(num_classes=Cn, mode=mode, n_data=N)
            for _ in range(2):
                o = e.step([s["f1"], s["f2"]], [s["W1"], s["W2"]], [s["b1"], s["b2"]], s["y"], idx=s.get("idx"), ogm_alpha=0.5)
            torch.cuda.synchronize()
            outs.append(dict(loss=o.loss.clone(), z=o.logits[1].clone(), dW=o.dweight[1].clone(), db=o.dbias[0].clone(),
                             df=o.dfeat[0].clone(), stats=o.stats.clone(), ema=e.ema_x.clone()))
        finally:
            os.environ.pop("LF_NO_NARROW", None)
    for k in outs[0]:
        assert_close(outs[0][k], outs[1][k], 2e-6, k)
    assert torch.equal(outs[0]["stats"][5:11], outs[1]["stats"][5:11])        # accuracy counts are exact


@pytest.mark.parametrize("name,mode,B,D,Cn,N,prec", [("K3", "jlogits", 8192, 512, 6, None, "fp32"),
                                                     ("K4", "qmf", 32768, 768, 101, 65536, "tf32"),
                                                     ("K2", "qmf", 64, 512, 6, 6698, "fp32")])
def test_full_size_properties(name, mode, B, D, Cn, N, prec):
    """BASELINE.json sizes.  Properties that hold for any input: softmax-CE gradients sum to zero over classes
    (mean fusion), or to the ranking gradient / 10 (QMF: d/dz of conf = softmax / 10); db is the column sum of dz;
    counts are integers in [0, B]; the step is bit-reproducible."""
    tol = 1e-4 if prec == "fp32" else 2e-2
    s = _batch(B, D, Cn, 5, N)
    e = _eng(num_classes=Cn, mode=mode, n_data=N, precision=prec)
    runs = []
    for rep in range(2):
        if N:
            e.qmf_state.correctness.zero_(); e.qmf_state.confidence.zero_(); e.qmf_state.last_writer.zero_()
        e.ema_x.zero_(); e.ema_offset.zero_()
        o = e.step([s["f1"], s["f2"]], [s["W1"], s["W2"]], [s["b1"], s["b2"]], s["y"], idx=s.get("idx"))
        torch.cuda.synchronize()
        runs.append([o.loss.clone(), o.dweight[0].clone(), o.dfeat[1].clone(), o.dbias[1].clone()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)                                             # deterministic reductions
    loss, dW, df, db = runs[1]
    assert torch.isfinite(loss) and float(loss) > 0
    counts = o.stats[5:11].cpu()
    assert torch.all(counts == counts.round()) and torch.all(counts >= 0) and torch.all(counts <= B)
    bufs = e._bufs
    dz = bufs["dz"][:, :, :Cn].double()
    if mode == "jlogits":
        assert float(dz[0].sum(1).abs().max()) < 1e-6 / B * 50               # rows of softmax - onehot sum to 0
        want_db = dz[0].sum(0)
        assert_close(o.dbias[0], want_db, 1e-4, "db1"); assert_close(o.dbias[1], want_db, 1e-4, "db2")
    else:
        g = bufs["qmf_g"].double() / 10.0
        for m in range(2):
            assert_close(dz[m].sum(1), g[m], 1e-3, f"row sums of dz{m + 1} == dL_reg/dconf / 10")
            assert_close(o.dbias[m], dz[m].sum(0), 1e-4, f"db{m + 1}")
    # dW = dz^T f and df = dz W, checked on a random probe (no full fp64 GEMM at this size)
    probe = torch.randn(D, dtype=torch.float64, device="cuda")
    m = 0
    want = dz[m if mode == "qmf" else 0].t() @ (s["f1"].double() @ probe)
    assert_close(o.dweight[0].double() @ probe, want, tol, "dW1 . probe")
    dzm = dz[1 if mode == "qmf" else 0]
    assert_close(o.dfeat[1].double() @ probe, dzm @ (s["W2"].double() @ probe), tol, "df2 . probe")


def test_stream_from_host_matches_eager_steps():
    """The host-fed pipeline (pinned batches -> copy stream -> graph replay, loss read back every step) yields the
    same losses and final state as eager steps on the same batches (the two capture warm-ups replay batch 0)."""
    B, D, Cn, N = 128, 512, 6, 400
    host = []
    for s in range(4):
        inp = O.make_inputs(B, D, Cn, seed=40 + s, n_data=N)
        host.append({k: inp[k].pin_memory() for k in ("f1", "f2", "y", "idx")})
    base = O.make_inputs(B, D, Cn, seed=1, n_data=N)
    W = [base["W1"].cuda(), base["W2"].cuda()]; b = [base["b1"].cuda(), base["b2"].cuda()]
    ea = _eng(num_classes=Cn, mode="qmf", n_data=N)
    # capture of the two staging sets runs 2 warm-up steps each on batch 0 -> 4 extra steps on batch 0
    want = []
    for hb in [host[0]] * 4 + host:
        o = ea.step([hb["f1"].cuda(), hb["f2"].cuda()], W, b, hb["y"].cuda(), idx=hb["idx"].cuda())
        want.append(float(o.loss))
    eg = _eng(num_classes=Cn, mode="qmf", n_data=N)
    got = list(eg.stream_from_host(iter(host), W, b))
    assert got == want[4:]
    assert torch.equal(eg.correctness, ea.correctness) and torch.equal(eg.ema_x, ea.ema_x)


def test_fused_sgd_in_the_dw_tail_matches_torch_sgd():
    """SGD(momentum 0.9, wd 1e-4) applied inside the tail of the dW kernel == torch.optim.SGD fed
    with the gradients the step returns (utils/BaseModel.py:275-285), including the bf16 copies the next forward uses."""
    B, D, Cn, N = 2048, 768, 101, 5000
    s = _batch(B, D, Cn, 11, N)
    f = [s["f1"].bfloat16(), s["f2"].bfloat16()]
    Wa = [s["W1"].clone(), s["W2"].clone()]; ba = [s["b1"].clone(), s["b2"].clone()]
    Wb = [s["W1"].clone().requires_grad_(True), s["W2"].clone().requires_grad_(True)]
    bb = [s["b1"].clone().requires_grad_(True), s["b2"].clone().requires_grad_(True)]
    ea = _eng(num_classes=Cn, mode="qmf", n_data=N, precision="bf16")
    ea.enable_sgd(lr=0.05, momentum=0.9, weight_decay=1e-4)
    eb = _eng(num_classes=Cn, mode="qmf", n_data=N, precision="bf16")
    opt = torch.optim.SGD(Wb + bb, lr=0.05, momentum=0.9, weight_decay=1e-4)
    for step in range(4):
        if step == 2:
            ea.set_lr(0.025)
            opt.param_groups[0]["lr"] = 0.025
        oa = ea.step(f, Wa, ba, s["y"], idx=s["idx"])
        ob = eb.step(f, [w.detach() for w in Wb], [x.detach() for x in bb], s["y"], idx=s["idx"])
        torch.cuda.synchronize()
        assert_close(oa.loss, ob.loss, 1e-6, f"loss step {step}")          # same heads -> same step
        assert_close(oa.dweight[1], ob.dweight[1], 1e-6, "dW2")
        for p, g in zip(Wb + bb, ob.dweight + ob.dbias):
            p.grad = g.clone()
        opt.step()
        for m in range(2):
            assert_close(Wa[m], Wb[m], 1e-6, f"W{m + 1} after step {step}")
            assert_close(ba[m], bb[m], 1e-6, f"b{m + 1} after step {step}")
            assert torch.equal(ea._w16[m], Wa[m].bfloat16())              # the copies the next forward consumes


def test_dw_tail_equals_separate_finalize_launch():
    B, D, Cn, N = 4096, 768, 101, 9000
    s = _batch(B, D, Cn, 3, N)
    f = [s["f1"].bfloat16(), s["f2"].bfloat16()]
    outs = []
    for tail in (True, False):
        if not tail:
            os.environ["LF_NO_DW_TAIL"] = "1"
        try:
            e = _eng(num_classes=Cn, mode="qmf", n_data=N, precision="bf16")
            o = e.step(f, [s["W1"], s["W2"]], [s["b1"], s["b2"]], s["y"], idx=s["idx"])
            torch.cuda.synchronize()
            outs.append([o.dweight[0].clone(), o.dweight[1].clone(), o.dbias[0].clone(), o.dbias[1].clone(), o.stats.clone()])
        finally:
            os.environ.pop("LF_NO_DW_TAIL", None)
    for a, b in zip(*outs):
        assert_close(a, b, 1e-6, "tail vs finalize_grads")
    assert torch.equal(outs[0][4][5:11], outs[1][4][5:11])
