"""Parity of the batch-sharded step: N ranks must reproduce ONE GPU running the concatenated batch, and the CPU oracle.

Used by ``bench.py`` (``parity_check`` in the JSON line, before anything is timed) and, under torchrun, by
``tests/test_multigpu_gpu.py``:

    torchrun --nproc-per-node 2 tools/parity_multigpu.py [--workload k4] [--batch 4096]

The global batch is generated on the device from one seed (identical on every rank); rank r feeds rows
[r*B, (r+1)*B) to the sharded engine and ALL rows to a second, un-sharded engine on the same GPU.  After `steps`
steps (with the fused SGD update when it is enabled) the two must agree: replicated state (EMA, QMF History,
statistics, heads) to fp32 rounding of the batch means, gradients to the rounding of a different summation
order, integer counts up to ties.  A second, small problem is compared with oracle/late_fusion.py (fp64 on the
same bf16-rounded inputs, tolerance 2e-2 / 1e-5 as in BASELINE.json) on rank 0.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def _rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    same_nan = torch.isnan(a) & torch.isnan(b)                       # NaN in the same place on both sides is agreement
    a, b = torch.where(same_nan, torch.zeros_like(a), a), torch.where(same_nan, torch.zeros_like(b), b)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _global_batch(w, Bg, dev, seed, bf16, step):
    g = torch.Generator(device=dev).manual_seed(seed + 977 * step)
    f1 = torch.randn(Bg, w["D"], generator=g, device=dev)
    f2 = torch.randn(Bg, w["D"], generator=g, device=dev)
    if bf16:
        f1, f2 = f1.bfloat16(), f2.bfloat16()
    y = torch.randint(0, w["C"], (Bg,), generator=g, device=dev, dtype=torch.int64)
    idx = None
    if w["N"]:
        if Bg <= 8192:       # sampling with replacement: duplicates inside and across shards (cremad/get_data.py:157)
            idx = torch.randint(0, w["N"], (Bg,), generator=g, device=dev, dtype=torch.int64)
        else:                # contiguous window (food101/run_training.py:39-45)
            idx = (torch.arange(Bg, device=dev, dtype=torch.int64) + 12345 * (step + 1)) % w["N"]
    return f1, f2, y, idx


def _heads(w, dev, seed=5):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (w["D"] ** 0.5)
    mk = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * bound).to(dev)
    return [mk(w["C"], w["D"]), mk(w["C"], w["D"])], [mk(w["C"]), mk(w["C"])]


def sharded_vs_single(w, precision, dev, rank, world, sgd=False, steps=2, seed=4242, comm="auto"):
    """-> dict of error measures (max over steps) between the N-rank engine and a 1-GPU engine on the whole batch."""
    from multimodal_clinical_b200.step import LateFusionStep
    from multimodal_clinical_b200._lib import STAT
    B, Bg = w["B"], w["B"] * world
    if w["N"]:
        # the dataset grows with the global batch: a batch that covers every History entry leaves max == min and the
        # reference's normalisation yields NaN -- legal, but it would hide the ranking terms from the check
        w = dict(w, N=max(w["N"], 2 * Bg + 1))
    bf16 = precision == "bf16"
    kw = dict(mode=w["mode"], n_data=w["N"], device=dev, precision=precision)
    eng_s = LateFusionStep(w["C"], comm=comm, **kw)
    eng_1 = LateFusionStep(w["C"], sharded=False, **kw)
    Ws, bs = _heads(w, dev)
    W1, b1 = [x.clone() for x in Ws], [x.clone() for x in bs]
    if sgd:
        eng_s.enable_sgd(lr=1e-2, momentum=0.9, weight_decay=1e-4)
        eng_1.enable_sgd(lr=1e-2, momentum=0.9, weight_decay=1e-4)
    sl = slice(rank * B, (rank + 1) * B)
    err = {}

    def upd(k, v):
        v = float(v)
        err[k] = max(err.get(k, 0.0), v if v == v else float("inf"))      # NaN on one side only is a failure

    for s in range(steps):
        f1, f2, y, idx = _global_batch(w, Bg, dev, seed, bf16, s)
        o1 = eng_1.step([f1, f2], W1, b1, y, idx=idx, need_dfeat=w["dfeat"], ogm_alpha=w["alpha"])
        os_ = eng_s.step([f1[sl], f2[sl]], Ws, bs, y[sl], idx=idx[sl] if idx is not None else None, need_dfeat=w["dfeat"],
                         ogm_alpha=w["alpha"])
        torch.cuda.synchronize()
        eng_s.check_peer()
        upd("loss", abs(float(os_.loss) - float(o1.loss)) / max(abs(float(o1.loss)), 1e-30))
        if os.environ.get("LF_PARITY_VERBOSE"):
            print(f"[rank {rank}] step {s}: loss single {float(o1.loss):.6g} sharded {float(os_.loss):.6g}  header single "
                  f"{[round(float(v), 3) for v in o1.stats[:5]]} sharded {[round(float(v), 3) for v in os_.stats[:5]]}", flush=True)
        for m in range(2):
            upd("logits", _rel(os_.logits[m], o1.logits[m][sl]))
            upd("dweight", _rel(os_.dweight[m], o1.dweight[m]))
            upd("dbias", _rel(os_.dbias[m], o1.dbias[m]))
            if w["dfeat"]:
                upd("dfeat", _rel(os_.dfeat[m].float(), o1.dfeat[m][sl].float()))
        st_s, st_1 = os_.stats.cpu(), o1.stats.cpu()
        for k in ("CE_JOINT", "CE_X1", "CE_X2", "SCORE_X1", "SCORE_X2") + (("REG_SUM",) if w["mode"] == "qmf" else ()):
            upd("stats", abs(float(st_s[STAT[k]]) - float(st_1[STAT[k]])) / max(abs(float(st_1[STAT[k]])), 1e-30))
        cnt = 0.0
        for k in ("CNT_X1", "CNT_X2", "CNT_JOINT", "CNT_X1_CAL", "CNT_X2_CAL") + (("CNT_DF",) if w["mode"] == "qmf" else ()):
            cnt = max(cnt, abs(float(st_s[STAT[k]]) - float(st_1[STAT[k]])))
        upd("counts_abs", cnt)
        upd("ema_x", _rel(eng_s.ema_x, eng_1.ema_x))
        if w["mode"] == "qmf":
            upd("history_correctness", _rel(eng_s.correctness, eng_1.correctness))
            upd("history_confidence", _rel(eng_s.confidence, eng_1.confidence))
        if w["alpha"] is not None:
            upd("ogm_coeff", float((eng_s.coeff - eng_1.coeff).abs().max()))
        if sgd:
            upd("heads_after_sgd", max(_rel(Ws[0], W1[0]), _rel(Ws[1], W1[1]), _rel(bs[0], b1[0]), _rel(bs[1], b1[1])))
    # every rank must hold the same replicated state, bit for bit
    rep = {"ema_x": eng_s.ema_x, "stats": eng_s.stats, "loss": os_.loss.reshape(1), "W1": Ws[0], "W2": Ws[1], "b1": bs[0], "b2": bs[1],
           "dW1": os_.dweight[0], "db2": os_.dbias[1]}
    if w["mode"] == "qmf":
        rep.update(correctness=eng_s.correctness, confidence=eng_s.confidence)
    if world > 1:
        differing = []
        for name, t in rep.items():
            mine = t.detach().double().flatten().contiguous().view(torch.int64)      # bit patterns: NaN == NaN
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            if not all(torch.equal(allr[0], x) for x in allr):
                bad = (torch.stack(allr) != allr[0]).any(0).nonzero().flatten()[:4].tolist()
                differing.append(f"{name}[{','.join(str(i) for i in bad)}]: " +
                                 " vs ".join(repr([float(x.view(torch.float64)[i]) for i in bad]) for x in allr))
        err["ranks_bit_identical"] = not differing
        if differing:
            err["differs_across_ranks"] = differing
        keys = [k for k in sorted(err) if k not in ("ranks_bit_identical", "differs_across_ranks")]
        t = torch.tensor([err[k] for k in keys], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for k, v in zip(keys, t.tolist()):
            err[k] = v
    else:
        err["ranks_bit_identical"] = True
    err["comm"] = "peer" if eng_s.peer is not None else ("nccl" if world > 1 else "none")
    eng_s.close()
    return err


def sharded_vs_oracle(w, precision, dev, rank, world, B_small=512, seed=777):
    """Small problem (global batch <= 4096): the N-rank engine against the fp64 CPU oracle -> dict of errors (rank 0)."""
    from multimodal_clinical_b200.step import LateFusionStep
    from oracle import late_fusion as O
    ws = dict(w, B=B_small, N=(4099 if w["N"] else None))
    Bg = B_small * world
    bf16 = precision == "bf16"
    eng = LateFusionStep(ws["C"], mode=ws["mode"], n_data=ws["N"], device=dev, precision=precision)
    W, b = _heads(ws, dev)
    rnd = (lambda x: x.bfloat16().float()) if bf16 else (lambda x: x.float())
    Wc, bc = [rnd(x).cpu() for x in W], [x.cpu() for x in b]
    hist = O.HistoryState(ws["N"]) if ws["N"] else None
    ema = torch.zeros(2, ws["C"], dtype=torch.float64)
    sl = slice(rank * B_small, (rank + 1) * B_small)
    err = {}
    for s in range(2):
        f1, f2, y, idx = _global_batch(ws, Bg, dev, seed, bf16, s)
        out = eng.step([f1[sl], f2[sl]], W, b, y[sl], idx=idx[sl] if idx is not None else None, need_dfeat=True, ogm_alpha=ws["alpha"])
        torch.cuda.synchronize()
        if rank != 0:
            continue
        fc = [f1.float().cpu(), f2.float().cpu()]
        if ws["mode"] == "qmf":
            ref = O.qmf_step(fc, Wc, bc, y.cpu(), idx.cpu(), hist, ema_x=ema, dtype=torch.float64)
        else:
            ref = O.jlogits_step(fc, Wc, bc, y.cpu(), ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        e = {"loss": abs(float(out.loss) - float(ref["loss"])) / abs(float(ref["loss"])),
             "logits": _rel(out.logits[0].cpu(), ref["logits"][0][sl]), "dweight": _rel(out.dweight[1].cpu(), ref["dW"][1]),
             "dbias": _rel(out.dbias[0].cpu(), ref["db"][0]), "dfeat": _rel(out.dfeat[0].float().cpu(), ref["dfeat"][0][sl]),
             "ema_x": _rel(eng.ema_x.cpu(), ref["ema_x"])}
        if ws["mode"] == "qmf":
            e["history_correctness"] = _rel(eng.correctness.cpu(), torch.as_tensor(hist.correctness))
        for k, v in e.items():
            err[k] = max(err.get(k, 0.0), v)
    eng.close()
    return err


def unsharded_in_group_vs_oracle(w, precision, dev, rank, world, B_small=384, seed=991):
    """The DDP layout (FusedLateFusionHead under a DDP wrapper): ``sharded=False`` engines inside an initialised process
    group, every rank with a DIFFERENT batch.  Each must behave like a lone GPU on its own batch -- loss, statistics, EMA and
    (QMF) the ranking gradients from its own data, never from another rank's -- checked against the CPU oracle on every rank."""
    from multimodal_clinical_b200.step import LateFusionStep
    from oracle import late_fusion as O
    ws = dict(w, B=B_small, N=(4099 if w["N"] else None))
    bf16 = precision == "bf16"
    eng = LateFusionStep(ws["C"], mode=ws["mode"], n_data=ws["N"], device=dev, precision=precision, sharded=False)
    W, b = _heads(ws, dev)
    rnd = (lambda x: x.bfloat16().float()) if bf16 else (lambda x: x.float())
    Wc, bc = [rnd(x).cpu() for x in W], [x.cpu() for x in b]
    hist = O.HistoryState(ws["N"]) if ws["N"] else None
    ema = torch.zeros(2, ws["C"], dtype=torch.float64)
    err = {}
    for s in range(2):
        f1, f2, y, idx = _global_batch(ws, B_small, dev, seed + 31 * rank, bf16, s)       # rank-specific data
        out = eng.step([f1, f2], W, b, y, idx=idx, need_dfeat=True, ogm_alpha=ws["alpha"])
        torch.cuda.synchronize()
        fc = [f1.float().cpu(), f2.float().cpu()]
        if ws["mode"] == "qmf":
            ref = O.qmf_step(fc, Wc, bc, y.cpu(), idx.cpu(), hist, ema_x=ema, dtype=torch.float64)
        else:
            ref = O.jlogits_step(fc, Wc, bc, y.cpu(), ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        e = {"loss": abs(float(out.loss) - float(ref["loss"])) / abs(float(ref["loss"])),
             "dweight": _rel(out.dweight[1].cpu(), ref["dW"][1]), "dfeat": _rel(out.dfeat[0].float().cpu(), ref["dfeat"][0]),
             "ema_x": _rel(eng.ema_x.cpu(), ref["ema_x"])}
        for k, v in e.items():
            err[k] = max(err.get(k, 0.0), v if v == v else float("inf"))
    if world > 1:
        keys = sorted(err)
        t = torch.tensor([err[k] for k in keys], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = dict(zip(keys, t.tolist()))
    return err


def run(w, precision, dev, rank, world, sgd):
    """Both checks -> the ``parity_check`` object of bench.py's JSON line (identical on every rank after the reductions)."""
    tol = 1e-5 if precision == "fp32" else 2e-2
    vs1 = sharded_vs_single(w, precision, dev, rank, world, sgd=sgd)
    vso = sharded_vs_oracle(w, precision, dev, rank, world)
    vsd = unsharded_in_group_vs_oracle(w, precision, dev, rank, world)
    # N ranks vs one GPU: same kernels, different tiling / summation order -> far inside the oracle tolerance
    lim = {"loss": 1e-5, "logits": 1e-6, "dweight": 2e-3 if precision != "fp32" else 1e-5, "dbias": 2e-3 if precision != "fp32" else 1e-5,
           "dfeat": 1e-2 if precision == "bf16" else 1e-5, "stats": 1e-5, "counts_abs": 4.0, "ema_x": 1e-5, "history_correctness": 1e-6,
           "history_confidence": 1e-6, "ogm_coeff": 1e-5, "heads_after_sgd": 1e-4}
    ok1 = vs1["ranks_bit_identical"] and all(v <= lim[k] for k, v in vs1.items() if k in lim)
    oko = all(v <= tol for v in vso.values())
    okd = all(v <= tol for v in vsd.values())
    flag = torch.tensor([1 if (ok1 and okd and (oko or rank != 0)) else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(int(flag.item())), "ranks": world, "batch_per_rank": w["B"], "steps": 2, "fused_sgd": bool(sgd),
            "vs_single_gpu_on_concatenated_batch": vs1, "vs_cpu_oracle_small": {k: float(v) for k, v in vso.items()},
            "unsharded_engines_in_the_group_vs_cpu_oracle": {k: float(v) for k, v in vsd.items()}, "oracle_tolerance": tol}


if __name__ == "__main__":
    import argparse
    import json
    from bench import WORKLOADS
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="k4")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--no-sgd", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    prec = args.precision if args.precision != "auto" else ("bf16" if w["C"] >= 32 else "fp32")
    sgd = (not args.no_sgd) and prec != "fp32" and w["C"] >= 32
    res = run(w, prec, dev, rank, world, sgd)
    if rank == 0:
        print(json.dumps({"parity_check": res}), flush=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if res["ok"] else 1)
