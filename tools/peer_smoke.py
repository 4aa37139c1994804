"""2+ ranks under torchrun: the peer-memory exchange path must reproduce the NCCL path (state bit for bit, gradients to
rounding: NCCL sums in its own order beyond 2 ranks) and leave every rank with identical results.  Run:  torchrun --nproc-per-node 2 tools/peer_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from multimodal_clinical_b200.step import LateFusionStep
from oracle import late_fusion as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for mode, B, D, Cn, N, prec in (("qmf", 96, 512, 6, 500, "fp32"), ("jlogits", 64, 256, 20, None, "fp32"), ("qmf", 256, 768, 101, 2000, "tf32")):
    Bg = B * world
    full = O.make_inputs(Bg, D, Cn, seed=3, n_data=N)
    sl = slice(rank * B, (rank + 1) * B)
    W = [full["W1"].to(dev), full["W2"].to(dev)]; b = [full["b1"].to(dev), full["b2"].to(dev)]
    res = {}
    for comm in ("nccl", "peer"):
        eng = LateFusionStep(Cn, mode=mode, n_data=N, device=dev, precision=prec, comm=comm)
        for s in range(3):
            inp = O.make_inputs(Bg, D, Cn, seed=10 + s, n_data=N)
            out = eng.step([inp["f1"][sl].to(dev), inp["f2"][sl].to(dev)], W, b, inp["y"][sl].to(dev),
                           idx=inp["idx"][sl].to(dev) if N else None, ogm_alpha=0.8 if not N else None)
        torch.cuda.synchronize()
        if comm == "peer":
            assert eng.peer is not None, "peer communicator was not built"
            eng.peer.check()
        res[comm] = [out.loss.clone(), out.dweight[0].clone(), out.dbias[1].clone(), out.stats.clone(), eng.ema_x.clone()] + \
                    ([eng.correctness.clone()] if N else [eng.coeff.clone()])
    # statistics, EMA and History are summed in rank order on both paths (bit-identical); the gradient all-reduce
    # of NCCL uses its own summation tree for > 2 ranks, so gradients agree to rounding
    def rel(a, c):
        a, c = a.double(), c.double()
        return float((a - c).norm() / c.norm().clamp_min(1e-30))
    exact = all(torch.equal(a, c) for a, c in zip(res["nccl"][3:], res["peer"][3:]))
    close = max(rel(a, c) for a, c in zip(res["nccl"][:3], res["peer"][:3]))
    # all ranks hold identical results on the peer path
    mine = torch.cat([t.double().flatten() for t in res["peer"]])
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    identical = all(torch.equal(allr[0], t) for t in allr)
    same = exact and close < 1e-5 and identical
    if rank == 0:
        print(mode, Cn, "state bit-identical:", exact, "grad rel diff vs nccl: %.2e" % close, "ranks identical:", identical,
              "loss", float(res["peer"][0]), flush=True)
    ok = ok and same
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("PEER SMOKE", "OK" if ok else "MISMATCH", flush=True)
os._exit(0 if ok else 1)
