"""Times the C-ABI calls of one step in isolation: N back-to-back calls of each, CUDA events around the
loop (no per-kernel event overhead), inputs rotating over sets larger than L2.
  python tools/bench_parts.py [--workload k4] [--precision tf32] [--iters 20]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_clinical_b200 import _lib
from multimodal_clinical_b200.step import LateFusionStep, _ptr, _stream

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="k4"); ap.add_argument("--precision", default="tf32"); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--batch", type=int); ap.add_argument("--dim", type=int); ap.add_argument("--classes", type=int)
a = ap.parse_args()
w = dict(bench.WORKLOADS[a.workload])
w.update(B=a.batch or w["B"], D=a.dim or w["D"], C=a.classes or w["C"])
dev = torch.device("cuda:0")
lib = _lib.load()
eng = LateFusionStep(w["C"], mode=w["mode"], n_data=w["N"], device=dev, precision=a.precision)
W, b = bench.head_params(w); W = [x.to(dev) for x in W]; b = [x.to(dev) for x in b]
n_sets = max(2, -(-2 * bench.L2_BYTES // (2 * w["B"] * w["D"] * 4)))
sets = [{k: v.to(dev) for k, v in s.items()} for s in bench.make_batches(w, n_sets, dev, 1)]
# run full steps once per set so every buffer / state exists, then capture the argument structs
calls = {}
orig = {n: getattr(lib, n) for n in ("lf_heads_forward", "lf_heads_backward", "lf_qmf_history_step")}
for s in sets:
    eng.step([s["f1"], s["f2"]], W, b, s["y"], idx=s.get("idx"), need_dfeat=w["dfeat"], ogm_alpha=w["alpha"])
torch.cuda.synchronize()

def args_for(s):
    B, D = s["f1"].shape; Cn = w["C"]; bufs = eng._buffers(B, D, w["dfeat"]); n = Cn * D; gf = bufs["grad_flat"]
    h = _lib.LfHeadsArgs()
    h.batch, h.batch_global, h.dim, h.classes = B, B, D, Cn
    h.mode, h.precision, h.need_dfeat, h.ld_dlogits = eng.mode, eng.precision, int(w["dfeat"]), bufs["ldz"]
    h.ld_logits = bufs["ldl"]; h.ld_fused = bufs["ldf"]
    for m, f in enumerate((s["f1"], s["f2"])):
        h.feat[m] = _ptr(f); h.weight[m] = _ptr(W[m]); h.bias[m] = _ptr(b[m]); h.logits[m] = _ptr(bufs["logits_store"][m])
        h.dfeat[m] = _ptr(bufs["dfeat"][m]) if w["dfeat"] else None
    h.dweight[0] = _ptr(gf[0:n]); h.dbias[0] = _ptr(gf[n:n + Cn]); h.dweight[1] = _ptr(gf[n + Cn:2 * n + Cn]); h.dbias[1] = _ptr(gf[2 * n + Cn:])
    h.label = _ptr(s["y"]); h.avg_logits = _ptr(bufs["avg_store"]); h.logits_df = _ptr(bufs["zdf_store"]); h.conf = _ptr(bufs["conf"])
    h.dlogits[0] = _ptr(bufs["dz"][0]); h.dlogits[1] = _ptr(bufs["dz"][1]) if eng.mode == 1 else None
    h.qmf_g = _ptr(bufs["qmf_g"]); h.ema_offset = _ptr(eng.ema_offset); h.stats = _ptr(eng.stats)
    h.workspace = _ptr(eng._ws); h.workspace_bytes = eng._ws.numel()
    return h
hs = [args_for(s) for s in sets]

def timeit(name, fn, bytes_per_call=None):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.lf_launch_count()
    e0.record()
    for i in range(a.iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.iters
    nl = (lib.lf_launch_count() - l0) / a.iters
    extra = f"  {bytes_per_call / us / 1e3:7.0f} GB/s" if bytes_per_call else ""
    print(f"{name:28s} {us:8.1f} us/call  ({nl:.0f} launches){extra}", flush=True)

B, D, Cn = w["B"], w["D"], w["C"]
nout = 4 if w["mode"] == "qmf" else 3
st = _stream()
timeit("lf_heads_forward", lambda i: _lib.check(lib.lf_heads_forward(C.byref(hs[i % n_sets]), st), "fwd"), B * (2 * D + nout * Cn) * 4)
timeit("lf_heads_backward", lambda i: _lib.check(lib.lf_heads_backward(C.byref(hs[i % n_sets]), st), "bwd"), B * (4 * D) * 4)
timeit("full step", lambda i: eng.step([sets[i % n_sets]["f1"], sets[i % n_sets]["f2"]], W, b, sets[i % n_sets]["y"], idx=sets[i % n_sets].get("idx"), need_dfeat=w["dfeat"], ogm_alpha=w["alpha"]))
