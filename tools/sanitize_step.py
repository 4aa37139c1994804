"""Small shapes through every kernel family, for compute-sanitizer memcheck:  compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_clinical_b200.step import LateFusionStep
from oracle import late_fusion as O

cases = [("jlogits", 70, 512, 6, None, "fp32"), ("qmf", 70, 512, 6, 100, "fp32"), ("jlogits", 33, 36, 20, None, "fp32"),
         ("qmf", 200, 768, 101, 500, "tf32"), ("jlogits", 150, 512, 309, None, "tf32"), ("qmf", 200, 768, 101, 500, "bf16"),
         ("jlogits", 150, 512, 309, None, "bf16"), ("qmf", 90, 128, 40, 300, "fp32")]
for mode, B, D, Cn, N, prec in cases:
    inp = O.make_inputs(B, D, Cn, seed=1, n_data=N)
    e = LateFusionStep(Cn, mode=mode, n_data=N, device="cuda:0", precision=prec)
    for s in range(2):
        o = e.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()],
                   inp["y"].cuda(), idx=inp["idx"].cuda() if N else None, ogm_alpha=0.5 if not N else None)
    torch.cuda.synchronize()
    print(mode, B, D, Cn, prec, "loss", float(o.loss), flush=True)
# round-2 paths: in-step SGD (dW tail), ensemble loss, a QMF global batch larger than the step_mid grid, pooling, epoch-end kernels
inp = O.make_inputs(300, 768, 101, seed=2, n_data=700)
e = LateFusionStep(101, mode="qmf", n_data=700, device="cuda:0", precision="bf16")
e.enable_sgd(lr=1e-2)
W = [inp["W1"].cuda(), inp["W2"].cuda()]; b = [inp["b1"].cuda(), inp["b2"].cuda()]
for s in range(2):
    o = e.step([inp["f1"].cuda(), inp["f2"].cuda()], W, b, inp["y"].cuda(), idx=inp["idx"].cuda())
torch.cuda.synchronize(); print("in-step sgd loss", float(o.loss), flush=True)
inp = O.make_inputs(70, 512, 6, seed=3)
e = LateFusionStep(6, mode="ensemble", device="cuda:0")
o = e.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), ogm_alpha=0.5)
torch.cuda.synchronize(); print("ensemble loss", float(o.loss), flush=True)
inp = O.make_inputs(1500, 64, 6, seed=4, n_data=100)          # N small -> one-CTA step_mid grid, 3 batch positions per thread
e = LateFusionStep(6, mode="qmf", n_data=100, device="cuda:0")
o = e.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), idx=inp["idx"].cuda())
torch.cuda.synchronize(); print("looping step_mid loss", float(o.loss), flush=True)
from multimodal_clinical_b200.cremad._pool import pool_features
a = torch.randn(3, 130, 7, 7, device="cuda", requires_grad=True); v = torch.randn(9, 130, 7, 7, device="cuda", requires_grad=True)
pa, pv = pool_features(a, v); (pa.sum() + pv.sum()).backward()
from multimodal_clinical_b200.utils.BaseModel import epoch_offset_correction
off, acc = epoch_offset_correction(torch.randn(333, 2, 11, device="cuda"), torch.randint(0, 11, (333,), device="cuda"))
torch.cuda.synchronize(); print("pool / epoch ok", acc.tolist(), flush=True)
g = [torch.randn(64, 3, 7, 7, device="cuda"), torch.randn(128, 64, 3, 3, device="cuda"), torch.randn(5, 5, 1, 1, device="cuda")]
e = LateFusionStep(6, mode="jlogits", device="cuda:0")
e.modulate(g, which=0, modulation="OGM_GE", seed=1, offset=0)
torch.cuda.synchronize()
print("modulate ok", flush=True)
# late round 2: exact-fp32 wide heads (3xTF32 split-K chunks with TMA reduce-adds), mean fusion beside the dfeat GEMM
# (bwd_phase 3 / 2 / 4 on two streams), multi-head kernel (3 modalities, unequal widths), fused hidden layers
for mode, B, D, Cn, N in [("qmf", 300, 768, 101, 600), ("jlogits", 260, 512, 309, None), ("jlogits", 1100, 1024, 40, None)]:
    inp = O.make_inputs(B, D, Cn, seed=6, n_data=N)
    e = LateFusionStep(Cn, mode=mode, n_data=N, device="cuda:0", precision="fp32")
    for s in range(2):
        o = e.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()],
                   inp["y"].cuda(), idx=inp["idx"].cuda() if N else None, ogm_alpha=0.5 if not N else None)
    torch.cuda.synchronize(); print("3xTF32", mode, B, D, Cn, "loss", float(o.loss), flush=True)
from multimodal_clinical_b200.multi import MultiHeadStep
for Cn, dims, B in [(2, (100, 100, 100), 77), (10, (48, 192), 130), (32, (36, 20, 8, 52), 65)]:
    g = torch.Generator().manual_seed(3)
    f = [torch.randn(B, d, generator=g).cuda() for d in dims]
    W = [(torch.randn(Cn, d, generator=g) * 0.1).cuda() for d in dims]
    b = [torch.zeros(Cn).cuda() for _ in dims]
    o = MultiHeadStep(Cn, device="cuda:0").step(f, W, b, torch.randint(0, Cn, (B,), generator=g).cuda())
    torch.cuda.synchronize(); print("multi heads", Cn, dims, "loss", float(o.loss), flush=True)
from multimodal_clinical_b200.hidden import FusedHiddenPair
for prec, B, Din, Dout in [("bf16", 200, 768, 512), ("tf32", 130, 512, 512), ("fp32", 90, 136, 72)]:
    l1, l2 = torch.nn.Linear(Din, Dout).cuda(), torch.nn.Linear(Din, Dout).cuda()
    hp = FusedHiddenPair(0.2, precision=prec).train()
    x1 = torch.randn(B, Din, device="cuda", requires_grad=True); x2 = torch.randn(B, Din, device="cuda", requires_grad=True)
    xs = (x1.bfloat16(), x2.bfloat16()) if prec == "bf16" else (x1, x2)
    h1, h2 = hp(xs[0], xs[1], l1, l2)
    (h1.float().sum() + h2.float().sum()).backward()
    torch.cuda.synchronize(); print("hidden", prec, B, Din, Dout, float(h1.float().abs().mean()), flush=True)
