"""One eager pass over the kernels added in the last part of round 2, for ncu and for CUDA-event timings (lf_profile_*):
  1. K4 in the exact tier: LateFusionStep(precision="fp32"), C = 101 -> tc_logits / tc_dfeat / tc_dweight through 3xTF32
  2. multi_heads_step: avmnist-shaped heads (48 + 192 -> 10) and mustard-shaped heads (3 x 100 -> 2), B = 262144
  3. the Food101 MLP hidden layers at the K4 batch (32768 x 768 -> 512 -> 512), bf16, forward + backward
Launch order is the order of the rows in the ncu summary."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_clinical_b200 import _lib
from multimodal_clinical_b200.step import LateFusionStep
from multimodal_clinical_b200.multi import MultiHeadStep
from multimodal_clinical_b200.food101._common import MLP, FusedMLPHidden

reps = int(os.environ.get("REPS", "3"))
lib = _lib.load()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
B, D, C, N = 32768, 768, 101, 65536
eng = LateFusionStep(C, mode="qmf", n_data=N, device=dev, precision="fp32")
f = [torch.randn(B, D, device=dev, generator=g) for _ in range(2)]
W = [torch.randn(C, D, device=dev, generator=g) / D ** 0.5 for _ in range(2)]
b = [torch.zeros(C, device=dev) for _ in range(2)]
y = torch.randint(0, C, (B,), device=dev, generator=g); idx = torch.arange(B, device=dev)
Bm = 262144
av = MultiHeadStep(10, device=dev); mu = MultiHeadStep(2, device=dev)
fa = [torch.randn(Bm, d, device=dev, generator=g) for d in (48, 192)]
Wa = [torch.randn(10, d, device=dev, generator=g) / d ** 0.5 for d in (48, 192)]; ba = [torch.zeros(10, device=dev) for _ in range(2)]
ya = torch.randint(0, 10, (Bm,), device=dev, generator=g)
fm = [torch.randn(Bm, 100, device=dev, generator=g) for _ in range(3)]
Wm = [torch.randn(2, 100, device=dev, generator=g) / 10 for _ in range(3)]; bm = [torch.zeros(2, device=dev) for _ in range(3)]
ym = torch.randint(0, 2, (Bm,), device=dev, generator=g)
m1, m2 = MLP(768, 512, 101).to(dev).train(), MLP(768, 512, 101).to(dev).train()
hid = FusedMLPHidden(precision="bf16")
e1 = torch.randn(B, D, device=dev, generator=g).bfloat16().requires_grad_(True)
e2 = torch.randn(B, D, device=dev, generator=g).bfloat16().requires_grad_(True)


def one_pass():
    eng.step(f, W, b, y, idx=idx)
    av.step(fa, Wa, ba, ya)
    mu.step(fm, Wm, bm, ym)
    h1, h2 = hid(m1, m2, e1, e2)
    (h1.float().sum() + h2.float().sum()).backward()


one_pass(); torch.cuda.synchronize()
lib.lf_profile_enable(1)
for _ in range(reps):
    one_pass()
torch.cuda.synchronize()
prof = _lib.profile_report(); lib.lf_profile_enable(0)
for k, (n, ms) in prof.items():
    print(f"{k:24s} launches {n:3d}  avg {1e3 * ms / n:8.1f} us")
