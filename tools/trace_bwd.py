"""LF_BWD_TRACE=1 python tools/trace_bwd.py : per-role %globaltimer stamps of the fused K4 backward (eager launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_clinical_b200.step import LateFusionStep

B, D, C, N = 32768, 768, 101, 65536
eng = LateFusionStep(C, mode="qmf", n_data=N, device="cuda:0", precision="bf16")
g = torch.Generator().manual_seed(0)
W = [torch.randn(C, D, generator=g).cuda() * 0.03 for _ in range(2)]
b = [torch.zeros(C).cuda() for _ in range(2)]
sets = []
for i in range(6):
    sets.append(([torch.randn(B, D, generator=g).cuda().bfloat16() for _ in range(2)], torch.randint(0, C, (B,), generator=g).cuda(),
                 ((torch.arange(B) + 1000 * i) % N).cuda()))
for i in range(14):
    f, y, idx = sets[i % 6]
    eng.step(f, W, b, y, idx=idx)
torch.cuda.synchronize()
