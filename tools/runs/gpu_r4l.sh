#!/bin/bash
# final state: whole GPU suite, smoke, the driver's default bench line; then two side measurements (ensemble loss on wide
# heads through mid_light_kernel with ~150-300 per-CTA statistic rows; step_mid without the cooperative launch attribute)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r4l_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4l_smoke.log 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/r4l_k4.json 2> gpurun_out/r4l_k4.err
timeout 120 python - > gpurun_out/r4l_ensemble.log 2>&1 <<'PY'
import torch
from multimodal_clinical_b200.step import LateFusionStep
from multimodal_clinical_b200 import _lib
lib = _lib.load()
B, D, C = 32768, 768, 101
g = torch.Generator().manual_seed(1)
f = [torch.randn(B, D, generator=g).cuda().bfloat16() for _ in range(2)]
W = [(torch.randn(C, D, generator=g) * 0.03).cuda() for _ in range(2)]; b = [torch.zeros(C).cuda() for _ in range(2)]
y = torch.randint(0, C, (B,), generator=g).cuda()
for mode in ("ensemble", "jlogits"):
    e = LateFusionStep(C, mode=mode, device="cuda:0", precision="bf16")
    for _ in range(3):
        e.step(f, W, b, y, ogm_alpha=0.5)
    lib.lf_profile_enable(1)
    for _ in range(10):
        e.step(f, W, b, y, ogm_alpha=0.5)
    rep = _lib.profile_report(); lib.lf_profile_enable(0)
    print(mode, {k: round(v[1] / v[0] * 1e3, 2) for k, v in rep.items()}, flush=True)
PY
LF_MID_NOCOOP=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r4l_k4_nocoop.json 2> gpurun_out/r4l_k4_nocoop.err
echo done
