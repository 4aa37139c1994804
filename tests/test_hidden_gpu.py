"""The Food101 MLPs' hidden layers on lf_hidden_forward / lf_hidden_backward (csrc/lf_hidden.cu):
``dropout(relu(linear(x)))`` for both modalities per call, against torch in fp64 on the same inputs (bf16-rounded where
the mode rounds).  Eval mode is compared exactly; in training mode the dropout mask is our own Philox stream, so the mask is
read back from the output (kept and active <=> h > 0) and torch is run with THAT mask."""
import pytest
import torch
import torch.nn as nn

from tests.util import assert_close, TOL_FP32, TOL_TENSOR

pytestmark = pytest.mark.gpu

TOL = {"fp32": TOL_FP32, "tf32": TOL_TENSOR, "bf16": TOL_TENSOR}


def _layers(Din, Dout, seed):
    torch.manual_seed(seed)
    return nn.Linear(Din, Dout).cuda(), nn.Linear(Din, Dout).cuda()


def _ref(x, lin, mask, scale, gout, rnd, h_ours):
    """torch in fp64.  The backward takes the ReLU's on/off pattern from OUR forward (h > 0): a unit whose pre-activation is
    within rounding of zero may land on either side in fp32 vs fp64, and ONE flipped unit out of B * Dout moves the norm of
    dx by ~1e-3 -- a property of comparing any fp32 ReLU with an fp64 one, not of the kernel (its value in h is ~1e-7)."""
    x = rnd(x.detach()).double().requires_grad_(True)
    W = rnd(lin.weight.detach()).double().requires_grad_(True)
    b = lin.bias.detach().double().requires_grad_(True)
    pre = x @ W.T + b
    h_fwd = torch.relu(pre.detach()) * mask * scale
    on = (h_ours.detach() > 0).double() if isinstance(mask, float) else mask
    ((pre * on * scale) * gout.double()).sum().backward()
    return h_fwd, x.grad, W.grad, b.grad


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("B,Din,Dout", [(1000, 768, 512), (37, 64, 128), (4096, 512, 512), (130, 72, 40)])
def test_hidden_pair_eval_mode_matches_torch(prec, B, Din, Dout):
    from multimodal_clinical_b200.hidden import FusedHiddenPair
    l1, l2 = _layers(Din, Dout, B + Din)
    fused = FusedHiddenPair(precision=prec).eval()
    g = torch.Generator(device="cuda").manual_seed(B)
    xs = [torch.randn(B, Din, device="cuda", generator=g).requires_grad_(True) for _ in range(2)]
    gouts = [torch.randn(B, Dout, device="cuda", generator=g) for _ in range(2)]
    rnd = (lambda t: t.bfloat16().float()) if prec == "bf16" else (lambda t: t)
    h1, h2 = fused(xs[0], xs[1], l1, l2)
    assert h1.dtype == (torch.bfloat16 if prec == "bf16" else torch.float32)
    (h1.float() * rnd(gouts[0]) + 0).sum().backward(retain_graph=True)
    (h2.float() * rnd(gouts[1])).sum().backward()
    torch.cuda.synchronize()
    for x, lin, h, go in ((xs[0], l1, h1, gouts[0]), (xs[1], l2, h2, gouts[1])):
        rh, rdx, rdW, rdb = _ref(x, lin, 1.0, 1.0, rnd(go), rnd, h)
        assert_close(h.float(), rh, TOL[prec], "h")
        assert_close(x.grad, rdx, TOL[prec], "dx")
        assert_close(lin.weight.grad, rdW, TOL[prec], "dW")
        assert_close(lin.bias.grad, rdb, TOL[prec], "db")


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_hidden_pair_training_mode_dropout(prec):
    from multimodal_clinical_b200.hidden import FusedHiddenPair
    B, Din, Dout, p = 3000, 256, 512, 0.2
    l1, l2 = _layers(Din, Dout, 7)
    fused = FusedHiddenPair(drop_p=p, precision=prec).train()
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(B, Din, device="cuda", generator=g).requires_grad_(True) for _ in range(2)]
    gouts = [torch.randn(B, Dout, device="cuda", generator=g) for _ in range(2)]
    rnd = (lambda t: t.bfloat16().float()) if prec == "bf16" else (lambda t: t)
    torch.manual_seed(123)
    h1, h2 = fused(xs[0], xs[1], l1, l2)
    (h1.float() * rnd(gouts[0])).sum().backward(retain_graph=True)
    (h2.float() * rnd(gouts[1])).sum().backward()
    torch.cuda.synchronize()
    for x, lin, h, go in ((xs[0], l1, h1, gouts[0]), (xs[1], l2, h2, gouts[1])):
        pre = torch.relu(rnd(x.detach()).double() @ rnd(lin.weight.detach()).double().T + lin.bias.detach().double())
        active = pre > 1e-2                                  # clearly active units: h == 0 there means "dropped"
        kept = (h.float() > 0)
        rate = float((kept & active).sum()) / float(active.sum())
        assert abs(rate - (1 - p)) < 0.01, rate               # ~7e5 Bernoulli draws: sigma ~ 5e-4
        mask = torch.where(pre > 0, kept.double(), torch.zeros_like(pre))
        rh, rdx, rdW, rdb = _ref(x, lin, mask, 1.0 / (1 - p), rnd(go), rnd, h)
        assert_close(h.float(), rh, TOL[prec], "h"); assert_close(x.grad, rdx, TOL[prec], "dx")
        assert_close(lin.weight.grad, rdW, TOL[prec], "dW"); assert_close(lin.bias.grad, rdb, TOL[prec], "db")
    # the two modalities draw different masks; a new call draws a new mask; the same seed and call index reproduce it
    assert (h1 > 0).ne(h2 > 0).float().mean() > 0.1
    again1, _ = fused(xs[0], xs[1], l1, l2)
    assert (again1 > 0).ne(h1 > 0).float().mean() > 0.1
    fused2 = FusedHiddenPair(drop_p=p, precision=prec).train()
    fused2.layer_id, fused2.calls = fused.layer_id, 0
    torch.manual_seed(123)
    rep1, _ = fused2(xs[0], xs[1], l1, l2)
    assert torch.equal(rep1, h1)
    # the mask of a row does not depend on the batch it is part of (it is a function of (row, column) only)
    fused2.calls = 0
    part1, _ = fused2(xs[0][:700].detach(), xs[1][:700].detach(), l1, l2)
    assert torch.equal(part1 > 0, h1[:700] > 0)


def test_k4_sized_hidden_layers_bf16_and_kernel_names():
    """The Food101 shape the bench is quoted on (B = 32768, 768 -> 512 -> 512) through both fused layers in bf16."""
    from multimodal_clinical_b200 import _lib
    from multimodal_clinical_b200.food101._common import MLP, FusedMLPHidden
    torch.manual_seed(5)
    m1, m2 = MLP(768, 512, 101).cuda().eval(), MLP(768, 512, 101).cuda().eval()
    hid = FusedMLPHidden(precision="bf16")
    e1 = torch.randn(32768, 768, device="cuda").bfloat16().requires_grad_(True)
    e2 = torch.randn(32768, 768, device="cuda").bfloat16().requires_grad_(True)
    lib = _lib.load(); lib.lf_profile_enable(1)
    h1, h2 = hid(m1, m2, e1, e2)
    (h1.float().sum() + 2 * h2.float().sum()).backward()
    torch.cuda.synchronize()
    prof = _lib.profile_report(); lib.lf_profile_enable(0)
    assert {"hidden_forward", "hidden_dpre", "hidden_dx", "hidden_dw"} <= set(prof), prof
    with torch.autocast("cuda", dtype=torch.bfloat16):
        r1 = m1.hidden(e1.detach()); r2 = m2.hidden(e2.detach())
    assert_close(h1.float(), r1.float(), TOL_TENSOR, "h1"); assert_close(h2.float(), r2.float(), TOL_TENSOR, "h2")
    ea = e1.detach().clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        m1.zero_grad(); m1.hidden(ea).float().sum().backward()
    # m1's parameters now hold torch's gradients; ours were accumulated before zero_grad -> recompute ours
    tg = [p.grad.clone() for p in m1.mlp[:6].parameters()]
    m1.zero_grad(); m2.zero_grad(); e1.grad = None
    h1, h2 = hid(m1, m2, e1, e2)
    (h1.float().sum() + 2 * h2.float().sum()).backward()
    for a, b in zip([p.grad for p in m1.mlp[:6].parameters()], tg):
        assert_close(a, b, TOL_TENSOR, "mlp gradient")
    assert_close(e1.grad.float(), ea.grad.float(), TOL_TENSOR, "d embeddings")
