"""The tcgen05/TMA TF32 GEMM kernel in isolation (through the C-ABI test hook lf_debug_tc_gemm), in the
three operand orientations the wide-head path uses, against an fp64 matmul.  Tolerance: TF32 class (2e-2
norm-wise is the contract; these shapes come out near 1e-3)."""
import pytest
import torch

from tests.util import relerr, TOL_TENSOR

pytestmark = pytest.mark.gpu


def run(A, B, bias, M, N, K, a_mn, b_mn, block_n, splits=1, x3=False):
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    out = torch.full((splits, M, N), float("nan"), device="cuda")
    fn = lib.lf_debug_tc_gemm_x3 if x3 else lib.lf_debug_tc_gemm
    rc = fn(A.data_ptr(), B.data_ptr(), bias.data_ptr() if bias is not None else None,
                              out.data_ptr(), M, N, K, A.stride(0), B.stride(0), N, a_mn, b_mn, block_n, splits,
                              M * N, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "lf_debug_tc_gemm")
    torch.cuda.synchronize()
    return out.sum(0)


# (M, N, K, block_n): logits-like  Z = F W^T + b   (A K-major, B K-major)
@pytest.mark.parametrize("M,N,K,bn", [(128, 16, 32, 16), (256, 112, 768, 112), (1000, 101, 768, 112),
                                      (300, 309, 512, 160), (128, 64, 40, 64), (4096, 101, 768, 112)])
def test_tc_gemm_logits_orientation(M, N, K, bn):
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    got = run(A, W, bias, M, N, K, 0, 0, bn)
    ref = A.double() @ W.double().T + bias.double()
    assert relerr(got, ref) < 5e-3, relerr(got, ref)


# dfeat-like  dF = dZ W  (A K-major with padded pitch, B MN-major)
@pytest.mark.parametrize("M,N,K,bn", [(128, 32, 8, 32), (256, 768, 101, 256), (777, 512, 309, 256), (130, 96, 6, 32)])
def test_tc_gemm_dfeat_orientation(M, N, K, bn):
    torch.manual_seed(M + K)
    ldz = (K + 3) // 4 * 4
    dZ = torch.randn(M, ldz, device="cuda")
    W = torch.randn(K, N, device="cuda")
    got = run(dZ[:, :K], W, None, M, N, K, 0, 1, bn)
    ref = dZ[:, :K].double() @ W.double()
    assert relerr(got, ref) < 5e-3, relerr(got, ref)


# dweight-like  dW = dZ^T F  (both MN-major, split-K)
@pytest.mark.parametrize("M,N,K,bn,splits", [(32, 32, 32, 32, 1), (101, 768, 2048, 256, 1), (101, 768, 5000, 256, 7),
                                             (309, 512, 1111, 256, 3), (6, 512, 700, 256, 2)])
def test_tc_gemm_dweight_orientation(M, N, K, bn, splits):
    torch.manual_seed(M + K)
    ldz = (M + 3) // 4 * 4
    dZ = torch.randn(K, ldz, device="cuda")
    F = torch.randn(K, N, device="cuda")
    got = run(dZ[:, :M], F, None, M, N, K, 1, 1, bn, splits)
    ref = dZ[:, :M].double().T @ F.double()
    assert relerr(got, ref) < 5e-3, relerr(got, ref)


# ---------------------------------------------------------------- bf16 operands (kind::f16), fp32 accumulate
def run16(A, B, M, N, K, a_mn, b_mn, block_n, splits=1, out_bf16=False):
    """A: (M,K) values, B: (K,N) values, both already rounded to bf16; stored in the orientation under test."""
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.contiguous() if b_mn else B.t().contiguous()
    out = torch.full((splits, M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    rc = lib.lf_debug_tc_gemm16(As.data_ptr(), Bs.data_ptr(), out.data_ptr(), M, N, K, As.stride(0), Bs.stride(0), N, a_mn, b_mn,
                                block_n, splits, M * N, int(out_bf16), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "lf_debug_tc_gemm16")
    torch.cuda.synchronize()
    return out.double().sum(0)


@pytest.mark.parametrize("a_mn,b_mn,M,N,K,bn,splits,out16", [
    (0, 0, 256, 112, 768, 112, 1, False),        # logits: both K-major
    (0, 0, 1000, 304, 512, 160, 1, False),
    (0, 1, 256, 768, 104, 256, 1, True),         # dfeat: A K-major (K = padded classes), B MN-major, bf16 out
    (0, 1, 777, 512, 312, 256, 1, True),
    (0, 1, 130, 128, 8, 64, 1, False),
    (1, 1, 104, 768, 4096, 256, 4, False),       # dweight: both MN-major, split-K partials
    (1, 1, 312, 512, 700, 256, 2, False),
    (1, 1, 8, 64, 300, 64, 1, False),
])
def test_tc_gemm_bf16_orientations(a_mn, b_mn, M, N, K, bn, splits, out16):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(K, N, device="cuda") / K ** 0.5).bfloat16()
    got = run16(A, B, M, N, K, a_mn, b_mn, bn, splits, out16)
    ref = A.double() @ B.double()
    tol = 6e-3 if out16 else 1e-5                # bf16 output rounding vs exact fp32 accumulation of bf16 products
    assert relerr(got, ref) < tol, relerr(got, ref)


# ---------------------------------------------------------------------------------- 3xTF32 (exact-fp32 tier, ABI v11)
# hi / lo tf32 halves formed in shared memory by the converter warps, three MMAs per k-step.  Contract: 1e-5 norm-wise
# (BASELINE.json north_star, fp32 path); the split leaves ~2^-23 per product, so these come out near fp32 rounding.
TOL_X3 = 5e-6


@pytest.mark.parametrize("M,N,K,bn", [(128, 16, 32, 16), (256, 112, 768, 112), (1000, 101, 768, 112),
                                      (300, 309, 512, 160), (128, 64, 40, 64), (4096, 101, 768, 112), (20000, 101, 768, 112)])
def test_tc_gemm_x3_logits_orientation(M, N, K, bn):
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    got = run(A, W, bias, M, N, K, 0, 0, bn, x3=True)
    ref = A.double() @ W.double().T + bias.double()
    assert relerr(got, ref) < TOL_X3, relerr(got, ref)
    # and it must beat plain fp32 FMA-grade error bounds by construction, not by luck: a single-pass tf32 result is ~1e-3
    assert relerr(run(A, W, bias, M, N, K, 0, 0, bn), ref) > 20 * relerr(got, ref)


@pytest.mark.parametrize("M,N,K,bn", [(128, 32, 8, 32), (256, 768, 101, 256), (777, 512, 309, 256), (130, 96, 6, 32), (9000, 768, 101, 256)])
def test_tc_gemm_x3_dfeat_orientation(M, N, K, bn):
    torch.manual_seed(M + K)
    ldz = (K + 3) // 4 * 4
    dZ = torch.randn(M, ldz, device="cuda")
    W = torch.randn(K, N, device="cuda")
    got = run(dZ[:, :K], W, None, M, N, K, 0, 1, bn, x3=True)
    ref = dZ[:, :K].double() @ W.double()
    assert relerr(got, ref) < TOL_X3, relerr(got, ref)


@pytest.mark.parametrize("M,N,K,bn,splits", [(32, 32, 32, 32, 1), (101, 768, 2048, 256, 1), (101, 768, 5000, 256, 7),
                                             (309, 512, 1111, 256, 3), (6, 512, 700, 256, 2), (101, 768, 32768, 256, 24),
                                             (309, 512, 131072, 256, 12)])
def test_tc_gemm_x3_dweight_orientation(M, N, K, bn, splits):
    torch.manual_seed(M + K)
    ldz = (M + 3) // 4 * 4
    dZ = torch.randn(K, ldz, device="cuda")
    F = torch.randn(K, N, device="cuda")
    got = run(dZ[:, :M], F, None, M, N, K, 1, 1, bn, splits, x3=True)
    ref = dZ[:, :M].double().T @ F.double()
    assert relerr(got, ref) < TOL_X3, relerr(got, ref)


def test_tc_gemm_x3_poisons_exactly_the_rows_with_non_finite_inputs():
    """An Inf or NaN feature poisons the outputs an fp32 product would poison, and no others.  (Inf comes out as NaN, not
    Inf: Inf * lo(w) with lo(w) of either sign or zero is part of the sum.  Documented in DESIGN.md.)"""
    M, N, K = 128, 64, 64
    torch.manual_seed(3)
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda").abs() + 0.1
    A[5, 7] = float("inf"); A[9, 1] = float("nan")
    got = run(A, W, None, M, N, K, 0, 0, 64, x3=True)
    assert (~torch.isfinite(got[5])).all()
    assert torch.isnan(got[9]).all()
    rest = torch.ones(M, dtype=torch.bool, device="cuda"); rest[5] = rest[9] = False
    assert torch.isfinite(got[rest]).all()
