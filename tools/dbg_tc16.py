"""bf16 tcgen05 GEMM bring-up: K-major sanity, then a sweep of the MN-major UMMA descriptor / TMA swizzle parameters
(one subprocess per configuration: a wrong descriptor can fault the context)."""
import itertools, os, subprocess, sys
sys.path.insert(0, os.getcwd())

def run_case(a_mn, b_mn, M, N, K, bn, splits=1, out_bf16=0):
    import torch
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda") / K ** 0.5
    A16, B16 = A.bfloat16(), B.bfloat16()
    ref = A16.double() @ B16.double()
    As = A16.t().contiguous() if a_mn else A16.contiguous()                 # stored (K,M) when MN-major
    Bs = B16.contiguous() if b_mn else B16.t().contiguous()                 # stored (K,N) when MN-major, else (N,K)
    out = torch.full((splits, M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    rc = lib.lf_debug_tc_gemm16(As.data_ptr(), Bs.data_ptr(), out.data_ptr(), M, N, K, As.stride(0), Bs.stride(0), N, a_mn, b_mn, bn,
                                splits, M * N, out_bf16, torch.cuda.current_stream().cuda_stream)
    if rc:
        return "rc=%d %s" % (rc, lib.lf_last_error().decode())
    torch.cuda.synchronize()
    got = out.double().sum(0)
    err = ((got - ref).norm() / ref.norm()).item()
    return "err %.4f" % err

if len(sys.argv) > 1:
    a_mn, b_mn, M, N, K, bn = [int(x) for x in sys.argv[1:7]]
    print(sys.argv[7] if len(sys.argv) > 7 else "", run_case(a_mn, b_mn, M, N, K, bn), flush=True)
else:
    def sub(args, env=None, tag=""):
        r = subprocess.run([sys.executable, __file__] + [str(a) for a in args] + [tag], env=dict(os.environ, **(env or {})),
                           capture_output=True, text=True, timeout=120)
        print((r.stdout.strip() or ("FAIL " + tag + " " + r.stderr.strip()[-200:])), flush=True)
    print("== K-major A and B (logits orientation)")
    sub([0, 0, 256, 112, 768, 112], tag="kmajor 256x112x768")
    sub([0, 0, 1000, 304, 512, 160], tag="kmajor 1000x304x512")
    print("== B MN-major sweep (dfeat orientation), M=128 N=64 K=16 then K=64")
    for lt, swz, sbo, lbo, step in itertools.product([2, 1], [3, 4], [1024, 512, 2048], [8192, 1024], [2048, 1024]):
        env = dict(LF_TC16_LT=str(lt), LF_TC16_SWZ=str(swz), LF_TC16_SBO=str(sbo), LF_TC16_LBO=str(lbo), LF_TC16_STEP=str(step))
        sub([0, 1, 128, 128, 64, 128], env, tag=f"lt={lt} swz={swz} sbo={sbo} lbo={lbo} step={step}")
