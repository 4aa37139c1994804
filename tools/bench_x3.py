"""Timing of the 3xTF32 GEMM (lf_debug_tc_gemm_x3) on the three K4 head products, CUDA events, inputs > L2 rotated."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_clinical_b200 import _lib
lib = _lib.load()
B, D, C = 32768, 768, 101
ldz = 104
def timeit(fn, n=10):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
st = torch.cuda.current_stream().cuda_stream
Fs = [torch.randn(B, D, device="cuda") for _ in range(4)]
W = torch.randn(C, D, device="cuda"); bias = torch.randn(C, device="cuda")
Z = torch.empty(B, ldz, device="cuda"); dZ = torch.randn(B, ldz, device="cuda")
dF = torch.empty(B, D, device="cuda"); part = torch.empty(64, C, D, device="cuda")
x3 = int(os.environ.get("X3", "1"))
fn = lib.lf_debug_tc_gemm_x3 if x3 else lib.lf_debug_tc_gemm
def logits(i): _lib.check(fn(Fs[i % 4].data_ptr(), W.data_ptr(), bias.data_ptr(), Z.data_ptr(), B, C, D, D, D, ldz, 0, 0, 112, 1, 0, st), "l")
def dfeat(i): _lib.check(fn(dZ.data_ptr(), W.data_ptr(), None, dF.data_ptr(), B, D, C, ldz, D, D, 0, 1, 256, 1, 0, st), "df")
def dweight(i): _lib.check(fn(dZ.data_ptr(), Fs[i % 4].data_ptr(), None, part.data_ptr(), C, D, B, ldz, D, D, 1, 1, 256, 48, C * D, st), "dw")
print(f"X3={x3} DBG={os.environ.get('LF_X3_DBG','0')}: logits {timeit(logits):.1f} us  dfeat {timeit(dfeat):.1f} us  dweight(48 splits) {timeit(dweight):.1f} us (one modality each)")
