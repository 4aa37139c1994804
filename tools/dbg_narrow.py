import sys, torch, ctypes as C
sys.path.insert(0, ".")
from multimodal_clinical_b200 import _lib
from multimodal_clinical_b200.step import LateFusionStep
from oracle import late_fusion as O
lib = _lib.load()
orig = {}
for n in ("lf_heads_forward", "lf_step_mid", "lf_heads_backward"):
    f = getattr(lib, n)
    def wrap(*a, _f=f, _n=n):
        print("call", _n, flush=True)
        rc = _f(*a)
        torch.cuda.synchronize()
        print("done", _n, rc, flush=True)
        return rc
    setattr(lib, n, wrap)
mode = sys.argv[1]
B, D, Cn, N = 64, 512, 6, 997
inp = O.make_inputs(B, D, Cn, seed=5, n_data=N)
eng = LateFusionStep(Cn, mode=mode, n_data=N if mode == "qmf" else None, device="cuda:0")
out = eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()],
               inp["y"].cuda(), idx=inp["idx"].cuda() if mode == "qmf" else None)
torch.cuda.synchronize()
print("loss", float(out.loss))
