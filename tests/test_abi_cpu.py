"""CPU-side checks of the C-ABI boundary: the library builds and loads, exports every symbol the header
declares (and nothing the ctypes table does not know), the POD structs agree in size with the C side,
argument validation fails loudly, and nothing in the product package reaches for the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lf_fusion.h")


@pytest.fixture(scope="module")
def lib():
    from multimodal_clinical_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from multimodal_clinical_b200 import _lib
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lf_fusion.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.lf_abi_version() == 12


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof computed by gcc from the header must equal the ctypes mirror."""
    from multimodal_clinical_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lf_fusion.h"\nint main(){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(LfSgdFused), offsetof(LfHeadsArgs, sgd), sizeof(LfHeadsArgs), offsetof(LfHeadsArgs, feat),'
                    'offsetof(LfHeadsArgs, stats), sizeof(LfQmfArgs), offsetof(LfQmfArgs, step_base),'
                    'offsetof(LfQmfArgs, workspace), sizeof(LfTensorList), sizeof(LfMidArgs), offsetof(LfMidArgs, stats),'
                    'offsetof(LfMidArgs, step_base), offsetof(LfHeadsArgs, fwd_only)); return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.LfSgdFused), _lib.LfHeadsArgs.sgd.offset, C.sizeof(_lib.LfHeadsArgs), _lib.LfHeadsArgs.feat.offset, _lib.LfHeadsArgs.stats.offset,
            C.sizeof(_lib.LfQmfArgs), _lib.LfQmfArgs.step_base.offset, _lib.LfQmfArgs.workspace.offset,
            C.sizeof(_lib.LfTensorList), C.sizeof(_lib.LfMidArgs), _lib.LfMidArgs.stats.offset,
            _lib.LfMidArgs.step_base.offset, _lib.LfHeadsArgs.fwd_only.offset]
    assert got == want
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lf_fusion.h"\nint main(){'
                    'printf("%zu %zu %zu %zu %zu\\n", sizeof(LfMultiHeadsArgs), offsetof(LfMultiHeadsArgs, feat), offsetof(LfMultiHeadsArgs, label),'
                    'offsetof(LfMultiHeadsArgs, dfeat), offsetof(LfMultiHeadsArgs, workspace_bytes)); return 0;}')
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    M = _lib.LfMultiHeadsArgs
    assert got == [C.sizeof(M), M.feat.offset, M.label.offset, M.dfeat.offset, M.workspace_bytes.offset]
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lf_fusion.h"\nint main(){'
                    'printf("%zu %zu %zu %zu %zu\\n", sizeof(LfHiddenArgs), offsetof(LfHiddenArgs, seed), offsetof(LfHiddenArgs, x),'
                    'offsetof(LfHiddenArgs, dx), offsetof(LfHiddenArgs, workspace_bytes)); return 0;}')
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    H = _lib.LfHiddenArgs
    assert got == [C.sizeof(H), H.seed.offset, H.x.offset, H.dx.offset, H.workspace_bytes.offset]


def test_workspace_sizes_are_monotone(lib):
    a = lib.lf_workspace_bytes(64, 512, 6)
    b = lib.lf_workspace_bytes(8192, 512, 6)
    c = lib.lf_workspace_bytes(8192, 768, 101)
    assert 0 < a <= b < c
    assert lib.lf_workspace_bytes(0, 512, 6) == 0
    assert lib.lf_qmf_workspace_bytes(1000) > 0 and lib.lf_modulate_workspace_bytes() >= 2 * 8 * 64


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    from multimodal_clinical_b200 import _lib
    a = _lib.LfHeadsArgs()
    assert lib.lf_heads_forward(C.byref(a), None) == -1            # LF_ERR_BAD_ARG
    assert b"bad sizes" in lib.lf_last_error()
    a.batch, a.batch_global, a.dim, a.classes = 8, 8, 30, 4          # dim not a multiple of 4
    assert lib.lf_heads_forward(C.byref(a), None) == -1
    q = _lib.LfQmfArgs()
    assert lib.lf_qmf_history_step(C.byref(q), None) == -1
    tl = _lib.LfTensorList()
    tl.count = 65
    assert lib.lf_ogm_modulate(C.byref(tl), None, 0, 0, 0, None, 0, None) == -1
    tl.count = 0                                                      # no 4-D grads (Food101 MLPs): no-op, OK
    assert lib.lf_ogm_modulate(C.byref(tl), None, 2, 0, 0, None, 0, None) == 0


def test_capability_queries_are_pure_host_functions(lib):
    """lf_heads_backward_splits_rows / lf_heads_backward_fuses_allreduce decide the step's launch plan on the host (no CUDA
    call): narrow heads and the fused QMF backward form dL/dz inside one kernel (no bwd_phase 3 / 4), wide mean-fusion heads
    take the row kernel as a separate phase, and the dW tail (with the fused all-reduce) exists where the tiles fit one wave."""
    from multimodal_clinical_b200 import _lib
    a = _lib.LfHeadsArgs()
    assert lib.lf_heads_backward_splits_rows(None) == 0 and lib.lf_heads_backward_splits_rows(C.byref(a)) == 0
    a.batch, a.batch_global, a.dim, a.classes, a.need_dfeat = 256, 256, 512, 6, 1
    a.mode, a.precision = _lib.LF_MODE_JLOGITS, _lib.LF_PREC_FP32
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 0          # K2 / K3: one fused FMA pass
    assert lib.lf_heads_backward_fuses_allreduce(C.byref(a)) == 0
    a.classes, a.dim, a.mode, a.precision = 101, 768, _lib.LF_MODE_QMF, _lib.LF_PREC_BF16
    a.batch = a.batch_global = 32768
    a.ld_logits, a.ld_dlogits = 104, 104
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 0          # K4: tc_backward_qmf forms dL/dz itself
    assert lib.lf_heads_backward_fuses_allreduce(C.byref(a)) == 1
    a.mode = _lib.LF_MODE_JLOGITS                                      # Food101 mean fusion: stand-alone row kernels
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 1
    a.classes, a.dim, a.ld_logits, a.ld_dlogits = 309, 512, 312, 312   # K5
    a.batch = a.batch_global = 131072
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 1
    assert lib.lf_heads_backward_fuses_allreduce(C.byref(a)) == 1
    a.bwd_phase = 5                                                    # phases are 0 .. 4
    a.precision = _lib.LF_PREC_TF32
    assert lib.lf_heads_backward(C.byref(a), None) != 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from multimodal_clinical_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LfError, match="no CPU/eager fallback"):
        _lib.load()


def test_step_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_clinical_b200 import _lib
    from multimodal_clinical_b200.step import LateFusionStep
    with pytest.raises(_lib.LfError):
        LateFusionStep(6, mode="jlogits")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal_clinical_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "/root/reference" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_sass_is_sm100a():
    so = os.path.join(ROOT, "multimodal_clinical_b200", "_lf_fusion.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_staged_reference_copy_is_unmodified_and_runs():
    """oracle/_ref (staged by oracle/vendor_ref.py from /root/reference, git-ignored) matches its sha256 manifest and the
    literal reference step runs on the CPU; skipped where neither the copy nor the reference tree exists."""
    import pytest
    from oracle import vendor_ref
    if not vendor_ref.verify() and not vendor_ref.stage(verbose=False):
        pytest.skip("no staged reference copy and no /root/reference")
    assert vendor_ref.verify()
    from oracle import literal
    v, n, dt = literal.time_step("qmf", 64, 32, 5, 100, None, budget_s=0.5, max_steps=2)
    assert v > 0 and n >= 1
