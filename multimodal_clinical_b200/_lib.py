"""ctypes binding of the C-ABI library ``_lf_fusion.so`` (declared in ``include/lf_fusion.h``).

There is deliberately NO fallback: if the shared object is missing or a call fails, an exception is
raised.  The product path never routes through PyTorch eager ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lf_fusion.so")

LF_MODE_JLOGITS, LF_MODE_QMF = 0, 1
LF_PREC_FP32, LF_PREC_TF32, LF_PREC_BF16 = 0, 1, 2
LF_LOSS_NO_JOINT, LF_LOSS_NO_UNI, LF_LOSS_NO_REG = 1, 2, 4          # LfHeadsArgs.loss_terms / LfMidArgs.loss_terms bits (QMF loss ablations)
LF_MOD_OGM_GE, LF_MOD_OGM, LF_MOD_NOISE = 0, 1, 2
LF_STATS_HEADER = 16
LF_MAX_TENSORS = 64
STAT = dict(CE_JOINT=0, CE_X1=1, CE_X2=2, SCORE_X1=3, SCORE_X2=4, CNT_X1=5, CNT_X2=6, CNT_JOINT=7,
            CNT_DF=8, CNT_X1_CAL=9, CNT_X2_CAL=10, REG_SUM=11)

_f32p = C.c_void_p   # device pointers travel as plain integers
_P2 = C.c_void_p * 2


class LfSgdFused(C.Structure):
    _fields_ = [("hyper", C.c_void_p), ("momentum_buf", C.c_void_p * 4), ("weight_bf16_out", _P2)]


LF_MAX_RANKS = 8
LF_PEER_FLAGS_BYTES = 2 * LF_MAX_RANKS * 8
_P8 = C.c_void_p * LF_MAX_RANKS


class LfPeerComm(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("flags", _P8), ("recv_payload", _P8), ("recv_grad", _P8),
                ("epoch", C.c_void_p), ("error", C.c_void_p)]


class LfHeadsArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("batch_global", C.c_int32), ("dim", C.c_int32), ("classes", C.c_int32),
        ("mode", C.c_int32), ("precision", C.c_int32), ("need_dfeat", C.c_int32), ("ld_dlogits", C.c_int32),
        ("feat", _P2), ("weight", _P2), ("bias", _P2), ("label", C.c_void_p),
        ("logits", _P2), ("avg_logits", C.c_void_p), ("logits_df", C.c_void_p), ("conf", C.c_void_p),
        ("dlogits", _P2), ("dfeat", _P2), ("dweight", _P2), ("dbias", _P2),
        ("qmf_g", C.c_void_p), ("ema_offset", C.c_void_p), ("stats", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("fwd_only", C.c_int32), ("bwd_phase", C.c_int32), ("ld_logits", C.c_int32), ("ld_fused", C.c_int32), ("loss_terms", C.c_int32), ("reserved3", C.c_int32),
        ("weight_bf16", _P2), ("sgd", C.POINTER(LfSgdFused)),
        ("grad_comm", C.POINTER(LfPeerComm)), ("reg_partial", C.c_void_p), ("loss_out", C.c_void_p),
        ("stats_rows_out", C.POINTER(C.c_uint64)),
    ]


class LfQmfArgs(C.Structure):
    _fields_ = [
        ("batch_global", C.c_int32), ("n_data", C.c_int32),
        ("idx", C.c_void_p), ("conf", C.c_void_p), ("correctness", C.c_void_p), ("confidence", C.c_void_p),
        ("last_writer", C.c_void_p), ("step_base", C.c_int64), ("stats", C.c_void_p), ("qmf_g", C.c_void_p),
        ("target_out", C.c_void_p), ("g_begin", C.c_int32), ("g_count", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("flags", C.c_int32), ("reserved", C.c_int32), ("loss_uni", _P2),
    ]


class LfPeerReduceArgs(C.Structure):
    _fields_ = [("comm", LfPeerComm), ("buf", C.c_void_p), ("n", C.c_int32), ("n_padded", C.c_int32),
                ("tail_dst", C.c_void_p), ("tail_n", C.c_int32), ("reserved", C.c_int32)]


class LfMidArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("classes", C.c_int32), ("batch_global", C.c_int32), ("n_ranks", C.c_int32),
        ("batch_local", C.c_int32), ("rank", C.c_int32), ("n_data", C.c_int32), ("update_ema", C.c_int32),
        ("stats_parts", C.c_void_p), ("stats_stride", C.c_int64), ("idx_parts", C.c_void_p), ("idx_stride", C.c_int64),
        ("conf_parts", C.c_void_p), ("conf_stride", C.c_int64), ("stats", C.c_void_p), ("ema_x", C.c_void_p),
        ("ema_offset", C.c_void_p), ("smoothing", C.c_float), ("alpha", C.c_float), ("coeff_out", C.c_void_p),
        ("correctness", C.c_void_p), ("confidence", C.c_void_p), ("last_writer", C.c_void_p), ("step_base", C.c_int64),
        ("qmf_g", C.c_void_p), ("loss_out", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("use_peer", C.c_int32), ("loss_terms", C.c_int32), ("payload_local", C.c_void_p), ("payload_bytes", C.c_int64),
        ("off_idx", C.c_int64), ("off_conf", C.c_int64), ("comm", LfPeerComm),
        ("reg_partial_out", C.c_void_p), ("payload_idx_src", C.c_void_p),
        ("stats_rows", C.c_void_p), ("n_stats_rows", C.c_int64),
    ]


class LfSgdArgs(C.Structure):
    _fields_ = [("count", C.c_int32), ("first_step", C.c_int32), ("lr", C.c_float), ("momentum", C.c_float),
                ("weight_decay", C.c_float), ("reserved", C.c_int32), ("param", C.c_void_p * 8), ("grad", C.c_void_p * 8),
                ("momentum_buf", C.c_void_p * 8), ("numel", C.c_int64 * 8)]


LF_QMF_UPDATE_X1, LF_QMF_UPDATE_X2, LF_QMF_REG, LF_QMF_ALL = 1, 2, 4, 7
LF_MAX_MODALITIES = 4
_P4 = C.c_void_p * LF_MAX_MODALITIES


class LfHiddenArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("dim_in", C.c_int32), ("dim_out", C.c_int32), ("precision", C.c_int32),
                ("training", C.c_int32), ("drop_p", C.c_float), ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("x", _P2), ("weight", _P2), ("bias", _P2), ("h", _P2), ("dh", _P2), ("dpre", _P2), ("dx", _P2),
                ("dweight", _P2), ("dbias", _P2), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class LfMultiHeadsArgs(C.Structure):
    _fields_ = [("modalities", C.c_int32), ("batch", C.c_int32), ("classes", C.c_int32), ("need_dfeat", C.c_int32),
                ("dim", C.c_int32 * LF_MAX_MODALITIES), ("feat", _P4), ("weight", _P4), ("bias", _P4), ("label", C.c_void_p),
                ("logits", _P4), ("avg_logits", C.c_void_p), ("dweight", _P4), ("dbias", _P4), ("dfeat", _P4),
                ("loss_out", C.c_void_p), ("stats", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class LfTensorList(C.Structure):
    _fields_ = [("count", C.c_int32), ("reserved", C.c_int32),
                ("data", C.c_void_p * LF_MAX_TENSORS), ("numel", C.c_int64 * LF_MAX_TENSORS)]


# name -> (restype, argtypes); also the list the "exports every declared symbol" test walks
SIGNATURES = {
    "lf_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "lf_heads_forward": (C.c_int, [C.POINTER(LfHeadsArgs), C.c_void_p]),
    "lf_heads_backward": (C.c_int, [C.POINTER(LfHeadsArgs), C.c_void_p]),
    "lf_heads_backward_fuses_allreduce": (C.c_int, [C.POINTER(LfHeadsArgs)]),
    "lf_heads_backward_splits_rows": (C.c_int, [C.POINTER(LfHeadsArgs)]),
    "lf_grad_exchange_floats": (C.c_size_t, [C.c_int32, C.c_int32]),
    "lf_cast_heads_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lf_loss_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lf_ema_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "lf_ogm_coeff": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "lf_qmf_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "lf_qmf_history_step": (C.c_int, [C.POINTER(LfQmfArgs), C.c_void_p]),
    "lf_comm_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "lf_comm_fill": (C.c_int, [C.c_void_p, C.c_int32, C.c_size_t]),
    "lf_comm_free": (C.c_int, [C.c_void_p]),
    "lf_comm_ipc_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lf_comm_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "lf_comm_ipc_close": (C.c_int, [C.c_void_p]),
    "lf_peer_allreduce": (C.c_int, [C.POINTER(LfPeerReduceArgs), C.c_void_p]),
    "lf_mid_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "lf_pool_mean": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "lf_pool_mean_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "lf_epoch_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "lf_epoch_offset_correction": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lf_step_mid": (C.c_int, [C.POINTER(LfMidArgs), C.c_void_p]),
    "lf_modulate_workspace_bytes": (C.c_size_t, []),
    "lf_ogm_modulate": (C.c_int, [C.POINTER(LfTensorList), C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "lf_qmf_df": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lf_ogm_scores_workspace_bytes": (C.c_size_t, []),
    "lf_ogm_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "lf_sgd_heads": (C.c_int, [C.POINTER(LfSgdArgs), C.c_void_p]),
    "lf_hidden_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "lf_hidden_forward": (C.c_int, [C.POINTER(LfHiddenArgs), C.c_void_p]),
    "lf_hidden_backward": (C.c_int, [C.POINTER(LfHiddenArgs), C.c_void_p]),
    "lf_multi_heads_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "lf_multi_heads_step": (C.c_int, [C.POINTER(LfMultiHeadsArgs), C.c_void_p]),
    "lf_launch_count": (C.c_int64, []),
    "lf_profile_enable": (None, [C.c_int32]),
    "lf_profile_report": (C.c_int32, [C.c_char_p, C.c_int32]),
    "lf_debug_tc_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 10 +
                         [C.c_int64, C.c_void_p]),
    "lf_debug_tc_gemm_x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 10 +
                         [C.c_int64, C.c_void_p]),
    "lf_debug_tc_gemm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 10 + [C.c_int64, C.c_int32, C.c_void_p]),
    "lf_last_error": (C.c_char_p, []),
    "lf_abi_version": (C.c_int32, []),
}

_lib = None


class LfError(RuntimeError):
    pass


def load():
    """Load the library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LfError(f"{LIB_PATH} is missing: build it with `python -m multimodal_clinical_b200.build` "
                      "(or __graft_entry__.build()). There is no CPU/eager fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().lf_last_error()
        raise LfError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def profile_report() -> dict:
    """{kernel name: (launches, total_ms)} for everything timed since the last report."""
    lib = load()
    buf = C.create_string_buffer(1 << 16)
    lib.lf_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
