"""Collective plumbing of the batch-sharded late-fusion step (one process per GPU, torch.distributed).

The reference defines no multi-GPU semantics; here the global result is DEFINED as the
single-GPU reference applied to the concatenation of all ranks' shards in rank order.  That needs:
  * all-reduce(sum) of the packed statistics (score sums for OGM-GE, CE sums for the QMF History, logit
    sums for the EMA, accuracy counts) before anything derived from a batch mean,
  * all-gather of (idx, conf) for QMF, because the History scatter, the global min/max and the
    flattened-roll neighbour (which crosses shard and modality boundaries) need the whole batch,
  * all-reduce(sum) of the head gradients.
Works on any backend (NCCL on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world(pg=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(pg), dist.get_world_size(pg)
    return 0, 1


def shard_range(rank: int, batch_local: int) -> Tuple[int, int]:
    """[begin, begin+count) of this rank's samples in the global batch order."""
    return rank * batch_local, batch_local


def allreduce_sum_(t: torch.Tensor, pg=None) -> torch.Tensor:
    if world(pg)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
    return t


def gather_payload(payload: torch.Tensor, pg=None, engine_world: Optional[int] = None) -> torch.Tensor:
    """Every rank's [stats | idx | conf] byte payload -> (world, nbytes), rank-major: the ONE exchange between
    the forward and the backward pass.  lf_step_mid consumes the rank-major layout directly.

    ``engine_world`` is the world size the CALLING ENGINE works with.  An engine that treats its batch as the whole
    batch (``LateFusionStep(sharded=False)``: the DDP layout) passes 1 and gets its own payload back without any
    collective, whatever torch.distributed says -- row 0 of a gathered buffer would be rank 0's statistics."""
    if engine_world is not None and engine_world == 1:
        return payload
    rank, ws = world(pg)
    if ws == 1:
        return payload
    out = torch.empty(ws * payload.numel(), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload, group=pg)
    return out


def gather_batch(idx_local: torch.Tensor, conf_local: torch.Tensor, pg=None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx (B,) int64 and conf (2,B) of every rank -> idx (Bg,), conf (2,Bg) in global batch order."""
    rank, ws = world(pg)
    if ws == 1:
        return idx_local, conf_local
    B = idx_local.numel()
    idx_g = torch.empty(ws * B, dtype=idx_local.dtype, device=idx_local.device)
    dist.all_gather_into_tensor(idx_g, idx_local.contiguous(), group=pg)
    cg = torch.empty(ws * 2, B, dtype=conf_local.dtype, device=conf_local.device)   # rank-major concat on dim 0
    dist.all_gather_into_tensor(cg, conf_local.contiguous().view(2, B), group=pg)
    return idx_g, cg.view(ws, 2, B).permute(1, 0, 2).reshape(2, ws * B).contiguous()


def pack_grad_exchange(grad_flat: torch.Tensor, n_head: int, stats: torch.Tensor, cal_lo: int, cal_hi: int,
                       pg=None) -> None:
    """ONE all-reduce for [dW1|db1|dW2|db2|calibrated counts]: the two fp64 counts ride in the tail of the
    fp32 gradient buffer (exact: counts < 2^24)."""
    if world(pg)[1] == 1:
        return
    grad_flat[n_head:n_head + (cal_hi - cal_lo)] = stats[cal_lo:cal_hi].to(grad_flat.dtype)
    dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=pg)
    stats[cal_lo:cal_hi] = grad_flat[n_head:n_head + (cal_hi - cal_lo)].to(stats.dtype)


class PeerComm:
    """Peer-memory communicator: one symmetric buffer per rank (flags + double-buffered receive areas for the
    payload and the gradient exchange), mapped into every peer with CUDA IPC.  torch.distributed is only used
    here, once, to exchange the 64-byte IPC handles; the per-step exchanges are the kernels in csrc/lf_peer.cu
    and the fused push inside lf_step_mid (NVLink stores + system-scope flags, no NCCL call)."""

    def __init__(self, payload_bytes: int, grad_floats: int, pg=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self.rank, self.world = world(pg)
        if self.world > _lib.LF_MAX_RANKS:
            raise _lib.LfError(f"PeerComm supports up to {_lib.LF_MAX_RANKS} ranks on one NVSwitch domain")
        self.payload_bytes = (payload_bytes + 15) // 16 * 16
        self.grad_padded = (grad_floats + 3) // 4 * 4
        flags_bytes = _lib.LF_PEER_FLAGS_BYTES            # two barrier sets (stand-alone lf_peer_allreduce)
        self.off_payload = (flags_bytes + 255) // 256 * 256
        self.off_grad = self.off_payload + (2 * self.world * self.payload_bytes + 255) // 256 * 256
        total = self.off_grad + 2 * self.world * self.grad_padded * 4
        assert flags_bytes <= self.off_payload
        base = C.c_void_p()
        _lib.check(lib.lf_comm_alloc(total, C.byref(base)), "lf_comm_alloc")
        self.local = base.value
        # receive areas start armed with the all-ones sentinel of the flag-less exchanges (csrc/lf_peer.cuh); flags stay zero
        _lib.check(lib.lf_comm_fill(self.local + self.off_payload, 0xFF, total - self.off_payload), "lf_comm_fill")
        handle = C.create_string_buffer(64)
        _lib.check(lib.lf_comm_ipc_handle(self.local, handle), "lf_comm_ipc_handle")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=pg)
        self.bases = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases.append(self.local)
            else:
                p = C.c_void_p()
                _lib.check(lib.lf_comm_ipc_open(C.create_string_buffer(h, 64), C.byref(p)), "lf_comm_ipc_open")
                self.bases.append(p.value)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.epoch = torch.zeros(2, dtype=torch.int64, device=dev)
        # pinned, device-visible host flag: the kernels store 1 into it before they trap on a lost peer, so the host
        # can read the cause without a CUDA call (the context is unusable after a trap)
        self.error = torch.zeros(1, dtype=torch.int32).pin_memory()
        dist.barrier(group=pg)                      # every rank has mapped every buffer before the first push

    def fill(self, comm) -> None:
        """Populate an LfPeerComm struct."""
        comm.n_ranks, comm.rank = self.world, self.rank
        for r, b in enumerate(self.bases):
            comm.flags[r] = b
            comm.recv_payload[r] = b + self.off_payload
            comm.recv_grad[r] = b + self.off_grad
        comm.epoch = self.epoch.data_ptr()
        comm.error = self.error.data_ptr()

    def check(self) -> None:
        """Host-memory read (no synchronisation): raises once a kernel has given up on a peer."""
        if int(self.error[0]) != 0:
            from . import _lib
            raise _lib.LfError("peer exchange gave up waiting for another rank (a peer died or stalled for minutes); "
                               "the CUDA context of this rank has been aborted")

    def close(self) -> None:
        """Unmap the peers' buffers and free the local one.  Safe to call twice; not a collective (a peer that
        still has this rank's buffer mapped keeps the allocation alive until it closes its own mapping)."""
        if getattr(self, "bases", None) is None:
            return
        from . import _lib
        lib = _lib.load()
        torch.cuda.synchronize()
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                lib.lf_comm_ipc_close(b)
        if self.local:
            lib.lf_comm_free(self.local)
        self.bases, self.local = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
