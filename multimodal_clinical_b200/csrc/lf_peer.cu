// Peer-memory exchanges of the batch-sharded step over NVLink / NVSwitch (one process per GPU).
//
// Both exchanges of a step are small (payload: 16 B/sample + statistics; gradients: 2(CD+C) floats), i.e.
// latency-bound, so they are written as ONE-SHOT pushes into peer memory instead of ring collectives:
//   * every rank stores its contribution into slot [parity][rank] of every peer's receive area
//     (plain st.global on peer-mapped pointers), fences at system scope, and raises its flag on every peer
//     (st.release.sys); it then spins (ld.acquire.sys) until all peers' flags show the current epoch;
//   * the reduction / the consumer then reads only LOCAL memory, in rank order, so all ranks compute
//     bit-identical sums.
// Receive areas are double-buffered by epoch parity: a rank can run at most one epoch ahead of the slowest
// peer (it needs that peer's flag for the epoch in between), so slot [e & 1] is never overwritten while a
// peer still reads it.  The epoch counter is device-resident, which makes the launches replayable from a
// CUDA graph.  The flag wait is that of a collective: it tolerates rank skew of seconds; only a peer missing for
// minutes makes the kernel record the failure in a host-visible flag and trap (lf_peer.cuh::peer_barrier).
#include <cooperative_groups.h>
#include <string.h>
#include "lf_common.cuh"
#include "lf_peer.cuh"

namespace cg = cooperative_groups;

namespace lf {

__global__ void __launch_bounds__(512) peer_allreduce_kernel(LfPeerReduceArgs a) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  const LfPeerComm& c = a.comm;
  const long long epoch = c.epoch[1] + 1;
  const int parity = (int)(epoch & 1);
  const size_t bytes = (size_t)a.n_padded * sizeof(float);
  peer_push(c, a.comm.recv_grad, a.buf, bytes, parity, tid, nthr);
  peer_barrier(c, 1, epoch, grid);
  const float4* base = reinterpret_cast<const float4*>((const float*)c.recv_grad[c.rank] + (size_t)parity * c.n_ranks * a.n_padded);
  const int n4 = a.n_padded / 4;
  for (int i = tid; i < n4; i += nthr) {
    float4 v[LF_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < LF_MAX_RANKS; ++r)
      if (r < c.n_ranks) v[r] = base[(size_t)r * n4 + i];                // all ranks' loads in flight together
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < LF_MAX_RANKS; ++r)                              // rank order: identical on every rank
      if (r < c.n_ranks) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    const float e[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = 4 * i + k;
      if (j < a.n) {
        a.buf[j] = e[k];
        if (a.tail_dst && j >= a.n - a.tail_n) a.tail_dst[j - (a.n - a.tail_n)] = (double)e[k];
      }
    }
  }
  grid.sync();
  if (tid == 0) c.epoch[1] = epoch;
}

}  // namespace lf

using namespace lf;

extern "C" int lf_comm_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) { set_error("lf_comm_alloc: bad argument"); return LF_ERR_BAD_ARG; }
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) { set_error("lf_comm_alloc: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  return LF_OK;
}
extern "C" int lf_comm_fill(void* ptr, int32_t byte_value, size_t bytes) {
  if (!ptr || bytes == 0) { set_error("lf_comm_fill: bad argument"); return LF_ERR_BAD_ARG; }
  cudaError_t e = cudaMemset(ptr, byte_value, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("lf_comm_fill: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  return LF_OK;
}
extern "C" int lf_comm_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? LF_OK : LF_ERR_CUDA; }
extern "C" int lf_comm_ipc_handle(void* ptr, void* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
  return LF_OK;
}
extern "C" int lf_comm_ipc_open(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  return LF_OK;
}
extern "C" int lf_comm_ipc_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? LF_OK : LF_ERR_CUDA; }

extern "C" int lf_peer_allreduce(const LfPeerReduceArgs* a, void* stream) {
  if (!a || !a->buf || a->n < 1 || a->n_padded < a->n || a->n_padded % 4 || a->comm.n_ranks < 1 ||
      a->comm.n_ranks > LF_MAX_RANKS || !a->comm.epoch || !a->comm.error) { set_error("lf_peer_allreduce: bad argument"); return LF_ERR_BAD_ARG; }
  cudaLaunchConfig_t cfg = {};
  int ctas = div_up(a->n_padded / 4, 512 * 2);          // ~2 float4 per thread per rank
  if (ctas > 64) ctas = 64;
  if (ctas < 1) ctas = 1;
  cfg.gridDim = dim3(ctas, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;          // grid-wide barrier between push, flag exchange and reduction
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaSuccess;
  LF_LAUNCH("peer_allreduce", cfg.stream, (e = cudaLaunchKernelEx(&cfg, peer_allreduce_kernel, *a)));
  if (e != cudaSuccess) { set_error("lf_peer_allreduce: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  return check_launch("lf_peer_allreduce");
}
