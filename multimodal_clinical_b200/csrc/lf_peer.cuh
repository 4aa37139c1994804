// Device helpers of the peer-memory exchanges (protocol described in lf_peer.cu), shared by lf_peer.cu and lf_mid.cu.
#pragma once
#include <cooperative_groups.h>
#include "lf_common.cuh"

namespace lf {
namespace cg = cooperative_groups;

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr unsigned long long kPeerWaitNs = 120ull * 1000000000ull;

// Called by every thread of the (cooperatively launched) grid after its peer stores.  flags_set: 0 = payload, 1 = gradients.
__device__ __forceinline__ void peer_barrier(const LfPeerComm& c, int flags_set, long long epoch, cg::grid_group& grid) {
  __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && (int)threadIdx.x < c.n_ranks) {
    const int r = threadIdx.x;
    st_release_sys((long long*)c.flags[r] + flags_set * LF_MAX_RANKS + c.rank, epoch);       // my flag on rank r
    const long long* mine = (const long long*)c.flags[c.rank] + flags_set * LF_MAX_RANKS + r;  // rank r's flag here
    // Wait like a collective does: rank skew of seconds is routine in training (checkpoint writes, data-loader
    // respawns, first-iteration lazy init), so the wait is NOT bounded by a short timeout.  Only after
    // kPeerWaitNs (2 minutes: a peer that is this late is gone) the kernel records the failure and TRAPS: the
    // launch fails loudly (sticky CUDA error at the next synchronisation, PeerComm.check() reports the cause) and
    // nothing after the barrier -- the receive area, epoch[], the replicated EMA / History -- is touched.
    unsigned long long t0 = 0, spins = 0;
    while (ld_acquire_sys(mine) < epoch) {
      __nanosleep(64);
      if ((++spins & 1023) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > kPeerWaitNs) {
          *(volatile int*)c.error = 1;
          __threadfence_system();
          __trap();
        }
      }
    }
  }
  grid.sync();
}

// Push `bytes` (multiple of 16) from src into slot [parity][rank] of every rank's receive area.
__device__ __forceinline__ void peer_push(const LfPeerComm& c, void* const recv[LF_MAX_RANKS], const void* src, size_t bytes, int parity,
                          int tid, int nthr) {
  const size_t n16 = bytes / 16;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  for (int r = 0; r < c.n_ranks; ++r) {
    uint4* d = reinterpret_cast<uint4*>((char*)recv[r] + ((size_t)parity * c.n_ranks + c.rank) * bytes);
    size_t i = tid;
    for (; i + 3 * (size_t)nthr < n16; i += 4 * (size_t)nthr) {          // four loads in flight, then four peer stores
      const uint4 v0 = s[i], v1 = s[i + nthr], v2 = s[i + 2 * (size_t)nthr], v3 = s[i + 3 * (size_t)nthr];
      d[i] = v0; d[i + nthr] = v1; d[i + 2 * (size_t)nthr] = v2; d[i + 3 * (size_t)nthr] = v3;
    }
    for (; i < n16; i += nthr) d[i] = s[i];
  }
}


}  // namespace lf
