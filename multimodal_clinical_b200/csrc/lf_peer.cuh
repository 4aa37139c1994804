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
// ---- sentinel ("data is the flag") exchanges: no fence, no flag, no grid barrier on the critical path.
// A receive slot is pre-filled with all-ones words (a NaN pattern no arithmetic produces; senders canonicalise the one
// colliding payload).  The sender stores plain data words into the peer's slot; the receiver polls the words it needs
// with volatile (L1-bypassing) loads until none of them is the sentinel, uses them, and re-arms the slot with the
// sentinel.  Every 32-bit (or 64-bit) word validates itself, so no store atomicity beyond the word is assumed.  Slots
// are double-buffered by epoch parity: a peer writes slot [e & 1] again at epoch e + 2, which it can only reach after it
// has received this rank's epoch e + 1 data -- sent after this rank finished reading (and re-arming) epoch e.
constexpr unsigned kSentinel32 = 0xFFFFFFFFu;
constexpr unsigned long long kSentinel64 = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float clean_f32(float v) { return __float_as_uint(v) == kSentinel32 ? __uint_as_float(0x7FC00000u) : v; }
__device__ __forceinline__ double clean_f64(double v) {
  return (unsigned long long)__double_as_longlong(v) == kSentinel64 ? __longlong_as_double(0x7FF8000000000000ll) : v;
}
__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_volatile_u32(const void* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// A collective's wait: rank skew of seconds is routine; only a peer missing for minutes aborts the launch loudly
// (sticky CUDA error; the host-visible flag names the cause).  Call every iteration of a polling loop.
static __device__ __noinline__ unsigned long long peer_spin_check(int* error, unsigned long long t0) {
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (t0 == 0) return now;
  if (now - t0 > kPeerWaitNs) { *(volatile int*)error = 1; __threadfence_system(); __trap(); }
  return t0;
}
struct PeerSpin {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  // `error` = LfPeerComm.error (passed as a plain pointer: taking the address of a kernel parameter would make every
  // thread copy the parameter block to local memory)
  __device__ __forceinline__ void wait(int* error) {
    __nanosleep(32);
    if ((++spins & 4095u) == 0) t0 = peer_spin_check(error, t0);     // out of line: the polling loops are inlined in many places
  }
};

// Same for one piece of the slot: `bytes` (multiple of 16) from src to byte offset `off` of slot [parity][rank] (slot size slot_bytes).
__device__ __forceinline__ void peer_push_at(const LfPeerComm& c, void* const recv[LF_MAX_RANKS], const void* src, size_t off, size_t bytes,
                                             size_t slot_bytes, int parity, int tid, int nthr) {
  const size_t n16 = bytes / 16;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  const size_t slot = ((size_t)parity * c.n_ranks + c.rank) * slot_bytes + off;
  for (size_t i = tid; i < n16; i += nthr) {
    const uint4 v = s[i];                                   // one load, one store per peer
#pragma unroll
    for (int r = 0; r < LF_MAX_RANKS; ++r)
      if (r < c.n_ranks) reinterpret_cast<uint4*>((char*)recv[r] + slot)[i] = v;
  }
}

}  // namespace lf
