// Counter-based random numbers shared by the OGM-GE noise (lf_modulate.cu) and the dropout epilogue (lf_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lf {

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: any element's draw is addressable ----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * ctr.x, p1 = (unsigned long long)M1 * ctr.z;   // one IMAD.WIDE each
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// Random words of the dropout epilogue (lf_tc.cu::relu_dropout): element e (= row * N + col) of a tensor draws the 16-bit
// half (e & 7) of the Philox block whose counter is (e >> 3, rng_offset); it is KEPT when the half >= round(p * 65536).
__device__ __forceinline__ uint4 dropout_words(unsigned long long group, unsigned long long seed, unsigned long long offset) {
  return philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

}  // namespace lf
