// Warp-level row math shared by the register-resident row kernels (lf_rows_reg.cu) and the fused
// narrow-head kernel (lf_narrow.cu): lane = class, NCH values per lane.
#pragma once
#include "lf_common.cuh"

namespace lf {

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f + 0.f);                 // + 0.f folds -0 into +0 (they compare equal)
  return b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float(u ^ ((u & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}
// warp max + first index attaining it (torch.argmax semantics); v[k] holds class lane + 32 k
template <int NCH>
__device__ __forceinline__ void warp_max_arg(const float (&v)[NCH], int lane, float& mx, int& arg) {
  float lm = v[0];
#pragma unroll
  for (int k = 1; k < NCH; ++k) lm = fmaxf(lm, v[k]);
  mx = ord2f(__reduce_max_sync(kFull, f2ord(lm)));
  unsigned idx = 0x7fffffffu;
#pragma unroll
  for (int k = NCH - 1; k >= 0; --k) idx = (v[k] == mx) ? (unsigned)(lane + 32 * k) : idx;
  arg = (int)__reduce_min_sync(kFull, idx);
}
// exp(x - m) with the scale folded into one FFMA: ex2(x * log2e - m * log2e)
__device__ __forceinline__ float exp_sub(float x, float m_log2e) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(x, 1.4426950408889634f, -m_log2e)));
  return r;
}
template <int NCH>
__device__ __forceinline__ float pick_class(const float (&v)[NCH], int y) {   // value of class y, broadcast
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) r = ((y >> 5) == k) ? v[k] : r;
  return __shfl_sync(kFull, r, y & 31);
}
__device__ __forceinline__ void warp_sum3(float& a, float& b, float& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o); b += __shfl_xor_sync(kFull, b, o); c += __shfl_xor_sync(kFull, c, o);
  }
}

// Sum v[i] over the warp for all i < CMAX with ~CMAX shuffles: at every step a lane keeps one half of its
// values and hands the other half to its partner.  Afterwards every lane holds the total of class
// lane / (32 / CMAX).
template <int CMAX>
__device__ __forceinline__ float transpose_reduce(float (&v)[CMAX], int lane) {
  int n = CMAX;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < CMAX / 2; ++i)
        if (i < n) {
          const float keep = up ? v[i + n] : v[i];
          const float send = up ? v[i] : v[i + n];
          v[i] = keep + __shfl_xor_sync(kFull, send, s);
        }
    } else {
      v[0] += __shfl_xor_sync(kFull, v[0], s);
    }
  }
  return v[0];
}

}  // namespace lf
