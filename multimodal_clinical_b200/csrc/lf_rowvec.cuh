// Lane-group row helpers shared by the vectorised row kernels (lf_rows_vec.cu) and the fused tensor-pipe backward
// (lf_tc_bwd.cu): G lanes per sample, each lane owning NK float4 chunks of the row (columns 4(l + G k) .. +3).
#pragma once
#include "lf_common.cuh"
#include "lf_rowmath.cuh"

namespace lf {
namespace rowvec {

constexpr float kLog2e = 1.4426950408889634f;

template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ int group_min(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// sum over the 32/G lane groups of a warp (lanes with equal l): column sums owned by chunk position
template <int G>
__device__ __forceinline__ float across_groups_sum(float v) {
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One sample's row for one lane: NK chunks of 4 columns; chunk positions past the pitch read as 0.
template <int G, int NK>
struct Row {
  float v[NK * 4];
  __device__ __forceinline__ void load(const float* __restrict__ rowp, int l, int nq) {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int q = l + G * k;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nq) t = __ldg(reinterpret_cast<const float4*>(rowp) + q);
      v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
  }
};

// columns >= C (pitch padding the GEMM never writes, chunk positions past the pitch) := fill.  One compare per
// chunk on the common path; only the chunk that straddles C takes the per-element selects.
template <int G, int NK>
__device__ __forceinline__ void mask_cols(float (&v)[NK * 4], int l, int C, float fill) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c0 = 4 * (l + G * k);
    if (c0 + 3 >= C) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c0 + e >= C) v[4 * k + e] = fill;
    }
  }
}

template <int G, int NK>
__device__ __forceinline__ int col_of(int l, int i) { return 4 * (l + G * (i >> 2)) + (i & 3); }

// max over the row and the first column attaining it (torch.argmax semantics)
template <int G, int NK>
__device__ __forceinline__ void row_max_arg(const float (&v)[NK * 4], int l, float& mx, int& arg) {
  float lm = v[0];
#pragma unroll
  for (int i = 1; i < NK * 4; ++i) lm = fmaxf(lm, v[i]);
  mx = group_max<G>(lm);
  int idx = 0x7fffffff;
#pragma unroll
  for (int i = NK * 4 - 1; i >= 0; --i) idx = (v[i] == mx) ? col_of<G, NK>(l, i) : idx;
  arg = group_min<G>(idx);
}

template <int G, int NK>
__device__ __forceinline__ void store_chunks_f32(float* __restrict__ rowp, const float (&v)[NK * 4], int l, int C, bool vec) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c0 = 4 * (l + G * k);
    if (c0 >= C) continue;
    if (vec) {
      *reinterpret_cast<float4*>(rowp + c0) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c0 + e < C) rowp[c0 + e] = v[4 * k + e];
    }
  }
}

}  // namespace rowvec
}  // namespace lf
