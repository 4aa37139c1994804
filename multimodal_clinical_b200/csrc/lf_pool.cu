// Feature-side pooling in front of the heads: the global average of an encoder's feature
// maps over space -- and, for the frame-stacked visual stream, over the T frames of a clip --
//   a (B, C, H, W)      -> (B, C)      F.adaptive_avg_pool2d(a, 1)             cremad/joint_model_qmf.py:52, 54
//   v (B*T, C, H, W)    -> (B, C)      view(B, T, C, H, W).permute(0, 2, 1, 3, 4) + adaptive_avg_pool3d(v, 1)   :48-53
// and its backward (every input element receives dOut[b, c] / (T H W)).
// Pure HBM streaming: one warp owns four consecutive channels of a sample, lanes stride the H*W plane of each (frame,
// channel) -- 4 x ceil(HW / 32) independent loads in flight per lane -- and a butterfly adds the lanes.  fp32 or bf16
// maps (what the encoders emit under autocast); accumulation in fp32, output in the input's type.
#include <cuda_bf16.h>
#include "lf_common.cuh"

namespace lf {

constexpr int kPoolCh = 4;          // channels per warp

template <class T> __device__ __forceinline__ float pool_ld(const T* p);
template <> __device__ __forceinline__ float pool_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float pool_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <class T> __device__ __forceinline__ void pool_st(T* p, float v);
template <> __device__ __forceinline__ void pool_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void pool_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <class T>
__global__ void __launch_bounds__(256) pool_mean_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Tn, int C, int HW, float inv) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  const int groups = (C + kPoolCh - 1) / kPoolCh;
  const long long units = (long long)B * groups;
  for (long long u = (long long)blockIdx.x * nwarp + warp; u < units; u += (long long)gridDim.x * nwarp) {
    const int b = (int)(u / groups), c0 = (int)(u - (long long)b * groups) * kPoolCh;
    float s[kPoolCh];
#pragma unroll
    for (int j = 0; j < kPoolCh; ++j) s[j] = 0.f;
    for (int t = 0; t < Tn; ++t) {
      const T* plane = in + ((size_t)((size_t)b * Tn + t) * C + c0) * HW;       // kPoolCh consecutive channel planes
      for (int k = lane; k < HW; k += 32) {
#pragma unroll
        for (int j = 0; j < kPoolCh; ++j)
          if (c0 + j < C) s[j] += pool_ld(plane + (size_t)j * HW + k);
      }
    }
#pragma unroll
    for (int j = 0; j < kPoolCh; ++j) s[j] = warp_sum(s[j]);
    if (lane < kPoolCh && c0 + lane < C) {
      float v = s[0];
#pragma unroll
      for (int j = 1; j < kPoolCh; ++j) if (lane == j) v = s[j];
      pool_st(out + (size_t)b * C + c0 + lane, v * inv);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(256) pool_mean_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, int B, int Tn, int C, int HW, float inv) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  const int groups = (C + kPoolCh - 1) / kPoolCh;
  const long long units = (long long)B * groups;
  for (long long u = (long long)blockIdx.x * nwarp + warp; u < units; u += (long long)gridDim.x * nwarp) {
    const int b = (int)(u / groups), c0 = (int)(u - (long long)b * groups) * kPoolCh;
    float g[kPoolCh];
#pragma unroll
    for (int j = 0; j < kPoolCh; ++j) g[j] = c0 + j < C ? pool_ld(dout + (size_t)b * C + c0 + j) * inv : 0.f;
    for (int t = 0; t < Tn; ++t) {
      T* plane = din + ((size_t)((size_t)b * Tn + t) * C + c0) * HW;
      for (int k = lane; k < HW; k += 32) {
#pragma unroll
        for (int j = 0; j < kPoolCh; ++j)
          if (c0 + j < C) pool_st(plane + (size_t)j * HW + k, g[j]);
      }
    }
  }
}

template <class T>
static int pool_launch(const void* in, void* out, int B, int Tn, int C, int HW, bool backward, cudaStream_t s) {
  const long long units = (long long)B * ((C + kPoolCh - 1) / kPoolCh);
  long long nb = (units + 7) / 8;
  if (nb > 148 * 8) nb = 148 * 8;                 // 8 CTAs of 8 warps per SM: 64 warps x 4 channel planes in flight
  if (nb < 1) nb = 1;
  const float inv = 1.f / ((float)Tn * (float)HW);
  if (!backward) {
    LF_LAUNCH("pool_mean", s, (pool_mean_kernel<T><<<(int)nb, 256, 0, s>>>((const T*)in, (T*)out, B, Tn, C, HW, inv)));
    return check_launch("pool_mean_kernel");
  }
  LF_LAUNCH("pool_mean_bwd", s, (pool_mean_bwd_kernel<T><<<(int)nb, 256, 0, s>>>((const T*)in, (T*)out, B, Tn, C, HW, inv)));
  return check_launch("pool_mean_bwd_kernel");
}

}  // namespace lf

using namespace lf;

extern "C" int lf_pool_mean(const void* maps, void* pooled, int32_t batch, int32_t frames, int32_t channels, int32_t hw,
                            int32_t elem_bytes, void* stream) {
  if (!maps || !pooled || batch < 1 || frames < 1 || channels < 1 || hw < 1 || (elem_bytes != 4 && elem_bytes != 2)) {
    set_error("lf_pool_mean: bad argument");
    return LF_ERR_BAD_ARG;
  }
  return elem_bytes == 4 ? pool_launch<float>(maps, pooled, batch, frames, channels, hw, false, (cudaStream_t)stream)
                         : pool_launch<__nv_bfloat16>(maps, pooled, batch, frames, channels, hw, false, (cudaStream_t)stream);
}

extern "C" int lf_pool_mean_backward(const void* dpooled, void* dmaps, int32_t batch, int32_t frames, int32_t channels, int32_t hw,
                                     int32_t elem_bytes, void* stream) {
  if (!dpooled || !dmaps || batch < 1 || frames < 1 || channels < 1 || hw < 1 || (elem_bytes != 4 && elem_bytes != 2)) {
    set_error("lf_pool_mean_backward: bad argument");
    return LF_ERR_BAD_ARG;
  }
  return elem_bytes == 4 ? pool_launch<float>(dpooled, dmaps, batch, frames, channels, hw, true, (cudaStream_t)stream)
                         : pool_launch<__nv_bfloat16>(dpooled, dmaps, batch, frames, channels, hw, true, (cudaStream_t)stream);
}
