// OGM-GE gradient modulation over the 4-D (conv) gradients of one encoder:
//   sigma_t = unbiased std of tensor t (unscaled) + 1e-8
//   OGM_GE: g <- k g + sigma_t xi     OGM: g <- k g     noise: g <- g + sigma_t xi,   xi ~ N(0,1)
// Reference: existing_algos/OGM_GE.py:42-54 — there one .std().item() host sync and four kernels per
// tensor (40 syncs for two ResNet18s).  Here: one multi-tensor statistics pass and one multi-tensor
// update pass, coefficient read from device memory, Philox4x32-10 + Box-Muller noise generated in
// registers.  HBM-bound: 12 B/param (4 read for std, 4 read + 4 write for the update); the second read
// hits L2 when an encoder's conv gradients (45 MB for ResNet18) fit the 126 MB L2.
#include "lf_common.cuh"

namespace lf {

struct ModTable {
  int count;
  int mode;
  float* data[LF_MAX_TENSORS];
  long long numel[LF_MAX_TENSORS];
  long long start[LF_MAX_TENSORS];   // first element's position in the encoder-wide noise stream (multiple of 4)
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: any element's draw is addressable ----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float4 normal4(unsigned long long group, unsigned long long seed,
                                          unsigned long long offset) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)offset,
                                           (uint32_t)(offset >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
  const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
  // (x + 0.5) * 2^-32 can round to 1.0f for the top few x: clamp so log() stays finite and <= 0
  const float a0 = sqrtf(-2.f * __logf(fminf(u0, 0.99999994f)));
  const float a1 = sqrtf(-2.f * __logf(fminf(u2, 0.99999994f)));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  return make_float4(a0 * c0, a0 * s0, a1 * c1, a1 * s1);
}

__global__ void __launch_bounds__(256) modulate_stats_kernel(ModTable tb, double* __restrict__ ws) {
  const int t = blockIdx.y;
  const float* __restrict__ g = tb.data[t];
  const long long n = tb.numel[t];
  double s = 0.0, ss = 0.0;
  const bool vec = ((((uintptr_t)g) & 15) == 0);
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);   // keep in L2 for the update pass
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += v; ss += v * v;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  __shared__ double a[8], b[8];
  if (threadIdx.x % 32 == 0) { a[threadIdx.x / 32] = s; b[threadIdx.x / 32] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s += a[w]; ss += b[w]; }
    if (s != 0.0 || ss != 0.0) { atomicAdd(&ws[2 * t], s); atomicAdd(&ws[2 * t + 1], ss); }
  }
}

__global__ void __launch_bounds__(256) modulate_apply_kernel(ModTable tb, const double* __restrict__ ws,
                                                             const float* __restrict__ coeff_dev,
                                                             unsigned long long seed, unsigned long long offset) {
  const int t = blockIdx.y;
  float* __restrict__ g = tb.data[t];
  const long long n = tb.numel[t];
  const int mode = tb.mode;
  const float k = (mode == LF_MOD_NOISE) ? 1.f : coeff_dev[0];
  float sigma = 0.f;
  if (mode != LF_MOD_OGM) {
    const double mean = ws[2 * t] / (double)n;
    double var = (ws[2 * t + 1] - (double)n * mean * mean) / (double)(n - 1);   // unbiased (torch.std default)
    if (var < 0.0) var = 0.0;
    sigma = (float)((double)(float)sqrt(var) + 1e-8);                            // OGM_GE.py:50
  }
  const unsigned long long gbase = (unsigned long long)tb.start[t] / 4;
  const bool vec = ((((uintptr_t)g) & 15) == 0);
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    if (mode == LF_MOD_OGM) {
      v.x *= k; v.y *= k; v.z *= k; v.w *= k;
    } else {
      const float4 z = normal4(gbase + (unsigned long long)i, seed, offset);
      v.x = v.x * k + z.x * sigma; v.y = v.y * k + z.y * sigma;
      v.z = v.z * k + z.z * sigma; v.w = v.w * k + z.w * sigma;
    }
    reinterpret_cast<float4*>(g)[i] = v;
  }
  // scalar tail / unaligned tensors: one Philox call per group of 4, lanes pick their component
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = g[i];
    if (mode == LF_MOD_OGM) v *= k;
    else {
      const float4 z = normal4(gbase + (unsigned long long)(i / 4), seed, offset);
      const int r = (int)(i & 3);
      v = v * k + (r == 0 ? z.x : r == 1 ? z.y : r == 2 ? z.z : z.w) * sigma;
    }
    g[i] = v;
  }
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_modulate_workspace_bytes(void) { return sizeof(double) * 2 * LF_MAX_TENSORS; }

extern "C" int lf_ogm_modulate(const LfTensorList* list, const float* coeff_dev, int32_t mode, uint64_t seed,
                               uint64_t offset, void* workspace, size_t workspace_bytes, void* stream) {
  if (!list || list->count < 0 || list->count > LF_MAX_TENSORS) { set_error("lf_ogm_modulate: bad tensor list"); return LF_ERR_BAD_ARG; }
  if (mode < LF_MOD_OGM_GE || mode > LF_MOD_NOISE) { set_error("lf_ogm_modulate: bad mode %d", mode); return LF_ERR_BAD_ARG; }
  if (mode != LF_MOD_NOISE && !coeff_dev) { set_error("lf_ogm_modulate: coeff_dev is null"); return LF_ERR_BAD_ARG; }
  if (list->count == 0) return LF_OK;   // Food101 / MLP encoders: no 4-D grads -> no-op like the reference
  if (!workspace || workspace_bytes < lf_modulate_workspace_bytes()) { set_error("lf_ogm_modulate: workspace too small"); return LF_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  ModTable tb;
  tb.count = list->count; tb.mode = mode;
  long long pos = 0, maxn = 0;
  for (int i = 0; i < list->count; ++i) {
    if (!list->data[i] || list->numel[i] < 0) { set_error("lf_ogm_modulate: bad tensor %d", i); return LF_ERR_BAD_ARG; }
    tb.data[i] = list->data[i]; tb.numel[i] = list->numel[i]; tb.start[i] = pos;
    pos += (list->numel[i] + 3) / 4 * 4;
    if (list->numel[i] > maxn) maxn = list->numel[i];
  }
  int gx = div_up(maxn, 256 * 4 * 4);
  if (gx < 1) gx = 1;
  if (gx > 296) gx = 296;
  dim3 grid(gx, list->count);
  if (mode != LF_MOD_OGM) {
    cudaMemsetAsync(workspace, 0, lf_modulate_workspace_bytes(), s);
    LF_LAUNCH("modulate_stats", s, (modulate_stats_kernel<<<grid, 256, 0, s>>>(tb, (double*)workspace)));
  }
  LF_LAUNCH("modulate_apply", s, (modulate_apply_kernel<<<grid, 256, 0, s>>>(tb, (const double*)workspace, coeff_dev, seed, offset)));
  return check_launch("lf_ogm_modulate");
}
