// OGM-GE gradient modulation over the 4-D (conv) gradients of one encoder:
//   sigma_t = unbiased std of tensor t (unscaled) + 1e-8
//   OGM_GE: g <- k g + sigma_t xi     OGM: g <- k g     noise: g <- g + sigma_t xi,   xi ~ N(0,1)
// Reference: existing_algos/OGM_GE.py:42-54 — there one .std().item() host sync and four kernels per
// tensor (40 syncs for two ResNet18s).  Here: one multi-tensor statistics pass and one multi-tensor
// update pass, coefficient read from device memory, Philox4x32-10 + Box-Muller noise generated in
// registers.  HBM-bound: 12 B/param (4 read for std, 4 read + 4 write for the update); the second read
// hits L2 when an encoder's conv gradients (45 MB for ResNet18) fit the 126 MB L2.
#include "lf_common.cuh"
#include "lf_philox.cuh"

namespace lf {

struct ModTable {
  int count;
  int mode;
  float* data[LF_MAX_TENSORS];
  long long numel[LF_MAX_TENSORS];
  long long start[LF_MAX_TENSORS];   // first element's position in the encoder-wide noise stream (multiple of 4)
  int chunk0[LF_MAX_TENSORS + 1];    // work items: tensor t owns chunks [chunk0[t], chunk0[t+1]) of kModChunk elements
};
constexpr int kModChunk = 16384;     // elements per work item: 256 threads x 4 rounds x 4 float4

__device__ __forceinline__ int mod_find_tensor(const ModTable& tb, int item) {
  int lo = 0, hi = tb.count - 1;       // largest t with chunk0[t] <= item
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tb.chunk0[mid] <= item) lo = mid; else hi = mid - 1;
  }
  return lo;
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

// Persistent grid over (tensor, 16K-element chunk) work items: a ResNet18's 20 conv tensors span 3 K .. 2.4 M
// elements, so one CTA per tensor slice would leave most CTAs nearly empty.  Four independent 128-bit loads in
// flight per thread (tools/membench.cu: ~64 KB in flight per SM are needed to cover HBM latency).
__global__ void __launch_bounds__(256) modulate_stats_kernel(ModTable tb, double* __restrict__ ws) {
  const int total = tb.chunk0[tb.count];
  __shared__ double a[8], b[8];
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int t = mod_find_tensor(tb, item);
    const float* __restrict__ g = tb.data[t];
    const long long n = tb.numel[t];
    const long long e0 = (long long)(item - tb.chunk0[t]) * kModChunk, e1 = min(n, e0 + kModChunk);
    double s = 0.0, ss = 0.0;
    if ((((uintptr_t)g) & 15) == 0) {
      const float4* g4 = reinterpret_cast<const float4*>(g);
      const long long q0 = e0 / 4, q1 = e1 / 4;                      // e0 is a multiple of 4
      for (long long i = q0 + threadIdx.x; i < q1; i += 4 * 256) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i + u * 256 < q1) ? __ldg(g4 + i + u * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          s += (double)v[u].x + (double)v[u].y + (double)v[u].z + (double)v[u].w;
          ss += (double)v[u].x * v[u].x + (double)v[u].y * v[u].y + (double)v[u].z * v[u].z + (double)v[u].w * v[u].w;
        }
      }
      for (long long i = q1 * 4 + threadIdx.x; i < e1; i += 256) { const double v = g[i]; s += v; ss += v * v; }
    } else {
      for (long long i = e0 + threadIdx.x; i < e1; i += 256) { const double v = g[i]; s += v; ss += v * v; }
    }
    s = warp_sum(s); ss = warp_sum(ss);
    __syncthreads();
    if (threadIdx.x % 32 == 0) { a[threadIdx.x / 32] = s; b[threadIdx.x / 32] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { s += a[w]; ss += b[w]; }
      if (s != 0.0 || ss != 0.0) { atomicAdd(&ws[2 * t], s); atomicAdd(&ws[2 * t + 1], ss); }
    }
  }
}

__global__ void __launch_bounds__(256) modulate_apply_kernel(ModTable tb, const double* __restrict__ ws,
                                                             const float* __restrict__ coeff_dev,
                                                             unsigned long long seed, unsigned long long offset) {
  const int total = tb.chunk0[tb.count];
  const int mode = tb.mode;
  const float k = (mode == LF_MOD_NOISE) ? 1.f : coeff_dev[0];
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int t = mod_find_tensor(tb, item);
    float* __restrict__ g = tb.data[t];
    const long long n = tb.numel[t];
    const long long e0 = (long long)(item - tb.chunk0[t]) * kModChunk, e1 = min(n, e0 + kModChunk);
    float sigma = 0.f;
    if (mode != LF_MOD_OGM) {
      const double mean = ws[2 * t] / (double)n;
      double var = (ws[2 * t + 1] - (double)n * mean * mean) / (double)(n - 1);   // unbiased (torch.std default)
      if (var < 0.0) var = 0.0;
      sigma = (float)((double)(float)sqrt(var) + 1e-8);                            // OGM_GE.py:50
    }
    const unsigned long long gbase = (unsigned long long)tb.start[t] / 4;
    if ((((uintptr_t)g) & 15) == 0) {
      float4* g4 = reinterpret_cast<float4*>(g);
      const long long q0 = e0 / 4, q1 = e1 / 4;
      for (long long i0 = q0 + threadIdx.x; i0 < q1; i0 += 4 * 256) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * 256 < q1) v[u] = g4[i0 + u * 256];      // loads first, all in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * 256;
          if (i >= q1) break;
          if (mode == LF_MOD_OGM) {
            v[u].x *= k; v[u].y *= k; v[u].z *= k; v[u].w *= k;
          } else {
            const float4 z = normal4(gbase + (unsigned long long)i, seed, offset);
            v[u].x = v[u].x * k + z.x * sigma; v[u].y = v[u].y * k + z.y * sigma;
            v[u].z = v[u].z * k + z.z * sigma; v[u].w = v[u].w * k + z.w * sigma;
          }
          g4[i] = v[u];
        }
      }
      for (long long i = q1 * 4 + threadIdx.x; i < e1; i += 256) {     // scalar tail: lanes pick their Philox component
        float v = g[i];
        if (mode == LF_MOD_OGM) v *= k;
        else {
          const float4 z = normal4(gbase + (unsigned long long)(i / 4), seed, offset);
          const int r = (int)(i & 3);
          v = v * k + (r == 0 ? z.x : r == 1 ? z.y : r == 2 ? z.z : z.w) * sigma;
        }
        g[i] = v;
      }
    } else {
      for (long long i = e0 + threadIdx.x; i < e1; i += 256) {
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
  long long pos = 0;
  int chunks = 0;
  for (int i = 0; i < list->count; ++i) {
    if (!list->data[i] || list->numel[i] < 0) { set_error("lf_ogm_modulate: bad tensor %d", i); return LF_ERR_BAD_ARG; }
    tb.data[i] = list->data[i]; tb.numel[i] = list->numel[i]; tb.start[i] = pos;
    pos += (list->numel[i] + 3) / 4 * 4;
    tb.chunk0[i] = chunks;
    chunks += div_up(list->numel[i], kModChunk);
  }
  tb.chunk0[list->count] = chunks;
  if (chunks == 0) return LF_OK;
  const int grid = chunks < 148 * 8 ? chunks : 148 * 8;
  if (mode != LF_MOD_OGM) {
    cudaMemsetAsync(workspace, 0, lf_modulate_workspace_bytes(), s);
    LF_LAUNCH("modulate_stats", s, (modulate_stats_kernel<<<grid, 256, 0, s>>>(tb, (double*)workspace)));
  }
  LF_LAUNCH("modulate_apply", s, (modulate_apply_kernel<<<grid, 256, 0, s>>>(tb, (const double*)workspace, coeff_dev, seed, offset)));
  return check_launch("lf_ogm_modulate");
}
