// Exact-fp32 FMA GEMMs for the three head products:
//   logits  Z_m  = F_m W_m^T + b_m      (B x C, K = D)
//   dfeat   dF_m = dZ_m W_m             (B x D, K = C)
//   dweight dW_m = dZ_m^T F_m           (C x D, K = B, split-K with a fixed-order reduction)
// This is the LF_PREC_FP32 path: plain FFMA, sequential k order inside a tile, deterministic.
// The narrow-head fused kernels (lf_narrow.cu) and the tcgen05 path (lf_tc.cu) replace it where they
// apply; this file stays as the generic, always-correct implementation.
#include <cuda_bf16.h>
#include "lf_common.cuh"
#include "lf_gemm.cuh"

namespace lf {

constexpr int BK = 16;

template <bool A_KC, bool B_KC, int TM, int TN>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int batch = blockIdx.z / g.splits;
  const int split = blockIdx.z % g.splits;
  const float* __restrict__ A = g.A[batch];
  const float* __restrict__ Bm = g.B[batch];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = split * g.k_chunk;
  const int k_end = min(g.K, k_begin + g.k_chunk);
  const int t = threadIdx.x, ty = t / 16, tx = t % 16;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage A tile: As[k][m]
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int e = t + i * 256;
      int m, k;
      if (A_KC) { m = e / BK; k = e % BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < k_end)
        v = A_KC ? A[(size_t)gm * g.lda + gk] : A[(size_t)gk * g.lda + gm];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < TN; ++i) {
      const int e = t + i * 256;
      int n, k;
      if (B_KC) { n = e / BK; k = e % BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < k_end)
        v = B_KC ? Bm[(size_t)gn * g.ldb + gk] : Bm[(size_t)gk * g.ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* __restrict__ Cout = g.C[batch] + (size_t)split * g.split_stride;
  const float* __restrict__ bias = g.bias[batch];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn < g.N) Cout[(size_t)gm * g.ldc + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
    }
  }
}

template <bool A_KC, bool B_KC, int TM, int TN>
static int launch(const GemmArgs& g, int nbatch, cudaStream_t s) {
  dim3 grid(div_up(g.M, 16 * TM), div_up(g.N, 16 * TN), nbatch * g.splits);
  LF_LAUNCH(g.name, s, (sgemm_kernel<A_KC, B_KC, TM, TN><<<grid, 256, 0, s>>>(g)));
  return check_launch(g.name);
}

// Z = F W^T + b : A = F (k-contig), B = W (k-contig)
int gemm_logits(GemmArgs g, int nbatch, cudaStream_t s) {
  g.splits = 1; g.k_chunk = g.K; g.split_stride = 0; g.name = "sgemm_logits";
  if (g.N <= 16) return launch<true, true, 4, 1>(g, nbatch, s);
  return launch<true, true, 4, 4>(g, nbatch, s);
}
// dF = dZ W : A = dZ (k-contig, K = C), B = W (n-contig)
int gemm_dfeat(GemmArgs g, int nbatch, cudaStream_t s) {
  g.splits = 1; g.k_chunk = g.K; g.split_stride = 0; g.name = "sgemm_dfeat";
  return launch<true, false, 4, 4>(g, nbatch, s);
}
// dW partials = dZ^T F : A = dZ (m-contig), B = F (n-contig); K = B split over blockIdx.z
int gemm_dweight(GemmArgs g, int nbatch, cudaStream_t s) {
  g.name = "sgemm_dweight";
  if (g.M <= 16) return launch<false, false, 1, 4>(g, nbatch, s);
  return launch<false, false, 4, 4>(g, nbatch, s);
}

// fp32 head weights -> bf16 [2][n] (LF_PREC_BF16): the per-step cast autocast performs for nn.Linear
__global__ void cast_weights_bf16_kernel(const float* __restrict__ w0, const float* __restrict__ w1, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (i >= n) return;
  out[(size_t)blockIdx.y * n + i] = __float2bfloat16_rn((blockIdx.y == 0 ? w0 : w1)[i]);
}
int cast_weights_bf16(const float* w0, const float* w1, void* out, size_t n, cudaStream_t s) {
  LF_LAUNCH("cast_weights_bf16", s, launch_pdl(cast_weights_bf16_kernel, dim3(div_up((long long)n, 256), 2), dim3(256), 0, s, w0, w1, (__nv_bfloat16*)out, n));
  return check_launch("cast_weights_bf16");
}

// dW[c][d] = sum_s part[s][c][d] in fixed order; db[c] = sum_s dbpart[s][c]
__global__ void reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int splits,
                                     size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
  out[i] = s;
}

// both modalities in one launch: blockIdx.y = modality, partials [m][max_splits][n].  32 columns x 8 split
// groups per CTA: group g adds splits g, g+8, ... in order, then the eight group sums are added in order --
// a fixed summation tree (bit-reproducible) whose dependent-load chains are 8x shorter than one thread
// walking all splits.
__global__ void __launch_bounds__(256) reduce_splits2_kernel(const float* __restrict__ part, float* __restrict__ out0,
                                                             float* __restrict__ out1, int splits, int max_splits, size_t n) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const size_t i = (size_t)blockIdx.x * 32 + tx;
  const float* p = part + (size_t)blockIdx.y * max_splits * n;
  float s = 0.f;
  if (i < n)
    for (int k = ty; k < splits; k += 8) s += p[(size_t)k * n + i];
  __shared__ float sm[8][33];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += sm[g][tx];
    (blockIdx.y == 0 ? out0 : out1)[i] = t;
  }
}

// few splits: one thread per element walks them (fewer, fatter CTAs)
__global__ void reduce_splits2_flat_kernel(const float* __restrict__ part, float* __restrict__ out0, float* __restrict__ out1,
                                           int splits, int max_splits, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = part + (size_t)blockIdx.y * max_splits * n;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += p[(size_t)k * n + i];
  (blockIdx.y == 0 ? out0 : out1)[i] = s;
}

int reduce_splits2(const float* part, float* out0, float* out1, int splits, int max_splits, size_t n, cudaStream_t s) {
  if (splits <= 32) {
    LF_LAUNCH("reduce_dweight", s, (reduce_splits2_flat_kernel<<<dim3(div_up((long long)n, 256), 2), 256, 0, s>>>(part, out0, out1, splits, max_splits, n)));
    return check_launch("reduce_splits2_flat_kernel");
  }
  LF_LAUNCH("reduce_dweight", s, (reduce_splits2_kernel<<<dim3(div_up((long long)n, 32), 2), 256, 0, s>>>(part, out0, out1, splits, max_splits, n)));
  return check_launch("reduce_splits2_kernel");
}

int reduce_splits(const float* part, float* out, int splits, size_t n, cudaStream_t s) {
  LF_LAUNCH("reduce_splits", s, (reduce_splits_kernel<<<div_up((long long)n, 256), 256, 0, s>>>(part, out, splits, n)));
  return check_launch("reduce_splits_kernel");
}

// db_m[c] = sum_b dZ_m[b][c]: column sums, two-stage and fixed-order like dW
__global__ void __launch_bounds__(256) colsum_kernel(const float* dz0, const float* dz1, int B, int C, int ldz, int rows_per_split,
                                                     float* __restrict__ part, int max_splits) {
  const int m = blockIdx.z, split = blockIdx.y;
  const float* __restrict__ dz = m == 0 ? dz0 : dz1;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  const int b0 = split * rows_per_split, b1 = min(B, b0 + rows_per_split);
  float s = 0.f;
  if (c < C)
    for (int b = b0 + ty; b < b1; b += 8) s += dz[(size_t)b * ldz + c];
  __shared__ float sm[8][33];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][tx];
    part[((size_t)m * max_splits + split) * C + c] = t;
  }
}

int colsum(const float* const dz[2], int B, int C, int ldz, float* part, float* const out[2], cudaStream_t s) {
  int splits = div_up(B, 256);
  if (splits > kMaxSplits) splits = kMaxSplits;
  const int rows = div_up(B, splits);
  LF_LAUNCH("colsum", s, (colsum_kernel<<<dim3(div_up(C, 32), splits, 2), 256, 0, s>>>(dz[0], dz[1], B, C, ldz, rows, part, kMaxSplits)));
  int rc = check_launch("colsum_kernel");
  if (rc) return rc;
  for (int m = 0; m < 2; ++m) {
    rc = reduce_splits(part + (size_t)m * kMaxSplits * C, out[m], splits, (size_t)C, s);
    if (rc) return rc;
  }
  return rc;
}

}  // namespace lf
