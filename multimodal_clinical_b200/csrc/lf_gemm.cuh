#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lf {

// C[batch](M x N) = A[batch](M x K) * B[batch](K x N) (+ bias[n]); element addressing is chosen by the
// launcher (k-contiguous or m/n-contiguous operands), see lf_gemm.cu.
struct GemmArgs {
  const float* A[2];
  const float* B[2];
  const float* bias[2];
  float* C[2];
  int M, N, K;
  int lda, ldb, ldc;
  int splits;           // split-K factor (dweight only)
  int k_chunk;          // K range per split
  size_t split_stride;  // elements between split partials in C
  const char* name;     // host-side label for launch accounting
};

int gemm_logits(GemmArgs g, int nbatch, cudaStream_t s);
int gemm_dfeat(GemmArgs g, int nbatch, cudaStream_t s);
int gemm_dweight(GemmArgs g, int nbatch, cudaStream_t s);
int colsum(const float* const dz[2], int B, int C, int ldz, float* part, float* const out[2], cudaStream_t s);
int reduce_splits2(const float* part, float* out0, float* out1, int splits, int max_splits, size_t n, cudaStream_t s);
int reduce_splits(const float* part, float* out, int splits, size_t n, cudaStream_t s);

}  // namespace lf
