// Epoch-end unimodal offset correction on the device (utils/BaseModel.py:168-185):
//   m_out  = mean over all N collected samples of logits (N, M = 2, C)
//   offset = mean_m(m_out) - m_out                                    (M, C)
//   accuracies of argmax(logits_m) and argmax(logits_m + offset_m) against the labels
// Two bandwidth-bound passes over the (N, 2, C) tensor, no host synchronisation:
//   epoch_colsum_kernel  per-CTA column sums (fp32 running sums over a CTA's rows), one partial row per CTA
//   epoch_acc_kernel     every CTA rebuilds the offset from the partial rows in CTA order (fp64, bit-reproducible),
//                        then one warp per sample: both argmaxes of both modalities, integer counts per CTA; the
//                        last CTA to finish adds the counts in order.
#include "lf_common.cuh"

namespace lf {

constexpr int kEpochCtas = 296;      // 2 CTAs per SM
constexpr int kEpochThreads = 256;

__global__ void __launch_bounds__(kEpochThreads) epoch_colsum_kernel(const float* __restrict__ z, long long n, int C,
                                                                     float* __restrict__ part) {
  // columns of the flattened (2C)-wide row are strided over the threads; rows over the CTAs
  const int W = 2 * C;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float s = 0.f;
    for (long long r = blockIdx.x; r < n; r += gridDim.x) s += z[r * W + c];
    part[(size_t)blockIdx.x * W + c] = s;
  }
}

__global__ void __launch_bounds__(kEpochThreads) epoch_acc_kernel(const float* __restrict__ z, const int64_t* __restrict__ label,
                                                                  long long n, int C, const float* __restrict__ part, int nparts,
                                                                  float* __restrict__ offset_out, unsigned* __restrict__ cnt_part,
                                                                  unsigned* __restrict__ ticket, double* __restrict__ acc_out) {
  extern __shared__ float s_off[];                       // [2C]
  const int W = 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int p = 0; p < nparts; ++p) { a += (double)part[(size_t)p * W + c]; b += (double)part[(size_t)p * W + C + c]; }
    const float m1 = (float)(a / (double)n), m2 = (float)(b / (double)n);      // torch.mean(logits, dim=0)
    const float mu = (m1 + m2) / 2.f;                                           // torch.mean(m_out, dim=0)
    s_off[c] = mu - m1; s_off[C + c] = mu - m2;
    if (blockIdx.x == 0) { offset_out[c] = mu - m1; offset_out[C + c] = mu - m2; }
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  unsigned c_u1 = 0, c_u2 = 0, c_c1 = 0, c_c2 = 0;
  for (long long r = (long long)blockIdx.x * nwarp + warp; r < n; r += (long long)gridDim.x * nwarp) {
    const float* row = z + r * W;
    const int y = (int)label[r];
    float b1 = -INFINITY, b2 = -INFINITY, d1 = -INFINITY, d2 = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff, j1 = 0x7fffffff, j2 = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v1 = row[c], v2 = row[C + c];
      const float w1 = v1 + s_off[c], w2 = v2 + s_off[C + c];
      if (v1 > b1) { b1 = v1; i1 = c; }
      if (v2 > b2) { b2 = v2; i2 = c; }
      if (w1 > d1) { d1 = w1; j1 = c; }
      if (w2 > d2) { d2 = w2; j2 = c; }
    }
    warp_argmax(b1, i1); warp_argmax(b2, i2); warp_argmax(d1, j1); warp_argmax(d2, j2);     // first index on ties (torch.argmax)
    if (lane == 0) { c_u1 += (i1 == y); c_u2 += (i2 == y); c_c1 += (j1 == y); c_c2 += (j2 == y); }
  }
  __shared__ unsigned sh[kEpochThreads / 32][4];
  __shared__ bool last;
  if (lane == 0) { sh[warp][0] = c_u1; sh[warp][1] = c_u2; sh[warp][2] = c_c1; sh[warp][3] = c_c2; }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned s = 0;
    for (int w = 0; w < nwarp; ++w) s += sh[w][threadIdx.x];
    cnt_part[blockIdx.x * 4 + threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    unsigned long long s = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&cnt_part[b * 4 + threadIdx.x]);
    acc_out[threadIdx.x] = (double)s / (double)n;          // x1 uncal, x2 uncal, x1 corrected, x2 corrected
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_epoch_workspace_bytes(int32_t classes) {
  if (classes < 1) return 0;
  return 256 + (size_t)kEpochCtas * 4 * sizeof(unsigned) + (size_t)kEpochCtas * 2 * classes * sizeof(float);
}

extern "C" int lf_epoch_offset_correction(const float* logits, const int64_t* labels, int64_t n, int32_t classes,
                                          float* offset_out, double* acc_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!logits || !labels || !offset_out || !acc_out || !workspace || n < 1 || classes < 1) { set_error("lf_epoch_offset_correction: bad argument"); return LF_ERR_BAD_ARG; }
  if (workspace_bytes < lf_epoch_workspace_bytes(classes)) { set_error("lf_epoch_offset_correction: workspace too small"); return LF_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* ticket = (unsigned*)workspace;                       // zero-initialised once by the caller, left zero
  unsigned* cnt = (unsigned*)((char*)workspace + 256);
  float* part = (float*)(cnt + (size_t)kEpochCtas * 4);
  int nb = n < kEpochCtas ? (int)n : kEpochCtas;
  LF_LAUNCH("epoch_colsum", s, (epoch_colsum_kernel<<<nb, kEpochThreads, 0, s>>>(logits, n, classes, part)));
  int rc = check_launch("epoch_colsum_kernel");
  if (rc) return rc;
  int nb2 = div_up(n, kEpochThreads / 32);
  if (nb2 > kEpochCtas) nb2 = kEpochCtas;
  LF_LAUNCH("epoch_acc", s, (epoch_acc_kernel<<<nb2, kEpochThreads, 2 * classes * sizeof(float), s>>>(logits, labels, n, classes, part, nb,
                                                                                                   offset_out, cnt, ticket, acc_out)));
  return check_launch("epoch_acc_kernel");
}
