// Stand-alone pieces of the reference's algorithm API (existing_algos/QMF.py, existing_algos/OGM_GE.py)
// for callers that do not go through the fused step: QMF.df and the ogm_ge score sums.
// Both are one warp per sample with lanes strided over classes (coalesced for any C).
#include "lf_common.cuh"

namespace lf {

constexpr int kScoreBlocks = 296;   // 2 CTAs per SM

// QMF.df (existing_algos/QMF.py:109-117): energy = log(sum(exp z)) -- NOT stabilised, like the reference --
// conf = energy / 10, z_df = z1 * conf1 + z2 * conf2.
__global__ void __launch_bounds__(256) qmf_df_kernel(const float* __restrict__ z1, const float* __restrict__ z2, int B,
                                                     int C, float* __restrict__ zdf, float* __restrict__ conf) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    const float* r1 = z1 + (size_t)b * C;
    const float* r2 = z2 + (size_t)b * C;
    float e1 = 0.f, e2 = 0.f;
    for (int c = lane; c < C; c += 32) { e1 += expf(r1[c]); e2 += expf(r2[c]); }
    const float c1 = logf(warp_sum(e1)) / 10.f, c2 = logf(warp_sum(e2)) / 10.f;
    for (int c = lane; c < C; c += 32) zdf[(size_t)b * C + c] = r1[c] * c1 + r2[c] * c2;
    if (lane == 0) { conf[b] = c1; conf[B + b] = c2; }
  }
}

// score_m = sum_b softmax(z_m)[b, y_b]  (existing_algos/OGM_GE.py:21-22).  Per-CTA partials, then the last
// CTA to finish adds them in index order (deterministic) and resets the ticket for the next call.
__global__ void __launch_bounds__(256) ogm_scores_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                         const int64_t* __restrict__ label, int B, int C,
                                                         double* __restrict__ stats, float* __restrict__ part,
                                                         unsigned int* __restrict__ ticket) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  float s1 = 0.f, s2 = 0.f;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    const float* r1 = z1 + (size_t)b * C;
    const float* r2 = z2 + (size_t)b * C;
    const int y = (int)label[b];
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int c = lane; c < C; c += 32) { m1 = fmaxf(m1, r1[c]); m2 = fmaxf(m2, r2[c]); }
    m1 = warp_max(m1); m2 = warp_max(m2);
    float e1 = 0.f, e2 = 0.f;
    for (int c = lane; c < C; c += 32) { e1 += __expf(r1[c] - m1); e2 += __expf(r2[c] - m2); }
    e1 = warp_sum(e1); e2 = warp_sum(e2);
    if (lane == 0 && y >= 0 && y < C) { s1 += __expf(r1[y] - m1) / e1; s2 += __expf(r2[y] - m2) / e2; }
  }
  __shared__ float sh[8][2];
  __shared__ bool last;
  if (lane == 0) { sh[warp][0] = s1; sh[warp][1] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < nwarp; ++w) { a += sh[w][0]; b += sh[w][1]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned i = 0; i < gridDim.x; ++i) { a += (double)part[2 * i]; b += (double)part[2 * i + 1]; }
    stats[LF_STAT_SCORE_X1] = a; stats[LF_STAT_SCORE_X2] = b;
    *ticket = 0;
  }
}

}  // namespace lf

using namespace lf;

extern "C" int lf_qmf_df(const float* z1, const float* z2, int32_t batch, int32_t classes, float* zdf, float* conf,
                         void* stream) {
  if (!z1 || !z2 || !zdf || !conf || batch < 1 || classes < 1) { set_error("lf_qmf_df: bad argument"); return LF_ERR_BAD_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  int nb = div_up(batch, 8);
  if (nb > 4 * 148) nb = 4 * 148;
  LF_LAUNCH("qmf_df", s, (qmf_df_kernel<<<nb, 256, 0, s>>>(z1, z2, batch, classes, zdf, conf)));
  return check_launch("qmf_df_kernel");
}

extern "C" size_t lf_ogm_scores_workspace_bytes(void) { return 256 + sizeof(float) * 2 * kScoreBlocks; }

extern "C" int lf_ogm_scores(const float* z1, const float* z2, const int64_t* label, int32_t batch, int32_t classes,
                             double* stats, void* workspace, size_t workspace_bytes, void* stream) {
  if (!z1 || !z2 || !label || !stats || !workspace || batch < 1 || classes < 1) { set_error("lf_ogm_scores: bad argument"); return LF_ERR_BAD_ARG; }
  if (workspace_bytes < lf_ogm_scores_workspace_bytes()) { set_error("lf_ogm_scores: workspace too small"); return LF_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  int nb = div_up(batch, 8);
  if (nb > kScoreBlocks) nb = kScoreBlocks;
  LF_LAUNCH("ogm_scores", s, (ogm_scores_kernel<<<nb, 256, 0, s>>>(z1, z2, label, batch, classes, stats,
                                  (float*)((char*)workspace + 256), (unsigned int*)workspace)));
  return check_launch("ogm_scores_kernel");
}

// ---- SGD(momentum, weight decay) for the head parameters (utils/BaseModel.py:275-285: torch.optim.SGD with
// momentum 0.9, weight_decay 1e-4, dampening 0, no nesterov), all head tensors in one launch:
//   d = g + wd * p ;  buf = first ? d : momentum * buf + d ;  p -= lr * buf
namespace lf {
struct SgdTable {
  int count;
  float* p[8];
  const float* g[8];
  float* buf[8];
  long long n[8];
};
__global__ void __launch_bounds__(256) sgd_heads_kernel(SgdTable t, float lr, float momentum, float wd, int first) {
  const int k = blockIdx.y;
  if (k >= t.count) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n[k]; i += (long long)gridDim.x * blockDim.x) {
    const float pv = t.p[k][i];
    const float d = fmaf(wd, pv, t.g[k][i]);
    const float b = first ? d : fmaf(momentum, t.buf[k][i], d);
    t.buf[k][i] = b;
    t.p[k][i] = pv - lr * b;
  }
}
}  // namespace lf

extern "C" int lf_sgd_heads(const LfSgdArgs* a, void* stream) {
  if (!a || a->count < 1 || a->count > 8) { set_error("lf_sgd_heads: bad tensor count"); return LF_ERR_BAD_ARG; }
  lf::SgdTable t;
  t.count = a->count;
  long long maxn = 0;
  for (int k = 0; k < a->count; ++k) {
    if (!a->param[k] || !a->grad[k] || !a->momentum_buf[k] || a->numel[k] < 0) { set_error("lf_sgd_heads: bad tensor %d", k); return LF_ERR_BAD_ARG; }
    t.p[k] = a->param[k]; t.g[k] = a->grad[k]; t.buf[k] = a->momentum_buf[k]; t.n[k] = a->numel[k];
    if (a->numel[k] > maxn) maxn = a->numel[k];
  }
  cudaStream_t s = (cudaStream_t)stream;
  int gx = div_up(maxn, 256 * 4);
  if (gx > 148) gx = 148;
  if (gx < 1) gx = 1;
  LF_LAUNCH("sgd_heads", s, (lf::sgd_heads_kernel<<<dim3(gx, a->count), 256, 0, s>>>(t, a->lr, a->momentum, a->weight_decay, a->first_step)));
  return check_launch("sgd_heads_kernel");
}
