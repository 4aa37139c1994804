// QMF History state and ranking regulariser on the device.
//
// The reference keeps History.correctness/confidence as host numpy fp64 arrays and pays ~10 D2H and
// 4 H2D copies plus four full-N numpy scans per step (existing_algos/QMF.py:20-68).  Here the arrays
// live in HBM as fp64 and the whole update -> min/max -> targets -> ranking loss -> dL/dconf chain is
// five small launches with no host round trip.  Semantics kept bit-for-bit where they are integer /
// comparison based (targets, duplicate handling), IEEE fp64/fp32 otherwise:
//   * gather-then-scatter update: every duplicate of an index writes the same correctness value and the
//     LAST duplicate's confidence wins (numpy fancy assignment).  Implemented with a per-index
//     "last writer" ticket (atomicMax of step_base + position) so exactly one thread updates an index;
//   * min/max over ALL N entries after the update (QMF.py:38-40);
//   * the flattened torch.roll and the row-n pick of the (B,B) broadcast (QMF.py:125-138), closed form
//     in SURVEY.md Appendix A.3.
#include "lf_common.cuh"

namespace lf {

constexpr int kMinMaxBlocks = 128;
constexpr int kRegBlocks = 512;

struct QmfWs {
  double* minmax;   // [2][kMinMaxBlocks][2]
  float* regpart;   // [kRegBlocks]
};

__global__ void qmf_bump_kernel(long long* counter, int Bg) { *counter += Bg; }

__global__ void qmf_mark_kernel(const int64_t* __restrict__ idx, int Bg, int N, long long base,
                                long long* __restrict__ last_writer) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Bg) return;
  if (base == 0) base = last_writer[N] + 1;           // device-resident ticket counter
  const int64_t i = idx[j];
  if ((unsigned long long)i < (unsigned long long)N) atomicMax(&last_writer[i], base + j);
}

__global__ void qmf_update_kernel(const int64_t* __restrict__ idx, const float* __restrict__ conf, int Bg,
                                  int N, long long base, const long long* __restrict__ last_writer,
                                  const double* __restrict__ stats, double* __restrict__ corr,
                                  double* __restrict__ confid, int flags, const float* __restrict__ loss0,
                                  const float* __restrict__ loss1) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Bg) return;
  if (base == 0) base = last_writer[N] + 1;
  const int64_t i = idx[j];
  if ((unsigned long long)i >= (unsigned long long)N) return;   // out-of-range index: ignored (numpy would raise)
  if (last_writer[i] != base + j) return;
  // batch-mean unimodal CE as the fp32 scalar the reference hands to numpy (cremad/joint_model_qmf.py:64-65)
  const double keep = 1.0 - 0.1, alpha = 0.1;           // QMF.py:17,26
  if (flags & LF_QMF_UPDATE_X1) {
    const double l0 = loss0 ? (double)loss0[0] : (double)(float)(stats[LF_STAT_CE_X1] / (double)Bg);
    corr[i] = keep * corr[i] + alpha * l0;
    confid[i] = (double)conf[j];
  }
  if (flags & LF_QMF_UPDATE_X2) {
    const double l1 = loss1 ? (double)loss1[0] : (double)(float)(stats[LF_STAT_CE_X2] / (double)Bg);
    corr[(size_t)N + i] = keep * corr[(size_t)N + i] + alpha * l1;
    confid[(size_t)N + i] = (double)conf[Bg + j];
  }
}

__global__ void qmf_minmax_kernel(const double* __restrict__ corr, int N, double* __restrict__ out) {
  const int m = blockIdx.y;
  const double* c = corr + (size_t)m * N;
  double lo = INFINITY, hi = -INFINITY;
  bool nan = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double v = c[i];
    nan |= (v != v);
    lo = fmin(lo, v); hi = fmax(hi, v);
  }
  __shared__ double slo[8], shi[8];
  __shared__ int snan;
  if (threadIdx.x == 0) snan = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(kFull, lo, o));
    hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
  }
  if (nan) atomicOr(&snan, 1);
  if (threadIdx.x % 32 == 0) { slo[threadIdx.x / 32] = lo; shi[threadIdx.x / 32] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < blockDim.x / 32; ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
    if (snan) { lo = NAN; hi = NAN; }   // numpy min/max propagate NaN
    out[((size_t)m * gridDim.x + blockIdx.x) * 2 + 0] = lo;
    out[((size_t)m * gridDim.x + blockIdx.x) * 2 + 1] = hi;
  }
}

struct RegShared {
  double lo[2], hi[2];
  float s0, q0, q1;
};

__device__ __forceinline__ double norm_corr(double x, double lo, double hi) { return (x - lo) / (hi - lo); }

// ranking target of pair (j, j+1) for modality m  (QMF.py:45-62)
__device__ __forceinline__ float pair_target(const double* __restrict__ corr, int N, int m,
                                             const int64_t* __restrict__ idx, int Bg, int j, double lo,
                                             double hi, float* margin) {
  const int j2 = (j + 1 == Bg) ? 0 : j + 1;
  const int64_t ia = idx[j], ib = idx[j2];
  const double ca = ((unsigned long long)ia < (unsigned long long)N) ? corr[(size_t)m * N + ia] : (double)NAN;
  const double cb = ((unsigned long long)ib < (unsigned long long)N) ? corr[(size_t)m * N + ib] : (double)NAN;
  const double a = norm_corr(ca, lo, hi);
  const double b = norm_corr(cb, lo, hi);
  if (margin) *margin = (float)fabs(a - b);
  return (a > b ? 1.f : 0.f) - (a < b ? 1.f : 0.f);
}

__device__ __forceinline__ void pair_terms(const RegShared& sh, const double* __restrict__ corr, int N,
                                           const int64_t* __restrict__ idx, const float* __restrict__ conf,
                                           int Bg, int j, float* x0, float* x1, float* t0, float* t1) {
  *t0 = pair_target(corr, N, 0, idx, Bg, j, sh.lo[0], sh.hi[0], nullptr);
  *t1 = pair_target(corr, N, 1, idx, Bg, j, sh.lo[1], sh.hi[1], nullptr);
  const float r = (j + 1 < Bg) ? conf[j + 1] : conf[Bg];   // flattened roll: wraps into modality 1, sample 0
  const float c0 = conf[j], c1 = conf[Bg + j];
  *x0 = *t0 * (c0 - (r + sh.s0));                           // MarginRankingLoss(x1, x2, -t), margin 0
  *x1 = *t1 * (c1 - ((r + sh.q0) + sh.q1));
}

__global__ void __launch_bounds__(256) qmf_reg_kernel(LfQmfArgs a, QmfWs ws) {
  __shared__ RegShared sh;
  const int Bg = a.batch_global, N = a.n_data;
  if (threadIdx.x < 2) {
    const int m = threadIdx.x;
    double lo = ws.minmax[((size_t)m * kMinMaxBlocks) * 2], hi = ws.minmax[((size_t)m * kMinMaxBlocks) * 2 + 1];
    bool nan = (lo != lo);
    for (int b = 1; b < kMinMaxBlocks; ++b) {
      const double l = ws.minmax[((size_t)m * kMinMaxBlocks + b) * 2], h = ws.minmax[((size_t)m * kMinMaxBlocks + b) * 2 + 1];
      nan |= (l != l);
      lo = fmin(lo, l); hi = fmax(hi, h);
    }
    if (nan) { lo = NAN; hi = NAN; }
    sh.lo[m] = lo; sh.hi[m] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m00, m11;
    const float t00 = pair_target(a.correctness, N, 0, a.idx, Bg, 0, sh.lo[0], sh.hi[0], &m00);
    const float t01 = pair_target(a.correctness, N, 0, a.idx, Bg, 1, sh.lo[0], sh.hi[0], nullptr);
    const float t11 = pair_target(a.correctness, N, 1, a.idx, Bg, 1, sh.lo[1], sh.hi[1], &m11);
    const float z00 = t00 == 0.f ? 1.f : t00, z01 = t01 == 0.f ? 1.f : t01, z11 = t11 == 0.f ? 1.f : t11;
    sh.s0 = m00 / z00;     // rank_margin[0] / rank_target_nonzero, row 0   (QMF.py:134, n = 0)
    sh.q0 = m00 / z01;     // same matrix, row 1 (picked up by n = 1)
    sh.q1 = m11 / z11;     // n = 1: rank_margin[1] / rank_target_nonzero, row 1
  }
  __syncthreads();

  float reg = 0.f;
  const float invB = 1.f / (float)Bg;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Bg; j += gridDim.x * blockDim.x) {
    float x0, x1, t0, t1;
    pair_terms(sh, a.correctness, N, a.idx, a.conf, Bg, j, &x0, &x1, &t0, &t1);
    reg += relu_nan(x0) + relu_nan(x1);
    if (a.target_out) { a.target_out[j] = t0; a.target_out[Bg + j] = t1; }
    if (j >= a.g_begin && j < a.g_begin + a.g_count) {
      // dL_reg/dconf: own pair directly, previous pair through the rolled operand (Appendix A.4)
      const float u0 = (x0 >= 0.f) ? t0 * invB : 0.f;       // clamp_min backward mask is (x >= 0)
      const float u1 = (x1 >= 0.f) ? t1 * invB : 0.f;
      const int jp = (j == 0) ? Bg - 1 : j - 1;
      float px0, px1, pt0, pt1;
      pair_terms(sh, a.correctness, N, a.idx, a.conf, Bg, jp, &px0, &px1, &pt0, &pt1);
      const float v = -(((px0 >= 0.f) ? pt0 * invB : 0.f) + ((px1 >= 0.f) ? pt1 * invB : 0.f));
      // pair jp's rolled operand is conf0[j] for j >= 1 and conf1[0] for j == 0
      const float g0 = u0 + (j >= 1 ? v : 0.f);
      const float g1 = u1 + (j == 0 ? v : 0.f);
      a.qmf_g[j - a.g_begin] = g0;
      a.qmf_g[a.g_count + (j - a.g_begin)] = g1;
    }
  }
  __shared__ float sred[8];
  reg = warp_sum(reg);
  if (threadIdx.x % 32 == 0) sred[threadIdx.x / 32] = reg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    bool nan = false;
    for (int w = 0; w < 8; ++w) { s += sred[w]; nan |= (sred[w] != sred[w]); }
    ws.regpart[blockIdx.x] = nan ? NAN : s;
  }
}

__global__ void qmf_reg_finalize_kernel(const float* __restrict__ part, int n, double* __restrict__ stats) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += (double)part[i];
  stats[LF_STAT_REG_SUM] = s;
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_qmf_workspace_bytes(int32_t) {
  return align_up(sizeof(double) * 2 * kMinMaxBlocks * 2, 256) + align_up(sizeof(float) * kRegBlocks, 256);
}

extern "C" int lf_qmf_history_step(const LfQmfArgs* a, void* stream) {
  if (!a || !a->idx || !a->conf || !a->correctness || !a->confidence || !a->last_writer || !a->stats ||
      !a->workspace) { set_error("lf_qmf_history_step: null argument"); return LF_ERR_BAD_ARG; }
  if ((a->flags & ~LF_QMF_ALL) || a->flags == 0) { set_error("lf_qmf_history_step: bad flags %d", a->flags); return LF_ERR_BAD_ARG; }
  if ((a->flags & LF_QMF_REG) && !a->qmf_g) { set_error("lf_qmf_history_step: LF_QMF_REG needs qmf_g"); return LF_ERR_BAD_ARG; }
  if (a->batch_global < 1 || ((a->flags & LF_QMF_REG) && a->batch_global < 2)) {
    // the reference raises for B == 1 only in reg_loss / get_target_margin (len() of a 0-d array);
    // History.correctness_update alone works on a batch of one (existing_algos/QMF.py:20-29)
    set_error("lf_qmf_history_step: batch_global must be >= 2 for the ranking loss (reference raises for a batch of one)");
    return LF_ERR_BAD_ARG;
  }
  if (a->n_data < 1 || a->step_base < 0 || a->g_begin < 0 || a->g_count < 0 ||
      a->g_begin + a->g_count > a->batch_global) { set_error("lf_qmf_history_step: bad sizes"); return LF_ERR_BAD_ARG; }
  if (a->workspace_bytes < lf_qmf_workspace_bytes(a->n_data)) { set_error("lf_qmf_history_step: workspace too small"); return LF_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  QmfWs ws;
  ws.minmax = (double*)a->workspace;
  ws.regpart = (float*)((char*)a->workspace + align_up(sizeof(double) * 2 * kMinMaxBlocks * 2, 256));
  const int Bg = a->batch_global, N = a->n_data;
  const int nb = div_up(Bg, 256);
  if (a->flags & (LF_QMF_UPDATE_X1 | LF_QMF_UPDATE_X2)) {
    LF_LAUNCH("qmf_mark", s, (qmf_mark_kernel<<<nb, 256, 0, s>>>(a->idx, Bg, N, (long long)a->step_base, (long long*)a->last_writer)));
    LF_LAUNCH("qmf_update", s, (qmf_update_kernel<<<nb, 256, 0, s>>>(a->idx, a->conf, Bg, N, (long long)a->step_base,
                                        (const long long*)a->last_writer, a->stats, a->correctness, a->confidence,
                                        a->flags, a->loss_uni[0], a->loss_uni[1])));
    if (a->step_base == 0) LF_LAUNCH("qmf_bump", s, (qmf_bump_kernel<<<1, 1, 0, s>>>((long long*)a->last_writer + N, Bg)));
  }
  if (!(a->flags & LF_QMF_REG)) return check_launch("lf_qmf_history_step");
  LF_LAUNCH("qmf_minmax", s, (qmf_minmax_kernel<<<dim3(kMinMaxBlocks, 2), 256, 0, s>>>(a->correctness, N, ws.minmax)));
  const int rb = nb < kRegBlocks ? nb : kRegBlocks;
  LF_LAUNCH("qmf_reg", s, (qmf_reg_kernel<<<rb, 256, 0, s>>>(*a, ws)));
  LF_LAUNCH("qmf_reg_finalize", s, (qmf_reg_finalize_kernel<<<1, 1, 0, s>>>(ws.regpart, rb, a->stats)));
  return check_launch("lf_qmf_history_step");
}
