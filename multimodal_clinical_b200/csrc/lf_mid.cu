// The middle of the step in ONE launch: everything that needs global-batch quantities between the
// forward and the backward pass (SURVEY.md §8 a5-a8, a11 and the loss assembly a4).
//
//   global statistics   = sum over ranks of the per-rank partial statistics, in rank order (bit-identical
//                         on every GPU, so replicated EMA / History state never diverges)
//   EMA update          utils/EMA.py:29-38          OGM-GE coefficients   existing_algos/OGM_GE.py:24-40
//   QMF History update  existing_algos/QMF.py:20-29 (scalar batch-mean CE, last duplicate wins)
//   global min/max, ranking targets, ranking loss + dL/dconf   QMF.py:37-68, 119-141
//   total loss          cremad/joint_model_qmf.py:70 / cremad/joint_model_ogm_ge.py:56
//
// The reference spreads this over ~15 host round trips; the previous version of this library over eight
// small launches and (sharded) three collectives.  Here the inputs arrive as ONE rank-major gathered
// buffer ([rank][stats | idx | conf]: a single all-gather, or the peer-memory push fused in below) and one
// cooperatively launched grid (up to 128 CTAs x 1024 threads, ~2 samples per thread however large the
// GLOBAL batch is) walks the dependent phases with grid-wide barriers in between:
//   [P1 ticket = last writer per index | P0 sum statistics, EMA, coefficients]  -- barrier --
//   [P2+P3 one owner-computes sweep over the History: update + min/max]         -- barrier --
//   [P4 ranking terms, dL/dconf]  -- barrier --  P5 loss
// (three grid barriers; an earlier version used five and a staging array of normalised correctness)
#include <cooperative_groups.h>
#include "lf_common.cuh"
#include "lf_peer.cuh"

namespace cg = cooperative_groups;

namespace lf {

constexpr int kMidMaxCtas = 128;     // cooperative grid: all CTAs co-resident (<= 2 x 148 at 1024 threads)
constexpr int kMidThreads = 1024;

// rank-major gathered batch: sample j of the global batch lives on rank j / Bl
struct Gathered {
  const int64_t* idx; long long idx_stride;     // elements between ranks
  const float* conf; long long conf_stride;     // conf of one rank: [2][Bl]
  int Bl, Bg;
  __device__ __forceinline__ int64_t idx_at(int j) const {
    const int r = j / Bl, i = j - r * Bl;
    return idx[(size_t)r * idx_stride + i];
  }
  __device__ __forceinline__ float conf_at(int m, int j) const {
    const int r = j / Bl, i = j - r * Bl;
    return conf[(size_t)r * conf_stride + (size_t)m * Bl + i];
  }
};

struct MidShared {
  double lo[2], hi[2];
  float s0, q0, q1;
  float l0, l1;          // fp32 batch-mean unimodal CE handed to the History
  float red[32];
  int nan;
};

struct MidParams {
  LfMidArgs a;
  double* minmax;     // [kMidMaxCtas][2][2]
  float* regpart;     // [kMidMaxCtas]
};

// Normalised correctness of batch position j, both modalities (QMF.py:37-42, 51-52): one index load and two
// History gathers (the History is L2-resident: 16 N bytes).  Out-of-range index -> NaN, like a failed lookup.
struct NormPair { double a0, a1; };
__device__ __forceinline__ NormPair norm_at(const LfMidArgs& a, const MidShared& sh, const Gathered& g, int j) {
  const int64_t i = g.idx_at(j);
  const bool ok = (unsigned long long)i < (unsigned long long)a.n_data;
  const double c0 = ok ? a.correctness[i] : (double)NAN;
  const double c1 = ok ? a.correctness[(size_t)a.n_data + i] : (double)NAN;
  NormPair r;
  r.a0 = (c0 - sh.lo[0]) / (sh.hi[0] - sh.lo[0]);
  r.a1 = (c1 - sh.lo[1]) / (sh.hi[1] - sh.lo[1]);
  return r;
}
__device__ __forceinline__ float pair_target(double a, double b) { return (a > b ? 1.f : 0.f) - (a < b ? 1.f : 0.f); }

// ranking terms of the pair (j, j+1) given the normalised correctness of both ends (QMF.py:124-139)
__device__ __forceinline__ void pair_terms(const MidShared& sh, const Gathered& g, int j, const NormPair& cur,
                                           const NormPair& nxt, float* x0, float* x1, float* t0, float* t1) {
  *t0 = pair_target(cur.a0, nxt.a0);
  *t1 = pair_target(cur.a1, nxt.a1);
  const float r = (j + 1 < g.Bg) ? g.conf_at(0, j + 1) : g.conf_at(1, 0);   // flattened roll wraps into modality 1
  *x0 = *t0 * (g.conf_at(0, j) - (r + sh.s0));                               // MarginRankingLoss(x1, x2, -t)
  *x1 = *t1 * (g.conf_at(1, j) - ((r + sh.q0) + sh.q1));
}

__global__ void __launch_bounds__(kMidThreads) mid_kernel(MidParams p) {
  const LfMidArgs& a = p.a;
  cg::grid_group grid = cg::this_grid();                 // grid-wide barriers (cooperative launch)
  const int ncta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int tid = cta * blockDim.x + threadIdx.x, nthr = ncta * blockDim.x;
  const int C = a.classes, len = LF_STATS_HEADER + 2 * C, Bg = a.batch_global, N = a.n_data;
  extern __shared__ double s_stats[];                 // [len] global statistics (each CTA keeps a copy)
  __shared__ MidShared sh;
  __shared__ double slo[2][32], shi[2][32];

  // ---- exchange (sharded runs): push this rank's [stats | idx | conf] to every peer, wait for all peers,
  // then read the local receive area, which has the same rank-major layout an all-gather would produce
  const double* stats_parts = a.stats_parts;
  long long stats_stride = a.stats_stride;
  const int64_t* idx_parts = a.idx_parts; long long idx_stride = a.idx_stride;
  const float* conf_parts = a.conf_parts; long long conf_stride = a.conf_stride;
  long long epoch = 0;
  if (a.use_peer) {
    epoch = a.comm.epoch[0] + 1;
    const int parity = (int)(epoch & 1);
    peer_push(a.comm, a.comm.recv_payload, a.payload_local, (size_t)a.payload_bytes, parity, tid, nthr);
    peer_barrier(a.comm, 0, epoch, grid);
    const char* base = (const char*)a.comm.recv_payload[a.comm.rank] + (size_t)parity * a.n_ranks * a.payload_bytes;
    stats_parts = (const double*)base; stats_stride = a.payload_bytes / 8;
    idx_parts = (const int64_t*)(base + a.off_idx); idx_stride = a.payload_bytes / 8;
    conf_parts = (const float*)(base + a.off_conf); conf_stride = a.payload_bytes / 4;
  }

  Gathered g;
  g.idx = idx_parts; g.idx_stride = idx_stride; g.conf = conf_parts; g.conf_stride = conf_stride;
  g.Bl = a.batch_local; g.Bg = Bg;
  const bool qmf = a.mode == LF_MODE_QMF;
  long long* lw = (long long*)a.last_writer;
  // tickets: host-provided base, or (step_base == 0) the device-resident counter at last_writer[N], which
  // makes the launch replayable from a CUDA graph
  const long long base = qmf ? (a.step_base ? a.step_base : lw[N] + 1) : 0;

  // ---- P1 (issued first: its atomics fly while the statistics are summed): the last duplicate of an index
  // wins (numpy fancy assignment); ticket = position in the batch
  if (qmf)
    for (int j = tid; j < Bg; j += nthr) {
      const int64_t i = g.idx_at(j);
      if ((unsigned long long)i < (unsigned long long)N) atomicMax(&lw[i], base + j);
    }

  // ---- P0: global statistics in rank order -- or, on one GPU, straight from the per-CTA partial rows of the forward
  // kernel (the separate finalize_stats launch folded in): 256 columns x blockDim / 256 row groups, every thread's
  // loads independent, partial sums combined in group order (fixed summation order)
  if (a.stats_rows) {
    __shared__ double s_part[4][256];
    const int ng = (int)blockDim.x / 256, g = threadIdx.x / 256, c0 = threadIdx.x % 256;
    const int nrow = (int)a.n_stats_rows;
    for (int cb = 0; cb < len; cb += 256) {
      const int c = cb + c0;
      double s = 0.0;
      if (c < len) {
        int r = g;
        for (; r + 7 * ng < nrow; r += 8 * ng) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = a.stats_rows[(size_t)(r + k * ng) * len + c];
#pragma unroll
          for (int k = 0; k < 8; ++k) s += (double)v[k];
        }
        for (; r < nrow; r += ng) s += (double)a.stats_rows[(size_t)r * len + c];
      }
      s_part[g][c0] = s;
      __syncthreads();
      if (g == 0 && c < len) {
        double t = s_part[0][c0];
        for (int k = 1; k < ng; ++k) t += s_part[k][c0];
        s_stats[c] = t;
      }
      __syncthreads();
    }
  } else {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < a.n_ranks; ++r) s += stats_parts[(size_t)r * stats_stride + i];
      s_stats[i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.l0 = (float)(s_stats[LF_STAT_CE_X1] / (double)Bg);     // cremad/joint_model_qmf.py:64
    sh.l1 = (float)(s_stats[LF_STAT_CE_X2] / (double)Bg);
    sh.nan = 0;
  }
  if (cta == 0) {
    for (int i = threadIdx.x; i < len; i += blockDim.x)
      if (i != LF_STAT_CNT_X1_CAL && i != LF_STAT_CNT_X2_CAL) a.stats[i] = s_stats[i];
    if (a.update_ema)
      for (int c = threadIdx.x; c < C; c += blockDim.x) {     // utils/EMA.py:33, 38
        const float beta = a.smoothing;
        const float mean1 = (float)(s_stats[LF_STATS_HEADER + c] / (double)Bg);
        const float mean2 = (float)(s_stats[LF_STATS_HEADER + C + c] / (double)Bg);
        const float x1 = mean1 * beta + a.ema_x[c] * (1.0f - beta);
        const float x2 = mean2 * beta + a.ema_x[C + c] * (1.0f - beta);
        a.ema_x[c] = x1; a.ema_x[C + c] = x2;
        const float mu = (x1 + x2) / 2.f;
        a.ema_offset[c] = mu - x1; a.ema_offset[C + c] = mu - x2;
      }
    if (a.coeff_out && threadIdx.x == 0) {                    // existing_algos/OGM_GE.py:24-40
      const float s1 = (float)s_stats[LF_STAT_SCORE_X1], s2 = (float)s_stats[LF_STAT_SCORE_X2];
      const float r1 = s1 / s2, r2 = 1.f / r1;
      float k1 = 1.f, k2 = 1.f;
      if (r1 > 1.f) k1 = 1.f - tanhf(a.alpha * fmaxf(r1, 0.f));
      else k2 = 1.f - tanhf(a.alpha * fmaxf(r2, 0.f));
      a.coeff_out[0] = k1; a.coeff_out[1] = k2;
    }
  }
  __syncthreads();

  if (!qmf) {
    if (cta == 0 && threadIdx.x == 0) {
      if (a.loss_out) a.loss_out[0] = (float)(s_stats[LF_STAT_CE_JOINT] / (double)Bg);
      if (a.use_peer) a.comm.epoch[0] = epoch;
    }
    return;
  }

  grid.sync();                                             // every ticket of this step is in place
  // ---- P2 + P3 in ONE sweep over the History, owner-computes: the thread that scans entry i applies this
  // step's update to it (History.correctness_update, QMF.py:20-29, alpha = 0.1; the winning sample is
  // ticket - base) and folds the resulting value into the min / max over ALL N entries (QMF.py:38-40,
  // NaN-propagating like numpy).  No random access, no second pass, no barrier between update and scan.
  {
    double lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
    bool nan0 = false, nan1 = false;
    const double u0 = 0.1 * (double)sh.l0, u1 = 0.1 * (double)sh.l1;
    for (int i = tid; i < N; i += nthr) {
      const long long t = lw[i];
      double c0 = a.correctness[i], c1 = a.correctness[(size_t)N + i];
      if (t >= base) {
        const int j = (int)(t - base);
        c0 = 0.9 * c0 + u0; c1 = 0.9 * c1 + u1;
        a.correctness[i] = c0; a.correctness[(size_t)N + i] = c1;
        a.confidence[i] = (double)g.conf_at(0, j);
        a.confidence[(size_t)N + i] = (double)g.conf_at(1, j);
      }
      nan0 |= (c0 != c0); nan1 |= (c1 != c1);
      lo0 = fmin(lo0, c0); hi0 = fmax(hi0, c0); lo1 = fmin(lo1, c1); hi1 = fmax(hi1, c1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo0 = fmin(lo0, __shfl_xor_sync(kFull, lo0, o)); hi0 = fmax(hi0, __shfl_xor_sync(kFull, hi0, o));
      lo1 = fmin(lo1, __shfl_xor_sync(kFull, lo1, o)); hi1 = fmax(hi1, __shfl_xor_sync(kFull, hi1, o));
    }
    if (nan0) atomicOr(&sh.nan, 1);
    if (nan1) atomicOr(&sh.nan, 2);
    if (threadIdx.x % 32 == 0) {
      const int w = threadIdx.x / 32;
      slo[0][w] = lo0; shi[0][w] = hi0; slo[1][w] = lo1; shi[1][w] = hi1;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int m = threadIdx.x;
      double lo = slo[m][0], hi = shi[m][0];
      for (int w = 1; w < (int)blockDim.x / 32; ++w) { lo = fmin(lo, slo[m][w]); hi = fmax(hi, shi[m][w]); }
      if (sh.nan & (1 << m)) { lo = NAN; hi = NAN; }
      p.minmax[(cta * 2 + m) * 2 + 0] = lo; p.minmax[(cta * 2 + m) * 2 + 1] = hi;
    }
  }
  grid.sync();                                             // History updated, per-CTA min / max published
  if (threadIdx.x < 64) {
    // warp m combines the per-CTA min / max of modality m: every lane's loads are independent (one memory round trip
    // instead of a chain of ncta), min / max are order-independent, NaN wins
    const int m = threadIdx.x / 32, ln = threadIdx.x % 32;
    double lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (int b = ln; b < ncta; b += 32) {
      const double l = p.minmax[(b * 2 + m) * 2], h = p.minmax[(b * 2 + m) * 2 + 1];
      nan |= (l != l);
      lo = fmin(lo, l); hi = fmax(hi, h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(kFull, lo, o)); hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
    }
    nan = __any_sync(kFull, nan);
    if (nan) { lo = NAN; hi = NAN; }
    if (ln == 0) { sh.lo[m] = lo; sh.hi[m] = hi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const NormPair n0 = norm_at(a, sh, g, 0), n1 = norm_at(a, sh, g, 1), n2 = norm_at(a, sh, g, 2 == Bg ? 0 : 2);
    const float m00 = (float)fabs(n0.a0 - n1.a0), m11 = (float)fabs(n1.a1 - n2.a1);
    const float t00 = pair_target(n0.a0, n1.a0), t01 = pair_target(n1.a0, n2.a0), t11 = pair_target(n1.a1, n2.a1);
    const float z00 = t00 == 0.f ? 1.f : t00, z01 = t01 == 0.f ? 1.f : t01, z11 = t11 == 0.f ? 1.f : t11;
    sh.s0 = m00 / z00;     // rank_margin[0] / rank_target_nonzero, row 0   (QMF.py:134, n = 0)
    sh.q0 = m00 / z01;     // same matrix, row 1 (picked up by n = 1)
    sh.q1 = m11 / z11;     // n = 1: rank_margin[1] / rank_target_nonzero, row 1
  }
  __syncthreads();
  // ---- P4: ranking terms and dL_reg/dconf for this rank's slice.  Each thread
  // normalises the correctness of positions j-1, j, j+1 itself (neighbours hit L1 / L2), so no staging array and
  // no barrier separate the History sweep from the pair terms.
  float reg = 0.f;
  const float invB = 1.f / (float)Bg;
  const int g_begin = a.rank * a.batch_local, g_end = g_begin + a.batch_local;
  for (int j = tid; j < Bg; j += nthr) {
    const int jn = (j + 1 == Bg) ? 0 : j + 1;
    const NormPair cur = norm_at(a, sh, g, j), nxt = norm_at(a, sh, g, jn);
    float x0, x1, t0, t1;
    pair_terms(sh, g, j, cur, nxt, &x0, &x1, &t0, &t1);
    reg += relu_nan(x0) + relu_nan(x1);
    if (j >= g_begin && j < g_end && a.qmf_g) {
      const float u0 = (x0 >= 0.f) ? t0 * invB : 0.f;       // clamp_min backward mask is (x >= 0)
      const float u1 = (x1 >= 0.f) ? t1 * invB : 0.f;
      const int jp = (j == 0) ? Bg - 1 : j - 1;
      const NormPair prv = norm_at(a, sh, g, jp);
      float px0, px1, pt0, pt1;
      pair_terms(sh, g, jp, prv, cur, &px0, &px1, &pt0, &pt1);
      const float v = -(((px0 >= 0.f) ? pt0 * invB : 0.f) + ((px1 >= 0.f) ? pt1 * invB : 0.f));
      a.qmf_g[j - g_begin] = u0 + (j >= 1 ? v : 0.f);       // pair jp's rolled operand is conf0[j] for j >= 1
      a.qmf_g[a.batch_local + (j - g_begin)] = u1 + (j == 0 ? v : 0.f);   // ... and conf1[0] for j == 0
    }
  }
  reg = warp_sum(reg);
  if (threadIdx.x % 32 == 0) sh.red[threadIdx.x / 32] = reg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    bool nan = false;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) { s += sh.red[w]; nan |= (sh.red[w] != sh.red[w]); }
    p.regpart[cta] = nan ? NAN : s;
  }
  grid.sync();
  // ---- P5: loss = CE(z_df) + CE(z1) + CE(z2) + L_reg, each a separate fp32 mean like the reference
  if (cta == 0 && threadIdx.x < 32) {
    // lane l adds the partials of CTAs l, l + 32, ... in order, then a fixed xor butterfly (bit-reproducible)
    double rs = 0.0;
    for (int b = threadIdx.x; b < ncta; b += 32) rs += (double)p.regpart[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(kFull, rs, o);
    if (threadIdx.x == 0) {
    a.stats[LF_STAT_REG_SUM] = rs;
    if (!a.step_base) lw[N] = base - 1 + Bg;
    if (a.use_peer) a.comm.epoch[0] = epoch;
    if (a.loss_out) {
      const double inv = 1.0 / (double)Bg;
      // loss-term ablations drop a term the way the reference does (cremad/joint_model_qmf_ablate_Ljoint.py:68-70)
      const float uni = (a.loss_terms & LF_LOSS_NO_UNI) ? 0.f : (float)(s_stats[LF_STAT_CE_X1] * inv) + (float)(s_stats[LF_STAT_CE_X2] * inv);
      const float joint = (a.loss_terms & LF_LOSS_NO_JOINT) ? 0.f : (float)(s_stats[LF_STAT_CE_JOINT] * inv);
      a.loss_out[0] = (joint + uni) + (float)(rs * inv);
    }
    }
  }
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_mid_workspace_bytes(int32_t batch_global) {
  (void)batch_global;        // per-CTA min/max and ranking-loss partials only (the staging array of an earlier version is gone)
  return 8192;
}

extern "C" int lf_step_mid(const LfMidArgs* a, void* stream) {
  if (a && a->use_peer && (!a->payload_local || a->payload_bytes < 16 || a->payload_bytes % 16 || !a->comm.epoch || !a->comm.error ||
                           a->comm.n_ranks != a->n_ranks || a->comm.rank != a->rank)) {
    set_error("lf_step_mid: bad peer-exchange arguments");
    return LF_ERR_BAD_ARG;
  }
  if (a && a->stats_rows && (a->n_ranks != 1 || a->use_peer || a->n_stats_rows < 1)) {
    set_error("lf_step_mid: stats_rows is a single-GPU input (n_ranks == 1, no peer exchange)");
    return LF_ERR_BAD_ARG;
  }
  if (!a || (!a->stats_parts && !a->use_peer && !a->stats_rows) || !a->stats || a->classes < 1 || a->batch_global < 1 || a->n_ranks < 1 ||
      a->batch_local < 1 || a->batch_local * a->n_ranks != a->batch_global || a->rank < 0 || a->rank >= a->n_ranks) {
    set_error("lf_step_mid: bad argument");
    return LF_ERR_BAD_ARG;
  }
  if (a->update_ema && (!a->ema_x || !a->ema_offset)) { set_error("lf_step_mid: update_ema needs ema_x / ema_offset"); return LF_ERR_BAD_ARG; }
  const bool qmf = a->mode == LF_MODE_QMF;
  if (qmf) {
    if ((!a->use_peer && (!a->idx_parts || !a->conf_parts)) || !a->correctness || !a->confidence || !a->last_writer || !a->workspace ||
        a->n_data < 1 || a->step_base < 0) { set_error("lf_step_mid: QMF mode needs idx/conf/History/workspace"); return LF_ERR_BAD_ARG; }
    if (a->batch_global < 2) {
      set_error("lf_step_mid: batch_global must be >= 2 (reference raises for a batch of one)");
      return LF_ERR_BAD_ARG;
    }
    if (a->workspace_bytes < lf_mid_workspace_bytes(a->batch_global)) { set_error("lf_step_mid: workspace too small"); return LF_ERR_WORKSPACE; }
  } else if (a->mode != LF_MODE_JLOGITS) { set_error("lf_step_mid: bad mode %d", a->mode); return LF_ERR_BAD_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  MidParams p;
  p.a = *a;
  p.minmax = qmf ? (double*)a->workspace : nullptr;
  p.regpart = qmf ? (float*)((char*)a->workspace + 4096) : nullptr;        // minmax: 128 x 4 doubles = 4096 B
  const size_t smem = sizeof(double) * (LF_STATS_HEADER + 2 * (size_t)a->classes);
  cudaLaunchConfig_t cfg = {};
  // QMF: ~2 samples (and ~2 History entries) per thread, so the dependent random accesses of the phases are
  // one or two round trips deep however large the GLOBAL batch is (every rank walks all of it)
  int ncta = 1;
  if (qmf) {
    const long long work = a->batch_global > a->n_data ? a->batch_global : a->n_data;
    ncta = div_up(work, 2 * kMidThreads);
    if (ncta > kMidMaxCtas) ncta = kMidMaxCtas;
    if (ncta < 1) ncta = 1;
  }
  cfg.gridDim = dim3(ncta, 1, 1);
  cfg.blockDim = dim3(qmf ? kMidThreads : 256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaSuccess;
  LF_LAUNCH("step_mid", s, (e = cudaLaunchKernelEx(&cfg, mid_kernel, p)));
  if (e != cudaSuccess) { set_error("lf_step_mid: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  return check_launch("lf_step_mid");
}
