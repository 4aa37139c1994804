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
// The reference spreads this over ~15 host round trips.  Here one grid walks three phases separated by two grid
// barriers:
//   [A  statistics, EMA, coefficients | tickets (atomicMax: last duplicate wins) + History.correctness update by the
//       FIRST thread to reach an entry]                                                      -- barrier 1 --
//   [C  min / max over the whole History | History.confidence of the winners | correctness gathers of this rank's pairs]
//                                                                                            -- barrier 2 --
//   [D  ranking terms, dL/dconf of this rank's slice; the last CTA to finish assembles the loss]
// Heads without a History (mean fusion, the ensemble loss) take mid_light_kernel at the end of this file: one CTA, an
// ordinary launch.
// Sharded runs start with the exchange of [statistics | idx | conf] over NVLink peer memory (lf_peer.cuh: plain stores
// into the peers' receive slots, readers validate against a sentinel: no fence, flag or barrier).
//
// The kernel runs ONCE per step, so every instruction on its path is an instruction-cache miss: measured (LF_MID_TRACE)
// the phases cost ~5x what their memory round trips explain, in proportion to the code they execute, not to the
// data.  Hence the style: rolled loops, one out-of-line copy of each helper, no fp64 division on the hot path, a
// hand-written grid barrier instead of three inlined cooperative-groups ones.
#include <stdlib.h>
#include "lf_common.cuh"
#include "lf_peer.cuh"

namespace lf {

constexpr int kMidMaxCtas = 148;     // every CTA must be resident: one per SM
constexpr int kMidThreads = 512;
constexpr int kMidCols = 8;          // statistics columns a CTA sums per pass (3 CE sums + 2 units, or 4 units)
constexpr int kMidRows = 256;        // statistics rows staged in shared memory per pass

struct Sample { long long idx; float c0, c1; };

// rank-major gathered batch: sample j of the global batch lives on rank j / Bl.  Two representations:
//   arrays   idx (int64) and conf (2, Bl) per rank -- one GPU, or the buffer an NCCL all-gather produced
//   records  the peer-memory receive area of this rank: per rank [len statistics doubles | Bl x {int32 idx, conf0, conf1, 0}],
//            every word pre-filled with the all-ones sentinel and validated by the reader (lf_peer.cuh)
struct Gathered {
  const int64_t* idx; long long idx_stride;     // elements between ranks
  const float* conf; long long conf_stride;     // conf of one rank: [2][Bl]
  const char* rec;                              // record mode: local receive area of the current parity (else null)
  long long slot_bytes, stats_bytes;            // bytes per rank slot / of its statistics part
  int* error;                                   // LfPeerComm.error
  int Bl, Bg, n_ranks;
};

// `gp` points to SHARED memory (one copy per CTA): an out-of-line callee needs an address, and the address of anything
// derived from the kernel parameters would force every thread to keep a local-memory copy of them
static __device__ __noinline__ Sample sample_at(const Gathered* gp, int j) {
  const Gathered& g = *gp;
  int r = 0, i = j;
  if (g.n_ranks > 1) { r = j / g.Bl; i = j - r * g.Bl; }
  Sample s;
  if (g.rec) {
    const uint4* p = reinterpret_cast<const uint4*>(g.rec + (size_t)r * g.slot_bytes + g.stats_bytes) + i;
    uint4 v = ld_volatile_u4(p);
    PeerSpin spin;
    while (v.x == kSentinel32 || v.y == kSentinel32 || v.z == kSentinel32) { spin.wait(g.error); v = ld_volatile_u4(p); }
    s.idx = (long long)(int)v.x; s.c0 = __uint_as_float(v.y); s.c1 = __uint_as_float(v.z);
  } else {
    s.idx = g.idx[(size_t)r * g.idx_stride + i];
    s.c0 = g.conf[(size_t)r * g.conf_stride + i]; s.c1 = g.conf[(size_t)r * g.conf_stride + (size_t)g.Bl + i];
  }
  return s;
}

// existing_algos/OGM_GE.py:24-40
__device__ __forceinline__ void ogm_coeff_write(const LfMidArgs& a, float s1, float s2) {
  const float r1 = s1 / s2, r2 = 1.f / r1;
  float k1 = 1.f, k2 = 1.f;
  if (r1 > 1.f) k1 = 1.f - tanhf(a.alpha * fmaxf(r1, 0.f));
  else k2 = 1.f - tanhf(a.alpha * fmaxf(r2, 0.f));
  a.coeff_out[0] = k1; a.coeff_out[1] = k2;
}

// utils/EMA.py:33, 38 for class c given the global logit sums of both modalities
__device__ __forceinline__ void ema_class(const LfMidArgs& a, int c, double s1, double s2) {
  const int C = a.classes;
  const float beta = a.smoothing;
  const float mean1 = (float)(s1 / (double)a.batch_global), mean2 = (float)(s2 / (double)a.batch_global);
  const float x1 = ema_mix(mean1, a.ema_x[c], beta);
  const float x2 = ema_mix(mean2, a.ema_x[C + c], beta);
  a.ema_x[c] = x1; a.ema_x[C + c] = x2;
  const float mu = (x1 + x2) / 2.f;
  a.ema_offset[c] = mu - x1; a.ema_offset[C + c] = mu - x2;
}

// ---- QMF pair terms.  A pair (j, j+1) of the flattened roll (QMF.py:124-139) needs the indices and confidences of its
// two ends (known before the step touches the History), the UPDATED correctness of both ends (after grid barrier 1)
// and the global min / max (after grid barrier 2).  The three stages are separate so that a thread can have the
// first two in registers by the time the barrier that gates the third one opens.
struct PairIn {
  long long ip, ic, in;        // dataset indices of positions j-1, j, j+1 (cyclic)
  float c0p, c1p, c0c, c1c;    // conf of modality 0 / 1 at j-1 and j
  float rc, rp;                // rolled operand of pair j and of pair j-1
};
struct PairCorr { double p0, p1, c0, c1, n0, n1; };

__device__ __forceinline__ PairIn pair_load(const Gathered* g, int j) {
  const int Bg = g->Bg;
  const int jp = j == 0 ? Bg - 1 : j - 1, jn = j + 1 == Bg ? 0 : j + 1;
  const Sample sp = sample_at(g, jp), sc = sample_at(g, j), sn = sample_at(g, jn);
  PairIn q;
  q.ip = sp.idx; q.ic = sc.idx; q.in = sn.idx;
  q.c0p = sp.c0; q.c1p = sp.c1; q.c0c = sc.c0; q.c1c = sc.c1;
  q.rc = (j + 1 < Bg) ? sn.c0 : sn.c1;      // flattened roll wraps into modality 1 (jn == 0 there: conf1[0])
  q.rp = (j >= 1) ? sc.c0 : sc.c1;          // pair j-1's rolled operand (pair Bg-1 when j == 0: conf1[0])
  return q;
}
// History gathers (L2-resident: 16 N bytes).  Out-of-range index -> NaN, like a failed lookup.
__device__ __forceinline__ double corr_at(const LfMidArgs& a, int m, long long i) {
  return ((unsigned long long)i < (unsigned long long)a.n_data) ? __ldcg(a.correctness + (size_t)m * a.n_data + i) : (double)NAN;
}
__device__ __forceinline__ PairCorr pair_gather(const LfMidArgs& a, const PairIn& q) {
  PairCorr c;
  c.p0 = corr_at(a, 0, q.ip); c.p1 = corr_at(a, 1, q.ip);
  c.c0 = corr_at(a, 0, q.ic); c.c1 = corr_at(a, 1, q.ic);
  c.n0 = corr_at(a, 0, q.in); c.n1 = corr_at(a, 1, q.in);
  return c;
}
__device__ __forceinline__ float pair_target(double a, double b) { return (a > b ? 1.f : 0.f) - (a < b ? 1.f : 0.f); }

struct MidShared {
  double lo[2], hi[2];
  double ce[3];          // global sums of CE(z1), CE(z2), CE(joint)
  double c012[6];        // updated correctness at batch positions 0, 1, 2 (both modalities)
  float s0, q0, q1;      // rank_margin / rank_target_nonzero rows 0 and 1 of the reference's (B,B) matrices (QMF.py:134)
  int fast;              // hi > lo, both finite, for both modalities: targets from the raw correctness values
  float red[kMidThreads / 32];
  int nan;
  int last;
};

// Ranking target of a pair from the correctness of its two ends.  The reference compares the NORMALISED values
// (c - lo) / (hi - lo) (QMF.py:37-42, 51-52, 59-60); for finite hi > lo that map is monotone, so the comparison of the raw
// values gives the same sign -- without a double-precision division per operand, which is most of this kernel's code.
// Degenerate ranges (hi == lo -> 0/0, NaN / inf anywhere) take the literal form.
__device__ __forceinline__ float target_of(const MidShared& sh, int m, double a, double b) {
  if (sh.fast) return pair_target(a, b);
  const double d = sh.hi[m] - sh.lo[m];
  return pair_target((a - sh.lo[m]) / d, (b - sh.lo[m]) / d);
}

// ranking terms of pairs j and j-1 -> relu sum of pair j and dL_reg/dconf of position j
static __device__ __noinline__ float pair_finish(const MidShared& sh, const PairIn& q, const PairCorr& c, int j, float invB,
                                                 float* g0, float* g1) {
  const float s0 = sh.s0, q0 = sh.q0, q1 = sh.q1;
  const float t0 = target_of(sh, 0, c.c0, c.n0), t1 = target_of(sh, 1, c.c1, c.n1);
  const float x0 = t0 * (q.c0c - (q.rc + s0));                    // MarginRankingLoss(x1, x2, -t)
  const float x1 = t1 * (q.c1c - ((q.rc + q0) + q1));
  const float pt0 = target_of(sh, 0, c.p0, c.c0), pt1 = target_of(sh, 1, c.p1, c.c1);
  const float px0 = pt0 * (q.c0p - (q.rp + s0));
  const float px1 = pt1 * (q.c1p - ((q.rp + q0) + q1));
  const float u0 = (x0 >= 0.f) ? t0 * invB : 0.f;                 // clamp_min backward mask is (x >= 0)
  const float u1 = (x1 >= 0.f) ? t1 * invB : 0.f;
  const float v = -(((px0 >= 0.f) ? pt0 * invB : 0.f) + ((px1 >= 0.f) ? pt1 * invB : 0.f));
  *g0 = u0 + (j >= 1 ? v : 0.f);        // pair j-1's rolled operand is conf0[j] for j >= 1 ...
  *g1 = u1 + (j == 0 ? v : 0.f);        // ... and conf1[0] for j == 0
  return relu_nan(x0) + relu_nan(x1);
}

// Grid-wide barrier on a monotone counter (every CTA resident: cooperative launch).  `target` = arrivals that complete it.
static __device__ __noinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

struct MidParams {
  LfMidArgs a;
  double* minmax;     // [kMidMaxCtas][2][2]
  float* regpart;     // [kMidMaxCtas]
  unsigned* done;     // [0] CTAs that have published their ranking-loss partial, [1] grid-barrier arrivals (zero between launches)
  unsigned long long* trace;   // LF_MID_TRACE=1: [grid][8] %globaltimer stamps
};

__global__ void __launch_bounds__(kMidThreads, 1) mid_kernel(MidParams p) {
  const LfMidArgs& a = p.a;
  const int ncta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int tid = cta * blockDim.x + threadIdx.x, nthr = ncta * blockDim.x;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = (int)blockDim.x / 32;
  const int C = a.classes, len = LF_STATS_HEADER + 2 * C, Bg = a.batch_global, N = a.n_data;
  __shared__ MidShared sh;
  __shared__ double slo[2][kMidThreads / 32], shi[2][kMidThreads / 32];
  __shared__ double s_rows[kMidRows][kMidCols], s_col[kMidCols];
  __shared__ int s_cols[kMidCols];

  unsigned long long* tr = p.trace ? p.trace + (size_t)cta * 8 : nullptr;
  auto stamp = [&](int k) { if (tr && threadIdx.x == 0) { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x)); tr[k] = x; } };
  if (tr && threadIdx.x == 0) { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x)); tr[7] = x; p.trace[kMidMaxCtas * 8 + cta] = x; }

  // ---- exchange (sharded runs): this rank's statistics and per-sample records {idx, conf0, conf1} are stored into
  // slot [parity][rank] of every rank's receive area (NVLink stores; the local copy too, so readers see one layout).
  // No fence, flag or barrier follows: readers validate the words they need against the sentinel (lf_peer.cuh).
  // (mean fusion, and QMF heads with LF_LOSS_NO_REG -- the ensemble loss -- have no History / ranking part: mid_light_kernel)
  long long epoch = 0;
  const char* rec_base = nullptr;              // local receive area of this epoch's parity
  const long long stats_bytes = ((long long)len * 8 + 15) / 16 * 16;
  if (a.use_peer) {
    epoch = a.comm.epoch[0] + 1;
    const int parity = (int)(epoch & 1);
    const size_t slot = ((size_t)parity * a.n_ranks + a.rank) * (size_t)a.payload_bytes;
    const double* st_src = (const double*)a.payload_local;
    // (with stats_rows the statistics are exchanged as column sums further down, by the CTA that owns the column)
    for (int c = a.stats_rows ? len : tid; c < len; c += nthr) {
      const double v = clean_f64(st_src[c]);
#pragma unroll 1
      for (int r = 0; r < a.n_ranks; ++r) reinterpret_cast<double*>((char*)a.comm.recv_payload[r] + slot)[c] = v;
    }
    {
      const int64_t* isrc = a.payload_idx_src ? a.payload_idx_src : (const int64_t*)((const char*)a.payload_local + a.off_idx);
      const float* csrc = (const float*)((const char*)a.payload_local + a.off_conf);
#pragma unroll 1
      for (int i = tid; i < a.batch_local; i += nthr) {
        const long long ix = isrc[i];
        uint4 v;
        v.x = ((unsigned long long)ix < (unsigned long long)a.n_data) ? (unsigned)ix : 0xFFFFFFFEu;     // out of range: skipped by every reader
        v.y = __float_as_uint(clean_f32(csrc[i])); v.z = __float_as_uint(clean_f32(csrc[a.batch_local + i])); v.w = 0u;
#pragma unroll 1
        for (int r = 0; r < a.n_ranks; ++r) reinterpret_cast<uint4*>((char*)a.comm.recv_payload[r] + slot + stats_bytes)[i] = v;
      }
    }
    rec_base = (const char*)a.comm.recv_payload[a.rank] + (size_t)parity * a.n_ranks * a.payload_bytes;
  }
  stamp(0);
  // statistic c of row r: the forward's per-CTA partial rows (one GPU), the ranks' finished statistics polled from the
  // receive area (peer exchange) or read from the gathered buffer (NCCL)
  const int nrow = a.stats_rows ? (int)a.n_stats_rows : a.n_ranks;
  auto stat_polled = [&](int r, int c) -> double {       // statistic c of rank r from the receive area
    const void* q = rec_base + (size_t)r * a.payload_bytes + (size_t)c * 8;
    unsigned long long v = ld_volatile_u64(q);
    PeerSpin spin;
    while (v == kSentinel64) { spin.wait(a.comm.error); v = ld_volatile_u64(q); }
    return __longlong_as_double((long long)v);
  };
  auto stat_at = [&](int r, int c) -> double {
    if (a.stats_rows) return (double)a.stats_rows[(size_t)r * len + c];
    if (!rec_base) return a.stats_parts[(size_t)r * a.stats_stride + c];
    return stat_polled(r, c);
  };
  // re-arm this epoch's receive slots once everything has been read (grid-strided; local stores)
  auto rearm = [&]() {
    const size_t n16 = (size_t)a.n_ranks * a.payload_bytes / 16;
    uint4* q = reinterpret_cast<uint4*>(const_cast<char*>(rec_base));
    const uint4 ones = make_uint4(kSentinel32, kSentinel32, kSentinel32, kSentinel32);
#pragma unroll 1
    for (size_t i = tid; i < n16; i += nthr) q[i] = ones;
  };

  // =========================================== QMF ===========================================
  __shared__ Gathered sg;
  if (threadIdx.x == 0) {
    sg.idx = a.idx_parts; sg.idx_stride = a.idx_stride; sg.conf = a.conf_parts; sg.conf_stride = a.conf_stride;
    sg.rec = rec_base; sg.slot_bytes = a.payload_bytes; sg.stats_bytes = stats_bytes; sg.error = a.comm.error;
    sg.Bl = a.batch_local; sg.Bg = Bg; sg.n_ranks = a.n_ranks;
    sh.nan = 0;
  }
  __syncthreads();
  const Gathered* g = &sg;
  long long* lw = (long long*)a.last_writer;
  // tickets: host-provided base, or (step_base == 0) the device-resident counter at last_writer[N], which
  // makes the launch replayable from a CUDA graph
  const long long base = a.step_base ? a.step_base : lw[N] + 1;
  const int g_begin = a.rank * a.batch_local;

  // ---- phase A.  Issued first (no dependency on the statistics, so the round trips overlap the column sums): the
  // ticket of this thread's first batch position -- the LAST duplicate of an index wins (numpy fancy assignment,
  // QMF.py:29), and the value the atomic returns tells the FIRST thread to reach an index this step, which applies the
  // correctness update to it below -- and the pair inputs of its first position in this rank's slice.
  long long my_i = -1, my_old = 0;
  float my_c0 = 0.f, my_c1 = 0.f;
  if (tid < Bg) {
    const Sample sm = sample_at(g, tid);
    my_c0 = sm.c0; my_c1 = sm.c1;
    if ((unsigned long long)sm.idx < (unsigned long long)N) { my_i = sm.idx; my_old = atomicMax(&lw[sm.idx], base + tid); }
  }
  PairIn pin;
  if (tid < a.batch_local) pin = pair_load(g, g_begin + tid);
  long long i012[3] = {0, 0, 0};
  if (threadIdx.x == 0) { i012[0] = sample_at(g, 0).idx; i012[1] = sample_at(g, 1).idx; i012[2] = sample_at(g, 2 == Bg ? 0 : 2).idx; }

  // ---- global statistics.  Thread t stages row t of up to kMidCols columns in shared memory (one memory round trip
  // for everything a CTA needs), one warp adds the rows in a fixed order: the same sums on every launch and rank.
  // Every CTA needs the CE sums (History update, loss); the other columns are spread over the CTAs as units:
  // unit u < HEADER = header column u, else class u - HEADER = both modalities' logit sums + the EMA update of that class.
  {
    const int n_units = LF_STATS_HEADER + C, upc = (n_units + ncta - 1) / ncta;
    int u = cta * upc;
    const int u_end = min(n_units, u + upc);
    bool first = true, coeff = a.coeff_out != nullptr && cta == ncta - 1;
#pragma unroll 1
    do {
      // columns of this pass: the three CE sums (first pass), then as many whole units as fit; the OGM-GE score sums
      // (ogm_ge_lreg: QMF loss + OGM-GE coefficients) ride in a last pass of the last CTA
      const int c_first = first ? 3 : 0;
      const int u0 = u;
      int nu = 0;
      const bool score_pass = !first && u >= u_end;
      if (!score_pass) { nu = min(u_end - u, (kMidCols - c_first) / 2); if (nu < 0) nu = 0; u += nu; }
      auto col_of = [&](int k) -> int {                       // column index of slot k of this pass, -1 = unused
        if (score_pass) return k == 0 ? LF_STAT_SCORE_X1 : k == 1 ? LF_STAT_SCORE_X2 : -1;
        if (k < c_first) return k == 0 ? LF_STAT_CE_X1 : k == 1 ? LF_STAT_CE_X2 : LF_STAT_CE_JOINT;
        const int uu = u0 + (k - c_first) / 2;
        if ((k - c_first) / 2 >= nu) return -1;
        if (uu < LF_STATS_HEADER) return ((k - c_first) & 1) ? -1 : uu;
        return uu + (((k - c_first) & 1) ? C : 0);
      };
      if (threadIdx.x < kMidCols) s_cols[threadIdx.x] = col_of((int)threadIdx.x);
      __syncthreads();
      double acc = 0.0;                                       // lanes 0..7 of warp 0: running sum of column `lane`
#pragma unroll 1
      for (int r0 = 0; r0 < nrow; r0 += kMidRows) {
        const int r = r0 + (int)threadIdx.x;
        if (threadIdx.x < kMidRows && r < nrow) {
          double v[kMidCols];                                 // all loads of the row in flight together
#pragma unroll
          for (int k = 0; k < kMidCols; ++k) { const int col = s_cols[k]; v[k] = col >= 0 ? stat_at(r, col) : 0.0; }
#pragma unroll
          for (int k = 0; k < kMidCols; ++k) s_rows[threadIdx.x][k] = v[k];
        }
        __syncthreads();
        if (warp == 0) {
          // lane = part * 8 + column: four interleaved row subsets per column, then a fixed two-step butterfly
          const int k = lane & 7, nr = min(kMidRows, nrow - r0);
          double s = 0.0;
#pragma unroll 1
          for (int rr = lane >> 3; rr < nr; rr += 4) s += s_rows[rr][k];
          s += __shfl_xor_sync(kFull, s, 8);
          s += __shfl_xor_sync(kFull, s, 16);
          acc += s;
        }
        __syncthreads();
      }
      if (warp == 0 && lane < kMidCols) {
        if (a.stats_rows && rec_base && s_cols[lane] >= 0) {
          // sharded step fed with the forward's per-CTA rows: `acc` is this RANK's sum of the column.  The CTA that owns
          // a column (CTA 0 for the CE sums) stores it into every rank's receive slot; everybody who needs the column
          // polls the ranks' values and adds them in rank order
          const int col = s_cols[lane];
          const bool ce = first && lane < 3;
          if (!ce || cta == 0) {
            const size_t slot = ((size_t)(epoch & 1) * a.n_ranks + a.rank) * (size_t)a.payload_bytes;
            const double v = clean_f64(acc);
#pragma unroll 1
            for (int r = 0; r < a.n_ranks; ++r) reinterpret_cast<double*>((char*)a.comm.recv_payload[r] + slot)[col] = v;
          }
          double t = 0.0;
#pragma unroll 1
          for (int r = 0; r < a.n_ranks; ++r) t += stat_polled(r, col);
          acc = t;
        }
        s_col[lane] = acc;
      }
      __syncthreads();
      if (score_pass) {
        if (threadIdx.x == 0) ogm_coeff_write(a, (float)s_col[0], (float)s_col[1]);
        coeff = false;
      } else {
        if (first && threadIdx.x == 0) { sh.ce[0] = s_col[0]; sh.ce[1] = s_col[1]; sh.ce[2] = s_col[2]; }
        if ((int)threadIdx.x < nu) {
          const int uu = u0 + (int)threadIdx.x;
          const double s1 = s_col[c_first + 2 * threadIdx.x], s2 = s_col[c_first + 2 * threadIdx.x + 1];
          if (uu >= LF_STATS_HEADER) {
            const int c = uu - LF_STATS_HEADER;
            a.stats[LF_STATS_HEADER + c] = s1; a.stats[LF_STATS_HEADER + C + c] = s2;
            if (a.update_ema) ema_class(a, c, s1, s2);
          } else if (uu != LF_STAT_CNT_X1_CAL && uu != LF_STAT_CNT_X2_CAL && uu != LF_STAT_REG_SUM) {
            a.stats[uu] = s1;
          }
        }
      }
      first = false;
      __syncthreads();                                        // s_col / s_rows are reused by the next pass
    } while (u < u_end || coeff);
  }
  // ---- History.correctness_update (QMF.py:20-29, alpha = 0.1): every touched entry moves towards the batch-mean
  // unimodal CE (cremad/joint_model_qmf.py:64; one scalar for the whole batch, so duplicates are harmless and the
  // first thread to reach an entry applies the update exactly once)
  const double u0 = 0.1 * (double)(float)(sh.ce[0] / (double)Bg), u1 = 0.1 * (double)(float)(sh.ce[1] / (double)Bg);
  if (my_i >= 0 && my_old < base) {
    const double c0 = a.correctness[my_i], c1 = a.correctness[(size_t)N + my_i];
    a.correctness[my_i] = 0.9 * c0 + u0; a.correctness[(size_t)N + my_i] = 0.9 * c1 + u1;
  }
#pragma unroll 1
  for (int j = tid + nthr; j < Bg; j += nthr) {              // global batches larger than the grid
    const long long i = sample_at(g, j).idx;
    if ((unsigned long long)i < (unsigned long long)N && atomicMax(&lw[i], base + j) < base) {
      const double c0 = a.correctness[i], c1 = a.correctness[(size_t)N + i];
      a.correctness[i] = 0.9 * c0 + u0; a.correctness[(size_t)N + i] = 0.9 * c1 + u1;
    }
  }
  stamp(1);
  grid_barrier(p.done + 1, (unsigned)ncta);                  // every ticket and every correctness update of this step is in place
  stamp(2);

  // ---- phase C: min / max over ALL N entries (QMF.py:38-40, NaN-propagating like numpy), History.confidence of the
  // winning duplicates, and the correctness gathers of this thread's pair (loads first, uses after)
  PairCorr pc;
  {
    const long long tk = my_i >= 0 ? __ldcg(&lw[my_i]) : -1;
    if (tid < a.batch_local) pc = pair_gather(a, pin);
    if (threadIdx.x == 0)
#pragma unroll 1
      for (int k = 0; k < 3; ++k) { sh.c012[2 * k] = corr_at(a, 0, i012[k]); sh.c012[2 * k + 1] = corr_at(a, 1, i012[k]); }
    double lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
    bool nan0 = false, nan1 = false;
#pragma unroll 2
    for (int i = tid; i < N; i += nthr) {
      const double c0 = __ldcg(a.correctness + i), c1 = __ldcg(a.correctness + (size_t)N + i);
      nan0 |= (c0 != c0); nan1 |= (c1 != c1);
      lo0 = fmin(lo0, c0); hi0 = fmax(hi0, c0); lo1 = fmin(lo1, c1); hi1 = fmax(hi1, c1);
    }
    if (my_i >= 0 && tk == base + tid) {                     // History.confidence (QMF.py:29): the last duplicate of an index wins
      a.confidence[my_i] = (double)my_c0;
      a.confidence[(size_t)N + my_i] = (double)my_c1;
    }
#pragma unroll 1
    for (int j = tid + nthr; j < Bg; j += nthr) {
      const Sample sm = sample_at(g, j);
      if ((unsigned long long)sm.idx < (unsigned long long)N && __ldcg(&lw[sm.idx]) == base + j) {
        a.confidence[sm.idx] = (double)sm.c0;
        a.confidence[(size_t)N + sm.idx] = (double)sm.c1;
      }
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      lo0 = fmin(lo0, __shfl_xor_sync(kFull, lo0, o)); hi0 = fmax(hi0, __shfl_xor_sync(kFull, hi0, o));
      lo1 = fmin(lo1, __shfl_xor_sync(kFull, lo1, o)); hi1 = fmax(hi1, __shfl_xor_sync(kFull, hi1, o));
    }
    if (nan0) atomicOr(&sh.nan, 1);
    if (nan1) atomicOr(&sh.nan, 2);
    if (lane == 0) { slo[0][warp] = lo0; shi[0][warp] = hi0; slo[1][warp] = lo1; shi[1][warp] = hi1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int m = threadIdx.x;
      double lo = slo[m][0], hi = shi[m][0];
#pragma unroll 1
      for (int w = 1; w < nwarp; ++w) { lo = fmin(lo, slo[m][w]); hi = fmax(hi, shi[m][w]); }
      if (sh.nan & (1 << m)) { lo = NAN; hi = NAN; }
      p.minmax[(cta * 2 + m) * 2 + 0] = lo; p.minmax[(cta * 2 + m) * 2 + 1] = hi;
    }
  }
  stamp(3);
  grid_barrier(p.done + 1, 2u * (unsigned)ncta);             // per-CTA min / max published
  stamp(4);

  // ---- phase D: global min / max, the margins of rows 0 and 1, then the ranking terms of this rank's slice
  if (threadIdx.x < 64) {
    // warp m combines the per-CTA min / max of modality m (order-independent; NaN wins); all of a lane's loads first
    const int m = threadIdx.x / 32;
    double lo = INFINITY, hi = -INFINITY;
    bool nan = false;
#pragma unroll 5
    for (int b = lane; b < ncta; b += 32) {
      const double l = __ldcg(&p.minmax[(b * 2 + m) * 2]), h = __ldcg(&p.minmax[(b * 2 + m) * 2 + 1]);
      nan |= (l != l);
      lo = fmin(lo, l); hi = fmax(hi, h);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(kFull, lo, o)); hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
    }
    nan = __any_sync(kFull, nan);
    if (nan) { lo = NAN; hi = NAN; }
    if (lane == 0) { sh.lo[m] = lo; sh.hi[m] = hi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // rank_margin / rank_target_nonzero rows 0 and 1 of the reference's (B,B) matrices (QMF.py:134, n = 0 and n = 1):
    // literal normalised values (one thread: the only fp64 divisions of the kernel)
    const double d0 = sh.hi[0] - sh.lo[0], d1 = sh.hi[1] - sh.lo[1];
    const double n0a0 = (sh.c012[0] - sh.lo[0]) / d0, n1a0 = (sh.c012[2] - sh.lo[0]) / d0, n2a0 = (sh.c012[4] - sh.lo[0]) / d0;
    const double n1a1 = (sh.c012[3] - sh.lo[1]) / d1, n2a1 = (sh.c012[5] - sh.lo[1]) / d1;
    const float m00 = (float)fabs(n0a0 - n1a0), m11 = (float)fabs(n1a1 - n2a1);
    const float t00 = pair_target(n0a0, n1a0), t01 = pair_target(n1a0, n2a0), t11 = pair_target(n1a1, n2a1);
    const float z00 = t00 == 0.f ? 1.f : t00, z01 = t01 == 0.f ? 1.f : t01, z11 = t11 == 0.f ? 1.f : t11;
    sh.s0 = m00 / z00; sh.q0 = m00 / z01; sh.q1 = m11 / z11;
    sh.fast = (d0 > 0.0 && d0 < (double)INFINITY && d1 > 0.0 && d1 < (double)INFINITY) ? 1 : 0;
  }
  __syncthreads();
  // with reg_partial_out the ranking-loss sum of the slice is handed to the gradient exchange, otherwise (one GPU, or
  // a sharded forward-only step) the sum runs over the global batch
  float reg = 0.f;
  const float invB = 1.f / (float)Bg;
  if (tid < a.batch_local) {
    float g0, g1;
    reg += pair_finish(sh, pin, pc, g_begin + tid, invB, &g0, &g1);
    if (a.qmf_g) { a.qmf_g[tid] = g0; a.qmf_g[a.batch_local + tid] = g1; }
  }
  const bool walk_global = a.n_ranks > 1 && !a.reg_partial_out;
  const bool late_reads = a.batch_local > nthr || walk_global;
  if (late_reads) {
#pragma unroll 1
    for (int j = walk_global ? tid : g_begin + tid + nthr; j < (walk_global ? Bg : g_begin + a.batch_local); j += nthr) {
      const int jl = j - g_begin;
      const bool mine = jl >= 0 && jl < a.batch_local;
      if (mine && jl < nthr) continue;                        // done above from registers
      const PairIn q = pair_load(g, j);
      const PairCorr c = pair_gather(a, q);
      float g0, g1;
      reg += pair_finish(sh, q, c, j, invB, &g0, &g1);
      if (mine && a.qmf_g) { a.qmf_g[jl] = g0; a.qmf_g[a.batch_local + jl] = g1; }
    }
  }
  if (rec_base) {
    // every record read of phases A and C precedes grid barrier 2; only the loop above reads records after it
    if (late_reads) grid_barrier(p.done + 1, 3u * (unsigned)ncta);
    rearm();
  }
  reg = warp_sum(reg);
  if (lane == 0) sh.red[warp] = reg;
  __syncthreads();
  stamp(5);
  if (threadIdx.x == 0) {
    float s = 0.f;
    bool nan = false;
#pragma unroll 1
    for (int w = 0; w < nwarp; ++w) { s += sh.red[w]; nan |= (sh.red[w] != sh.red[w]); }
    p.regpart[cta] = nan ? NAN : s;
    __threadfence();
    sh.last = (atomicAdd(p.done, 1u) == (unsigned)ncta - 1) ? 1 : 0;
  }
  __syncthreads();
  // ---- the last CTA to publish its partial assembles the loss = CE(z_df) + CE(z1) + CE(z2) + L_reg, each a separate
  // fp32 mean like the reference (no third grid barrier: nobody else waits for it)
  if (sh.last && threadIdx.x < 32) {
    __threadfence();
    // lane l adds the partials of CTAs l, l + 32, ... in order, then a fixed xor butterfly (bit-reproducible)
    double rs = 0.0;
#pragma unroll 1
    for (int b = lane; b < ncta; b += 32) rs += (double)__ldcg(&p.regpart[b]);
    rs = warp_sum(rs);
    if (lane == 0) {
      p.done[0] = 0u; p.done[1] = 0u;
      if (a.reg_partial_out) a.reg_partial_out[0] = (float)rs; else a.stats[LF_STAT_REG_SUM] = rs;
      if (!a.step_base) lw[N] = base - 1 + Bg;
      if (a.use_peer) a.comm.epoch[0] = epoch;
      if (a.loss_out) {
        const double inv = 1.0 / (double)Bg;
        // loss-term ablations drop a term the way the reference does (cremad/joint_model_qmf_ablate_Ljoint.py:68-70)
        const float uni = (a.loss_terms & LF_LOSS_NO_UNI) ? 0.f : (float)(sh.ce[0] * inv) + (float)(sh.ce[1] * inv);
        const float joint = (a.loss_terms & LF_LOSS_NO_JOINT) ? 0.f : (float)(sh.ce[2] * inv);
        // with reg_partial_out the ranking term is added after the gradient exchange (dW kernel tail)
        a.loss_out[0] = (joint + uni) + (a.reg_partial_out ? 0.f : (float)(rs * inv));
      }
    }
  }
  stamp(6);
}

// ---- mean fusion, and the ensemble loss (QMF heads without History / ranking term): global statistics in rank order,
// EMA, OGM-GE coefficients, loss.  One CTA, no grid barrier -> an ordinary (non-cooperative) launch, and 128 bytes of
// shared memory: the step runs this kernel and the calibrated-count pass on a second stream BESIDE the dfeat GEMM
// (LfHeadsArgs.bwd_phase 3), whose CTAs leave ~2.7 KB of every SM's shared memory free.  A thread sums the column(s)
// it needs itself (the class threads sum their partner column a second time) instead of staging all sums.
__global__ void __launch_bounds__(256, 1) mid_light_kernel(MidParams p) {
  const LfMidArgs& a = p.a;
  const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
  const int C = a.classes, len = LF_STATS_HEADER + 2 * C, Bg = a.batch_global;
  __shared__ double s_hdr[LF_STATS_HEADER];
  long long epoch = 0;
  const char* rec_base = nullptr;
  if (a.use_peer) {
    // this rank's finished statistics into slot [parity][rank] of every rank's receive area (lf_peer.cuh: plain stores,
    // the readers validate every word against the sentinel)
    epoch = a.comm.epoch[0] + 1;
    const int parity = (int)(epoch & 1);
    const size_t slot = ((size_t)parity * a.n_ranks + a.rank) * (size_t)a.payload_bytes;
    const double* st_src = (const double*)a.payload_local;
    for (int c = tid; c < len; c += nthr) {
      const double v = clean_f64(st_src[c]);
#pragma unroll 1
      for (int r = 0; r < a.n_ranks; ++r) reinterpret_cast<double*>((char*)a.comm.recv_payload[r] + slot)[c] = v;
    }
    rec_base = (const char*)a.comm.recv_payload[a.rank] + (size_t)parity * a.n_ranks * a.payload_bytes;
  }
  const int nrow = a.stats_rows ? (int)a.n_stats_rows : a.n_ranks;
  auto col_sum = [&](int c) -> double {                // statistic c summed over the rows / ranks in order
    double s = 0.0;
    int r = 0;
    if (a.stats_rows) {
      // per-CTA rows of a fused forward (up to 296 of them): eight loads in flight, added in row order
      for (; r + 8 <= nrow; r += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = a.stats_rows[(size_t)(r + u) * len + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
      }
    }
#pragma unroll 1
    for (; r < nrow; ++r) {
      double v;
      if (a.stats_rows) v = (double)a.stats_rows[(size_t)r * len + c];
      else if (!rec_base) v = a.stats_parts[(size_t)r * a.stats_stride + c];
      else {
        const void* q = rec_base + (size_t)r * a.payload_bytes + (size_t)c * 8;
        unsigned long long w = ld_volatile_u64(q);
        PeerSpin spin;
        while (w == kSentinel64) { spin.wait(a.comm.error); w = ld_volatile_u64(q); }
        v = __longlong_as_double((long long)w);
      }
      s += v;
    }
    return s;
  };
  for (int i = tid; i < len; i += nthr) {
    const double s = col_sum(i);
    if (i < LF_STATS_HEADER) s_hdr[i] = s;
    if (i != LF_STAT_CNT_X1_CAL && i != LF_STAT_CNT_X2_CAL) a.stats[i] = s;
    if (a.update_ema && i >= LF_STATS_HEADER && i < LF_STATS_HEADER + C) ema_class(a, i - LF_STATS_HEADER, s, col_sum(i + C));
  }
  __syncthreads();                                     // every read of the receive slots is behind this barrier
  if (rec_base) {
    const size_t n16 = (size_t)a.n_ranks * a.payload_bytes / 16;
    uint4* q = reinterpret_cast<uint4*>(const_cast<char*>(rec_base));
    const uint4 ones = make_uint4(kSentinel32, kSentinel32, kSentinel32, kSentinel32);
#pragma unroll 1
    for (size_t i = tid; i < n16; i += nthr) q[i] = ones;
  }
  if (tid == 0) {
    if (a.coeff_out) ogm_coeff_write(a, (float)s_hdr[LF_STAT_SCORE_X1], (float)s_hdr[LF_STAT_SCORE_X2]);
    if (a.loss_out) {
      if (a.mode == LF_MODE_QMF) {         // ensemble: CE(z1) + CE(z2) (+ CE(z_df) unless ablated), each a separate fp32 mean
        const double inv = 1.0 / (double)Bg;
        const float uni = (a.loss_terms & LF_LOSS_NO_UNI) ? 0.f : (float)(s_hdr[LF_STAT_CE_X1] * inv) + (float)(s_hdr[LF_STAT_CE_X2] * inv);
        const float joint = (a.loss_terms & LF_LOSS_NO_JOINT) ? 0.f : (float)(s_hdr[LF_STAT_CE_JOINT] * inv);
        a.loss_out[0] = joint + uni;
      } else {
        a.loss_out[0] = (float)(s_hdr[LF_STAT_CE_JOINT] / (double)Bg);
      }
    }
    if (a.use_peer) a.comm.epoch[0] = epoch;
  }
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_mid_workspace_bytes(int32_t batch_global) {
  (void)batch_global;        // per-CTA min/max and ranking-loss partials, two counters
  return 16384;
}

extern "C" int lf_step_mid(const LfMidArgs* a, void* stream) {
  if (a && a->use_peer && a->mode == LF_MODE_QMF && !(a->loss_terms & LF_LOSS_NO_REG) &&
      (a->off_idx < (int64_t)8 * (LF_STATS_HEADER + 2 * a->classes) || a->off_conf < a->off_idx + (int64_t)8 * a->batch_local ||
       a->payload_bytes < ((int64_t)8 * (LF_STATS_HEADER + 2 * a->classes) + 15) / 16 * 16 + (int64_t)16 * a->batch_local)) {
    set_error("lf_step_mid: payload layout [stats | idx | conf] does not fit payload_bytes");
    return LF_ERR_BAD_ARG;
  }
  if (a && a->use_peer && (!a->payload_local || a->payload_bytes < 16 || a->payload_bytes % 16 || !a->comm.epoch || !a->comm.error ||
                           a->comm.n_ranks != a->n_ranks || a->comm.rank != a->rank)) {
    set_error("lf_step_mid: bad peer-exchange arguments");
    return LF_ERR_BAD_ARG;
  }
  if (a && a->stats_rows && ((a->n_ranks != 1 && (!a->use_peer || (a->loss_terms & LF_LOSS_NO_REG))) || a->n_stats_rows < 1 || a->mode != LF_MODE_QMF)) {
    set_error("lf_step_mid: stats_rows (QMF) needs one rank or the peer exchange (the column sums are exchanged in the kernel)");
    return LF_ERR_BAD_ARG;
  }
  if (!a || (!a->stats_parts && !a->use_peer && !a->stats_rows) || !a->stats || a->classes < 1 || a->batch_global < 1 || a->n_ranks < 1 ||
      a->batch_local < 1 || a->batch_local * a->n_ranks != a->batch_global || a->rank < 0 || a->rank >= a->n_ranks) {
    set_error("lf_step_mid: bad argument");
    return LF_ERR_BAD_ARG;
  }
  if (a->update_ema && (!a->ema_x || !a->ema_offset)) { set_error("lf_step_mid: update_ema needs ema_x / ema_offset"); return LF_ERR_BAD_ARG; }
  if (a->loss_terms & ~(LF_LOSS_NO_JOINT | LF_LOSS_NO_UNI | LF_LOSS_NO_REG)) { set_error("lf_step_mid: bad loss_terms %d", a->loss_terms); return LF_ERR_BAD_ARG; }
  const bool qmf = a->mode == LF_MODE_QMF && !(a->loss_terms & LF_LOSS_NO_REG);
  if (qmf) {
    if ((!a->use_peer && (!a->idx_parts || !a->conf_parts)) || !a->correctness || !a->confidence || !a->last_writer || !a->workspace ||
        a->n_data < 1 || a->step_base < 0) { set_error("lf_step_mid: QMF mode needs idx/conf/History/workspace"); return LF_ERR_BAD_ARG; }
    if (a->batch_global < 2) {
      set_error("lf_step_mid: batch_global must be >= 2 (reference raises for a batch of one)");
      return LF_ERR_BAD_ARG;
    }
    if (a->workspace_bytes < lf_mid_workspace_bytes(a->batch_global)) { set_error("lf_step_mid: workspace too small"); return LF_ERR_WORKSPACE; }
  } else if (a->mode != LF_MODE_JLOGITS && a->mode != LF_MODE_QMF) { set_error("lf_step_mid: bad mode %d", a->mode); return LF_ERR_BAD_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  MidParams p;
  p.a = *a;
  p.minmax = qmf ? (double*)a->workspace : nullptr;
  p.regpart = qmf ? (float*)((char*)a->workspace + 8192) : nullptr;         // minmax: <= 256 x 4 doubles = 8192 B
  p.done = qmf ? (unsigned*)((char*)a->workspace + 8192 + 1024) : nullptr;  // regpart: <= 256 floats
  p.trace = nullptr;
  if (!qmf) {
    LF_LAUNCH("step_mid", s, (mid_light_kernel<<<1, 256, 0, s>>>(p)));
    return check_launch("lf_step_mid");
  }
  cudaLaunchConfig_t cfg = {};
  // QMF: one batch position per thread where the grid allows it (the History sweep and larger global batches loop)
  int ncta = 1;
  if (qmf) {
    long long work = a->batch_global;
    if (a->n_data / 4 > work) work = a->n_data / 4;
    ncta = div_up(work, kMidThreads);
    if (ncta > kMidMaxCtas) ncta = kMidMaxCtas;
    if (ncta < 1) ncta = 1;
  }
  cfg.gridDim = dim3(ncta, 1, 1);
  cfg.blockDim = dim3(kMidThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of the grid (its barriers spin)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = getenv("LF_MID_NOCOOP") ? 0 : 1;      // (experiment: cost of the cooperative launch itself)
  static unsigned long long* trace_buf = nullptr;
  static int trace_calls = 0;
  p.trace = nullptr;
  if (qmf && getenv("LF_MID_TRACE")) {
    if (!trace_buf) { cudaMalloc(&trace_buf, kMidMaxCtas * 9 * sizeof(unsigned long long)); cudaMemset(trace_buf, 0, kMidMaxCtas * 9 * sizeof(unsigned long long)); }
    p.trace = trace_buf;
  }
  cudaError_t e = cudaSuccess;
  LF_LAUNCH("step_mid", s, (e = cudaLaunchKernelEx(&cfg, mid_kernel, p)));
  if (e != cudaSuccess) { set_error("lf_step_mid: %s", cudaGetErrorString(e)); return LF_ERR_CUDA; }
  if (p.trace && ++trace_calls == 12) {
    // LF_MID_TRACE=1 (eager launches only): per-phase %globaltimer stamps of the 12th call, relative to the first CTA's entry
    cudaStreamSynchronize(s);
    static unsigned long long h[kMidMaxCtas * 9];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < ncta; ++c) if (h[kMidMaxCtas * 8 + c] < t0) t0 = h[kMidMaxCtas * 8 + c];      // kernel entry
    const char* nm[8] = {"exchanged", "phaseA_done", "barrier1", "phaseC_done", "barrier2", "phaseD_done", "exit", "entry"};
    for (int k = 0; k < 8; ++k) {
      double mn = 1e30, mx = 0, sum = 0;
      for (int c = 0; c < ncta; ++c) { const double v = (double)(h[c * 8 + k] - t0) / 1000.0; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
      fprintf(stderr, "[mid trace] %-12s min %7.2f  avg %7.2f  max %7.2f us  (%d CTAs)\n", nm[k], mn, sum / ncta, mx, ncta);
    }
  }
  return check_launch("lf_step_mid");
}
