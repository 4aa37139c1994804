// extern "C" entry points of the late-fusion step library (include/lf_fusion.h): argument checking,
// workspace carving and kernel sequencing.  No torch types, no global state beyond a thread-local
// error string.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "lf_common.cuh"
#include "lf_gemm.cuh"
#include "lf_rows.cuh"
#include "lf_tc.cuh"

namespace lf {
int cast_weights_bf16(const float* w0, const float* w1, void* out, size_t n, cudaStream_t s);   // lf_gemm.cu
// lf_narrow.cu
int narrow_tile(int C, int D, bool fwd_only);
size_t narrow_dw_floats(int C, int D);
int narrow_run(const LfHeadsArgs* a, int pass, float* partials, float* dbpart, float* calpart, float* dwpart,
               float* rowstat, int nb_total, int* grid_out, cudaStream_t s);
void finalize_forward_stats(const float* partials, int nblocks, int C, double* stats, cudaStream_t s);  // lf_rows.cu
// lf_tc_fwd.cu: logits of both modalities + the QMF row math in one tcgen05 kernel (32 <= C <= 128)
bool tc_fwd_supported(int mode, int B, int D, int C, int ld_z, int ld_f, int elem);
int tc_heads_forward_qmf(const void* const feat[2], const void* const weight[2], const float* const bias[2], int elem, int B, int D,
                         int C, float* const z[2], int ld_z, float* avg, float* zdf, int ld_f, float* conf, float* rowstat,
                         const int64_t* label, float* partials, int nb_total, int* grid_out, double* stats, unsigned* sync,
                         cudaStream_t s);
// lf_tc_bwd.cu: dL/dz formed inside the dfeat GEMM (QMF, bf16, 32 <= C <= 128)
bool tc_bwd_supported(int mode, int precision, int B, int D, int C, int ld_z, int ldz, int need_dfeat);
int tc_backward_qmf(const RowsArgs& a, const void* const w16[2], void* const dfeat[2], int D, int* grid_out, cudaStream_t s);
}

namespace lf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("LF_PDL") != nullptr;     // opt-in: measured no gain under CUDA-graph replay
  return on;
}

// bit 0: features evict_last in the fused forward; 1: avg / z_df streaming stores; 2: dF evict_first stores;
// 3: the dW GEMM's operands evict_first.  LF_NO_L2_HINTS=1 = none, LF_L2_HINTS=<mask> for experiments.
int l2_hints_mask() {
  static const int mask = getenv("LF_NO_L2_HINTS") ? 0 : (getenv("LF_L2_HINTS") ? atoi(getenv("LF_L2_HINTS")) : 15);
  return mask;
}
bool l2_hints_enabled() { return l2_hints_mask() != 0; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LF_ERR_CUDA;
  }
  return LF_OK;
}

// ---- launch accounting + optional per-kernel event timing -------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static thread_local cudaEvent_t g_pending = nullptr;

// events come from a pool refilled in lf_profile_report: cudaEventCreate on the launch path made the profiled
// (eager) steps host-bound, and the idle gaps landed inside the kernels' event brackets
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(const char*, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  g_pending = take_event();
  cudaEventRecord(g_pending, s);
}
void prof_end(const char* name, cudaStream_t s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed) || !g_pending) return;
  ProfRec r; r.name = name; r.e0 = g_pending; g_pending = nullptr;
  r.e1 = take_event();
  cudaEventRecord(r.e1, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}

int dw_splits(int B, int D, int C) {
  // enough CTAs to fill the machine: tiles(C,D) * splits >= ~2 waves, each split >= 256 samples
  const int tiles = div_up(C, C <= 16 ? 16 : 64) * div_up(D, 64);
  int s = div_up(2 * 148, tiles);
  const int by_rows = div_up(B, 256);
  if (s > by_rows) s = by_rows;
  if (s > kMaxSplits) s = kMaxSplits;
  if (s < 1) s = 1;
  return s;
}

HeadsWorkspace carve_heads_workspace(void* base, int B, int D, int C) {
  HeadsWorkspace w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
  const int part_rows = kMaxRowBlocks;
  w.sync = (unsigned*)take(LF_WS_SYNC_BYTES);
  w.row_partials = (float*)take((size_t)part_rows * stat_len(C) * sizeof(float));
  w.dw_partials = (float*)take((size_t)2 * kMaxSplits * C * D * sizeof(float));
  w.db_partials = (float*)take((size_t)part_rows * 2 * C * sizeof(float));
  w.cal_partials = (float*)take((size_t)kMaxRowBlocks * 2 * sizeof(float));
  w.w16 = take((size_t)2 * C * D * 2);
  w.narrow_dw = narrow_tile(C, D, false) > 0 ? (float*)take(narrow_dw_floats(C, D) * sizeof(float)) : nullptr;
  w.total = off + (size_t)B * 4 * sizeof(float) + 256;  // + rowstat
  return w;
}

static float* rowstat_ptr(void* base, int B, int D, int C) {
  HeadsWorkspace w = carve_heads_workspace(base, B, D, C);
  return (float*)((char*)base + (w.total - (size_t)B * 4 * sizeof(float) - 256));
}

// The tensor pipe only pays for wide heads (C = 6/20 is HBM-bound on FMA).
// LF_PREC_FP32 on wide heads takes the same kernels through the 3xTF32 operand split (fp32-grade products, parity class
// 1e-5) when the dL/dlogits scratch has the TMA row pitch; with the legacy dense pitch it stays on the FMA GEMMs.
static bool use_x3(const LfHeadsArgs* a) {
  return a->precision == LF_PREC_FP32 && a->classes >= 32 && a->ld_dlogits != 0 && a->ld_dlogits % 4 == 0 && !getenv("LF_NO_X3");
}
static bool use_tensor_pipe(const LfHeadsArgs* a) { return (a->precision != LF_PREC_FP32 && a->classes >= 32) || use_x3(a); }
static bool is_bf16(const LfHeadsArgs* a) { return a->precision == LF_PREC_BF16; }
// Narrow heads (C <= 32, HBM-bound on FMA): one fused kernel per pass over the features (lf_narrow.cu).
static bool use_narrow(const LfHeadsArgs* a) {
  return !use_tensor_pipe(a) && a->classes <= 32 && narrow_tile(a->classes, a->dim, false) > 0 && !getenv("LF_NO_NARROW");
}
static int narrow_grid(const LfHeadsArgs* a, bool fwd_only) {
  const int S = narrow_tile(a->classes, a->dim, fwd_only);
  int grid = div_up(a->batch, S);
  if (grid > 148) grid = 148;
  const int rpc = div_up(a->batch, grid);
  return div_up(a->batch, rpc);
}

// bf16 copy of head m: the caller's (kept current across optimizer steps) or the per-call cast in the workspace
static const void* heads_w16(const LfHeadsArgs* a, const HeadsWorkspace& w, int m) {
  if (a->weight_bf16[0] && a->weight_bf16[1]) return a->weight_bf16[m];
  return (const char*)w.w16 + (size_t)m * a->classes * a->dim * 2;
}

static int tc_block_n(int n) {            // N tile: <= 256, multiple of 16, balanced over the tiles
  const int tiles = div_up(n, 256);
  return div_up(div_up(n, tiles), 16) * 16;
}

// Split-K plan of the tensor-pipe dW GEMM and whether its tail (reduction of the partials, db, calibrated counts,
// optional gradient all-reduce and SGD step) can run inside the kernel: all CTAs must be co-resident (one wave).
static bool dw_tail_plan(const LfHeadsArgs* a, int* splits_out, int* block_n_out) {
  const int block_n = div_up(tc_block_n(a->dim), 64) * 64;
  // split-K so that (C tiles x D tiles x 2 modalities x splits) covers the 148 SMs about once
  const int tiles = div_up(a->classes, 128) * div_up(a->dim, block_n) * 2;
  int splits = 148 / tiles;                    // floor: one full wave, no second-wave tail
  const int by_rows = div_up(a->batch, 128);
  if (splits > by_rows) splits = by_rows;
  if (splits > kMaxSplits) splits = kMaxSplits;
  if (splits < 1) splits = 1;
  if (splits_out) *splits_out = splits;
  if (block_n_out) *block_n_out = block_n;
  const size_t cd = (size_t)a->classes * a->dim;
  return cd % 4 == 0 && tiles * splits <= 148 && !getenv("LF_NO_DW_TAIL");
}

static int check_heads(const LfHeadsArgs* a, bool backward) {
  if (!a) { set_error("null LfHeadsArgs"); return LF_ERR_BAD_ARG; }
  if (a->batch < 1 || a->batch_global < a->batch || a->dim < 4 || a->dim % 4 || a->classes < 1) {
    set_error("bad sizes: batch=%d batch_global=%d dim=%d (multiple of 4) classes=%d", a->batch,
              a->batch_global, a->dim, a->classes);
    return LF_ERR_BAD_ARG;
  }
  if (a->mode != LF_MODE_JLOGITS && a->mode != LF_MODE_QMF) { set_error("bad mode %d", a->mode); return LF_ERR_BAD_ARG; }
  if (a->precision != LF_PREC_FP32 && a->precision != LF_PREC_TF32 && a->precision != LF_PREC_BF16) { set_error("bad precision %d", a->precision); return LF_ERR_BAD_ARG; }
  if (a->precision == LF_PREC_BF16 && (a->classes < 32 || a->dim % 8 || a->ld_dlogits % 8 || a->ld_dlogits == 0)) {
    set_error("LF_PREC_BF16 needs classes >= 32 and dim / ld_dlogits multiples of 8 (got C=%d D=%d ld=%d)", a->classes, a->dim, a->ld_dlogits);
    return LF_ERR_UNSUPPORTED;
  }
  for (int m = 0; m < 2; ++m)
    if (!a->feat[m] || !a->weight[m] || !a->bias[m] || !a->logits[m] || !a->dweight[m] || !a->dbias[m]) {
      set_error("null per-modality pointer (modality %d)", m);
      return LF_ERR_BAD_ARG;
    }
  if (!a->label || !a->avg_logits || !a->dlogits[0] || !a->stats || !a->workspace) { set_error("null pointer argument"); return LF_ERR_BAD_ARG; }
  if (a->mode == LF_MODE_QMF && (!a->logits_df || !a->conf || !a->dlogits[1])) { set_error("QMF mode needs logits_df, conf, dlogits[1]"); return LF_ERR_BAD_ARG; }
  if (a->need_dfeat && (!a->dfeat[0] || !a->dfeat[1])) { set_error("need_dfeat set but dfeat is null"); return LF_ERR_BAD_ARG; }
  if (backward && !a->ema_offset) { set_error("backward needs ema_offset"); return LF_ERR_BAD_ARG; }
  if (backward && a->mode == LF_MODE_QMF && !a->qmf_g) { set_error("QMF backward needs qmf_g"); return LF_ERR_BAD_ARG; }
  if (a->ld_logits != 0 && (a->ld_logits < a->classes || (a->ld_logits != a->classes && (!use_tensor_pipe(a) )))) {
    set_error("ld_logits %d: must be >= classes, and a padded pitch is only supported on the tensor-pipe path", a->ld_logits);
    return LF_ERR_BAD_ARG;
  }
  if (a->loss_terms & ~(LF_LOSS_NO_JOINT | LF_LOSS_NO_UNI | LF_LOSS_NO_REG)) { set_error("bad loss_terms %d", a->loss_terms); return LF_ERR_BAD_ARG; }
  if (a->ld_fused != 0 && a->ld_fused != a->classes && !use_tensor_pipe(a)) { set_error("ld_fused: a padded pitch is only supported on the tensor-pipe path"); return LF_ERR_BAD_ARG; }
  if (a->ld_fused != 0 && a->ld_fused < a->classes) { set_error("ld_fused %d < classes %d", a->ld_fused, a->classes); return LF_ERR_BAD_ARG; }
  if (a->ld_dlogits != 0 && a->ld_dlogits < a->classes) { set_error("ld_dlogits %d < classes %d", a->ld_dlogits, a->classes); return LF_ERR_BAD_ARG; }
  if (use_tensor_pipe(a) && (a->ld_dlogits % 4 != 0 || a->ld_dlogits == 0)) {
    set_error("LF_PREC_TF32 needs ld_dlogits to be a non-zero multiple of 4 (TMA row pitch), got %d", a->ld_dlogits);
    return LF_ERR_BAD_ARG;
  }
  if (a->workspace_bytes < lf_workspace_bytes(a->batch, a->dim, a->classes)) {
    set_error("workspace too small: %zu < %zu", a->workspace_bytes, lf_workspace_bytes(a->batch, a->dim, a->classes));
    return LF_ERR_WORKSPACE;
  }
  return LF_OK;
}

static RowsArgs rows_args(const LfHeadsArgs* a, const HeadsWorkspace& w) {
  RowsArgs r;
  r.z[0] = a->logits[0]; r.z[1] = a->logits[1];
  r.avg = a->avg_logits; r.zdf = a->logits_df; r.conf = a->conf;
  r.rowstat = rowstat_ptr(a->workspace, a->batch, a->dim, a->classes);
  r.dz[0] = a->dlogits[0]; r.dz[1] = a->dlogits[1];
  r.label = a->label; r.qmf_g = a->qmf_g; r.ema_off = a->ema_offset;
  r.partials = w.row_partials; r.stats = a->stats;
  r.dbpart = w.db_partials; r.calpart = w.cal_partials;
  r.B = a->batch; r.B_global = a->batch_global; r.C = a->classes;
  r.ld_z = a->ld_logits > 0 ? a->ld_logits : a->classes;
  r.ldz = a->ld_dlogits > 0 ? a->ld_dlogits : a->classes;
  r.ld_f = a->ld_fused > 0 ? a->ld_fused : a->classes;
  r.w_joint = (a->loss_terms & LF_LOSS_NO_JOINT) ? 0.f : 1.f;
  r.w_uni = (a->loss_terms & LF_LOSS_NO_UNI) ? 0.f : 1.f;
  r.dz_bf16 = a->precision == LF_PREC_BF16;
  r.nb_total = row_blocks(a->batch);
  return r;
}

}  // namespace lf

using namespace lf;

extern "C" const char* lf_last_error(void) { return g_err; }
extern "C" int32_t lf_abi_version(void) { return LF_ABI_VERSION; }
extern "C" int64_t lf_launch_count(void) { return g_launches.load(); }
extern "C" void lf_profile_enable(int32_t on) {
  if (on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    while (g_event_pool.size() < 2048) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; g_event_pool.push_back(e); }
  }
  g_prof_on.store(on ? 1 : 0);
}

// Synchronises the device, then writes "name count total_ms\n" lines for every kernel timed since the
// last report and clears the records.  Returns the number of bytes written (excluding the NUL).
extern "C" int32_t lf_profile_report(char* buf, int32_t buf_bytes) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> agg;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        auto& a = agg[r.name];
        a.first += 1; a.second += ms;
      }
      g_event_pool.push_back(r.e0); g_event_pool.push_back(r.e1);
    }
    g_prof.clear();
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (!buf || buf_bytes <= 0) return 0;
  const size_t n = out.size() < (size_t)buf_bytes - 1 ? out.size() : (size_t)buf_bytes - 1;
  memcpy(buf, out.data(), n);
  buf[n] = 0;
  return (int32_t)n;
}

extern "C" size_t lf_workspace_bytes(int32_t batch, int32_t dim, int32_t classes) {
  if (batch < 1 || dim < 1 || classes < 1) return 0;
  return carve_heads_workspace(nullptr, batch, dim, classes).total;
}

extern "C" int lf_heads_forward(const LfHeadsArgs* a, void* stream) {
  int rc = check_heads(a, false);
  if (rc) return rc;
  if (a->stats_rows_out) { a->stats_rows_out[0] = 0; a->stats_rows_out[1] = 0; }
  cudaStream_t s = (cudaStream_t)stream;
  HeadsWorkspace w = carve_heads_workspace(a->workspace, a->batch, a->dim, a->classes);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  for (int m = 0; m < 2; ++m) { g.A[m] = a->feat[m]; g.B[m] = a->weight[m]; g.bias[m] = a->bias[m]; g.C[m] = a->logits[m]; }
  g.M = a->batch; g.N = a->classes; g.K = a->dim;
  g.lda = a->dim; g.ldb = a->dim; g.ldc = a->classes;
  if (use_narrow(a)) {
    // JLOGITS: forward AND backward of the heads in this one pass (unless the caller only wants the forward);
    // QMF: forward pass, the backward pass follows the mid-step exchange
    const int pass = a->mode == LF_MODE_QMF ? 1 : 0;
    LfHeadsArgs b = *a;
    if (a->mode == LF_MODE_JLOGITS && a->fwd_only) b.need_dfeat = 0;
    int grid = 0;
    rc = narrow_run(&b, (a->mode == LF_MODE_JLOGITS && a->fwd_only) ? 1 : pass, w.row_partials, w.db_partials, w.cal_partials,
                    w.narrow_dw, rowstat_ptr(a->workspace, a->batch, a->dim, a->classes), 0, &grid, s);
    if (rc) return rc;
    finalize_forward_stats(w.row_partials, grid, a->classes, a->stats, s);
    return check_launch("finalize_stats");
  }
  if (use_tensor_pipe(a)) {
    TcGemmDesc d;
    d.nbatch = 2;
    const size_t cd = (size_t)a->classes * a->dim;
    if (is_bf16(a)) {
      if (!(a->weight_bf16[0] && a->weight_bf16[1])) {
        rc = cast_weights_bf16(a->weight[0], a->weight[1], w.w16, cd, s);     // what autocast does for nn.Linear every step
        if (rc) return rc;
      }
      d.elem = 2;
    }
    for (int m = 0; m < 2; ++m) {
      d.A[m] = a->feat[m];
      d.B[m] = is_bf16(a) ? heads_w16(a, w, m) : (const void*)a->weight[m];
      d.bias[m] = a->bias[m]; d.out[m] = a->logits[m];
    }
    {
      const int ldz = a->ld_logits > 0 ? a->ld_logits : a->classes, ldf = a->ld_fused > 0 ? a->ld_fused : a->classes;
      if (!use_x3(a) && tc_fwd_supported(a->mode, a->batch, a->dim, a->classes, ldz, ldf, d.elem)) {
        // QMF, 32 <= C <= 128: the row math runs in the GEMM's epilogue (one thread per sample straight from TMEM)
        int grid = 0;
        const void* fp[2] = {a->feat[0], a->feat[1]};
        float* zp[2] = {a->logits[0], a->logits[1]};
        rc = tc_heads_forward_qmf(fp, d.B, a->bias, d.elem, a->batch, a->dim, a->classes, zp, ldz, a->avg_logits, a->logits_df, ldf,
                                  a->conf, rowstat_ptr(a->workspace, a->batch, a->dim, a->classes), a->label, w.row_partials, 0,
                                  &grid, a->stats_rows_out ? nullptr : a->stats, w.sync, s);
        if (rc) return rc;
        // single-GPU steps hand the per-CTA rows to lf_step_mid (LfMidArgs.stats_rows), which sums them with its
        // whole grid; the exchange of a sharded step needs finished statistics, so they are summed here
        // (a sharded step exchanges finished statistics: there the last CTA of the forward kernel to finish sums them)
        if (a->stats_rows_out) { a->stats_rows_out[0] = (uint64_t)(uintptr_t)w.row_partials; a->stats_rows_out[1] = (uint64_t)grid; }
        return LF_OK;
      }
    }
    d.M = a->batch; d.N = a->classes; d.K = a->dim;
    d.lda = a->dim; d.ldb = a->dim; d.ld_out = a->ld_logits > 0 ? a->ld_logits : a->classes;
    d.a_mn_major = 0; d.b_mn_major = 0; d.block_n = tc_block_n(a->classes);
    if (a->classes > d.block_n) d.block_n = div_up(d.block_n, 32) * 32;     // several N tiles: whole 128-byte store chunks
    d.splits = 1; d.split_stride = 0; d.balance_m = 1; d.name = "tc_logits"; d.x3 = use_x3(a);
    rc = tc_gemm(d, s);
  } else {
    rc = gemm_logits(g, 2, s);
  }
  if (rc) return rc;
  return rows_forward(rows_args(a, w), a->mode, s);
}

extern "C" int lf_heads_backward(const LfHeadsArgs* a, void* stream) {
  int rc = check_heads(a, true);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  HeadsWorkspace w = carve_heads_workspace(a->workspace, a->batch, a->dim, a->classes);
  const int phase = a->bwd_phase;
  if (phase < 0 || phase > 4) { set_error("bad bwd_phase %d", phase); return LF_ERR_BAD_ARG; }
  if (phase >= 3 && !lf_heads_backward_splits_rows(a)) {
    set_error("bwd_phase %d: this shape forms dL/dz inside a fused kernel (ask lf_heads_backward_splits_rows first)", phase);
    return LF_ERR_UNSUPPORTED;
  }
  if (use_narrow(a)) {
    if (phase == 2) return LF_OK;                 // dfeat is produced inside the fused narrow kernel
    const size_t cdn = (size_t)a->classes * a->dim;
    int grid = narrow_grid(a, false), nb_cal;
    if (a->mode == LF_MODE_QMF) {
      rc = narrow_run(a, 2, w.row_partials, w.db_partials, w.cal_partials, w.narrow_dw,
                      rowstat_ptr(a->workspace, a->batch, a->dim, a->classes), 0, &grid, s);
      nb_cal = grid;
    } else {
      // dz, dfeat, dW and db partials were produced by the fused forward pass; only the calibrated counts
      // (which need this step's EMA offsets) are left
      if (a->fwd_only) { set_error("lf_heads_backward called for a step whose forward ran with fwd_only"); return LF_ERR_BAD_ARG; }
      rc = rows_backward(rows_args(a, w), a->mode, s);
      nb_cal = row_blocks(a->batch);
    }
    if (rc) return rc;
    rc = reduce_splits2(w.narrow_dw, a->dweight[0], a->dweight[1], grid, 148, cdn, s);
    if (rc) return rc;
    return finalize_db_cal(w.db_partials, grid, a->classes, w.cal_partials, nb_cal, a->dbias[0], a->dbias[1], a->stats, s);
  }
  const int ldz = a->ld_dlogits > 0 ? a->ld_dlogits : a->classes;
  const bool tc = use_tensor_pipe(a);
  // QMF / bf16 / C <= 128: dz is formed inside the dfeat GEMM (lf_tc_bwd.cu); dz, db partials and the calibrated counts
  // come out of the same kernel, so phase 1 of a sharded run already produces dfeat and phase 2 has nothing left to do
  const bool fused_bwd = tc && tc_bwd_supported(a->mode, a->precision, a->batch, a->dim, a->classes,
                                                a->ld_logits > 0 ? a->ld_logits : a->classes, ldz, a->need_dfeat);
  int nb_parts = row_blocks(a->batch);           // per-CTA partial rows (db, calibrated counts) of the kernel that made dz
  if (fused_bwd) {
    if (phase == 2) return LF_OK;
    const void* w16[2] = {heads_w16(a, w, 0), heads_w16(a, w, 1)};
    void* df[2] = {a->dfeat[0], a->dfeat[1]};
    rc = tc_backward_qmf(rows_args(a, w), w16, df, a->dim, &nb_parts, s);
    if (rc) return rc;
  } else if (phase != 2 && phase != 4) {
    rc = rows_backward(rows_args(a, w), a->mode, s);
    if (rc) return rc;
  }
  if (phase == 3) return LF_OK;
  const float* dz[2] = {a->dlogits[0], a->mode == LF_MODE_QMF ? a->dlogits[1] : a->dlogits[0]};
  GemmArgs g;
  if (a->need_dfeat && phase != 1 && phase != 4 && !fused_bwd) {
    memset(&g, 0, sizeof(g));
    for (int m = 0; m < 2; ++m) { g.A[m] = dz[m]; g.B[m] = a->weight[m]; g.bias[m] = nullptr; g.C[m] = a->dfeat[m]; }
    g.M = a->batch; g.N = a->dim; g.K = a->classes;
    g.lda = ldz; g.ldb = a->dim; g.ldc = a->dim;
    if (tc) {
      TcGemmDesc d;
      d.nbatch = 2;
      if (is_bf16(a)) { d.elem = 2; d.out_elem = 2; }
      for (int m = 0; m < 2; ++m) {
        d.A[m] = dz[m];
        d.B[m] = is_bf16(a) ? heads_w16(a, w, m) : (const void*)a->weight[m];
        d.bias[m] = nullptr; d.out[m] = a->dfeat[m];
      }
      d.M = a->batch; d.N = a->dim; d.K = a->classes;
      d.lda = ldz; d.ldb = a->dim; d.ld_out = a->dim;
      d.a_mn_major = 0; d.b_mn_major = 1; d.block_n = div_up(tc_block_n(a->dim), 64) * 64;
      d.splits = 1; d.split_stride = 0; d.balance_m = 0; d.name = "tc_dfeat"; d.x3 = use_x3(a);
      // sharded runs (phase 2) overlap dfeat with the peer all-reduce on a side stream: keep the CTA small enough
      // (192 threads) for the all-reduce CTAs to be co-resident, or the exchange waits for dfeat to drain
      d.max_epi_halves = phase == 2 ? 1 : 2;
      rc = tc_gemm(d, s);
    } else {
      rc = gemm_dfeat(g, 2, s);
    }
    if (rc) return rc;
  }
  if (phase == 2) return LF_OK;
  // dW_m = dZ_m^T F_m : split-K over the batch, partials reduced in fixed order
  int splits = dw_splits(a->batch, a->dim, a->classes);
  const size_t cd = (size_t)a->classes * a->dim;
  memset(&g, 0, sizeof(g));
  for (int m = 0; m < 2; ++m) { g.A[m] = dz[m]; g.B[m] = a->feat[m]; g.bias[m] = nullptr; g.C[m] = w.dw_partials + (size_t)m * kMaxSplits * cd; }
  g.M = a->classes; g.N = a->dim; g.K = a->batch;
  g.lda = ldz; g.ldb = a->dim; g.ldc = a->dim;
  if (tc) {
    TcGemmDesc d;
    d.nbatch = 2;
    if (is_bf16(a)) d.elem = 2;
    for (int m = 0; m < 2; ++m) { d.A[m] = dz[m]; d.B[m] = a->feat[m]; d.bias[m] = nullptr; d.out[m] = w.dw_partials + (size_t)m * kMaxSplits * cd; }
    d.M = a->classes; d.N = a->dim; d.K = a->batch;
    d.lda = ldz; d.ldb = a->dim; d.ld_out = a->dim;
    d.a_mn_major = 1; d.b_mn_major = 1;
    // the reduction of the split-K partials, db, the calibrated counts (and the optional all-reduce / SGD step) run in the kernel's tail
    const bool tail = dw_tail_plan(a, &splits, &d.block_n);
    d.splits = splits; d.split_stride = (long long)cd; d.balance_m = 0; d.name = "tc_dweight"; d.x3 = use_x3(a);
    d.l2_last_use = (l2_hints_mask() & 8) ? 3 : 0;       // F and dz are read for the last time in the step
    const bool peer = tail && a->grad_comm && a->batch_global != a->batch;
    if (a->grad_comm && a->batch_global != a->batch && !peer) {
      set_error("grad_comm given but this shape cannot fuse the all-reduce (ask lf_heads_backward_fuses_allreduce first)");
      return LF_ERR_UNSUPPORTED;
    }
    if (a->sgd && (!tail || (a->batch_global != a->batch && !peer))) {
      set_error("fused SGD needs the tensor-pipe dW tail, and on a sharded step the all-reduce fused into it (grad_comm)");
      return LF_ERR_UNSUPPORTED;
    }
    if (tail) {
      TcTail& t = d.tail;
      t.on = 1; t.sync = w.sync + 4;
      t.dw[0] = a->dweight[0]; t.dw[1] = a->dweight[1]; t.n = (long long)cd;
      t.dbpart = w.db_partials; t.calpart = w.cal_partials; t.nb_db = nb_parts; t.nb_cal = nb_parts; t.C = a->classes;
      t.db[0] = a->dbias[0]; t.db[1] = a->dbias[1]; t.stats = a->stats;
      if (peer) {
        const LfPeerComm& c = *a->grad_comm;
        if (c.n_ranks < 2 || c.n_ranks > LF_MAX_RANKS || c.rank < 0 || c.rank >= c.n_ranks || !c.epoch || !c.error ||
            a->batch_global != a->batch * c.n_ranks) {
          set_error("lf_heads_backward: bad grad_comm");
          return LF_ERR_BAD_ARG;
        }
        t.peer_on = 1; t.comm = c;
        t.n_padded = (int)lf_grad_exchange_floats(a->dim, a->classes);
        t.reg_local = a->mode == LF_MODE_QMF ? a->reg_partial : nullptr;
        t.loss_out = a->mode == LF_MODE_QMF ? a->loss_out : nullptr;
        t.batch_global = a->batch_global;
      }
      if (a->sgd) {
        t.hyper = a->sgd->hyper;
        for (int m = 0; m < 2; ++m) {
          t.param_w[m] = const_cast<float*>(a->weight[m]); t.param_b[m] = const_cast<float*>(a->bias[m]);
          t.mom_w[m] = a->sgd->momentum_buf[2 * m]; t.mom_b[m] = a->sgd->momentum_buf[2 * m + 1];
          t.w16[m] = a->sgd->weight_bf16_out[m];
        }
      }
    }
    rc = tc_gemm(d, s);
    if (rc) return rc;
    if (tail) return LF_OK;
  } else {
    g.splits = splits; g.k_chunk = div_up(a->batch, splits); g.split_stride = cd;
    rc = gemm_dweight(g, 2, s);
  }
  if (rc) return rc;
  // db_m (column sums of dZ_m, accumulated by the kernel that produced dZ) and the calibrated counts
  const int nb_db = nb_parts;
  if (finalize_grads_supported(w.dw_partials, a->dweight[0], a->dweight[1], splits, cd) && (kMaxSplits * cd) % 4 == 0)
    return finalize_grads(w.dw_partials, a->dweight[0], a->dweight[1], splits, kMaxSplits, cd, w.db_partials, nb_db, a->classes,
                          w.cal_partials, nb_parts, a->dbias[0], a->dbias[1], a->stats, s);
  rc = reduce_splits2(w.dw_partials, a->dweight[0], a->dweight[1], splits, kMaxSplits, cd, s);
  if (rc) return rc;
  return finalize_db_cal(w.db_partials, nb_db, a->classes, w.cal_partials, nb_parts, a->dbias[0], a->dbias[1],
                         a->stats, s);
}

// bwd_phase 3 / 4: the row kernel (dL/dz of QMF heads, calibrated counts) and the dW GEMM as separate calls, so that a
// caller whose dL/dz is already final after the forward pass (mean fusion) can run lf_step_mid + the calibrated-count pass
// on a second stream beside the dfeat GEMM.  Not for shapes that form dL/dz inside a fused kernel.
extern "C" int lf_heads_backward_splits_rows(const LfHeadsArgs* a) {
  if (!a || a->dim < 4 || a->classes < 1 || a->batch < 1) return 0;
  if (use_narrow(a)) return 0;
  const int ldz = a->ld_dlogits > 0 ? a->ld_dlogits : a->classes;
  if (use_tensor_pipe(a) && tc_bwd_supported(a->mode, a->precision, a->batch, a->dim, a->classes,
                                              a->ld_logits > 0 ? a->ld_logits : a->classes, ldz, a->need_dfeat)) return 0;
  return 1;
}

extern "C" int lf_heads_backward_fuses_allreduce(const LfHeadsArgs* a) {
  if (!a || a->dim < 4 || a->classes < 1 || a->batch < 1) return 0;
  if (!use_tensor_pipe(a) || use_narrow(a)) return 0;
  return dw_tail_plan(a, nullptr, nullptr) ? 1 : 0;
}

/* floats per rank slot of LfPeerComm.recv_grad for the fused all-reduce: [dW1 | dW2 | db1 | db2 | cal x2 | reg], padded to 16 B */
extern "C" size_t lf_grad_exchange_floats(int32_t dim, int32_t classes) {
  if (dim < 1 || classes < 1) return 0;
  return ((size_t)2 * classes * dim + 2 * (size_t)classes + 3 + 3) / 4 * 4;
}

extern "C" int lf_cast_heads_bf16(const float* w0, const float* w1, void* out16, size_t n_each, void* stream) {
  if (!w0 || !w1 || !out16 || n_each == 0) { set_error("lf_cast_heads_bf16: bad argument"); return LF_ERR_BAD_ARG; }
  return cast_weights_bf16(w0, w1, out16, n_each, (cudaStream_t)stream);
}

extern "C" int lf_loss_finalize(const double* stats, int32_t mode, int32_t batch_global, float* loss_out, void* stream) {
  if (!stats || !loss_out || batch_global < 1) { set_error("lf_loss_finalize: bad argument"); return LF_ERR_BAD_ARG; }
  return loss_finalize(stats, mode, batch_global, loss_out, (cudaStream_t)stream);
}

extern "C" int lf_ema_update(float* ema_x, float* ema_offset, const double* stats, int32_t classes,
                             int32_t batch_global, float smoothing, void* stream) {
  if (!ema_x || !ema_offset || !stats || classes < 1 || batch_global < 1) { set_error("lf_ema_update: bad argument"); return LF_ERR_BAD_ARG; }
  return ema_update(ema_x, ema_offset, stats, classes, batch_global, smoothing, (cudaStream_t)stream);
}

extern "C" int lf_ogm_coeff(const double* stats, float alpha, float* coeff_out, void* stream) {
  if (!stats || !coeff_out) { set_error("lf_ogm_coeff: bad argument"); return LF_ERR_BAD_ARG; }
  return ogm_coeff(stats, alpha, coeff_out, (cudaStream_t)stream);
}
