// Fused forward kernel for wide heads (32 <= C <= 256) on the tensor pipe:
//   both heads' logits (tcgen05 kind::tf32, TMA-fed)  ->  per-sample softmax statistics, mean / QMF
//   fusion, cross-entropy terms, OGM-GE scores, accuracy counts, EMA column sums, and (JLOGITS) dL/dz
// all in ONE pass over the features: the logits never round-trip through HBM before the row math.
//
// Persistent and warp-specialised, one CTA per SM:
//   warp 0      TMA producer; runs ahead across tile boundaries through a multi-stage full/empty ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; the two heads' accumulators sit side by
//               side in TMEM and the pair is DOUBLE-BUFFERED (C <= 128), so tile i+1's loads and MMAs run
//               under tile i's epilogue
//   warps 2-9   epilogue.  Thread = sample (TMEM lane), so every reduction over classes is a private
//               register loop; the two warps that share a lane quarter split the class range in halves
//               and exchange their partial row statistics through shared memory.  Outputs are transposed
//               through padded smem so global stores are contiguous, and the same read-back loop yields
//               the per-class column sums the EMA needs.
// The batch is cut into one contiguous row range per CTA (rows_per_cta = ceil(B / #CTAs)): every SM
// streams the same number of feature bytes, instead of 256 tiles landing 2-vs-1 on 148 SMs.  A range is
// walked in 128-row tiles; the last tile uses a second tensor map whose box holds only the remaining rows.
//
// Reference arithmetic: cremad/joint_model_qmf.py:57-75 (QMF), cremad/joint_model_ogm_ge.py:50-58
// (mean fusion), existing_algos/QMF.py:113-117, existing_algos/OGM_GE.py:21-22, utils/BaseModel.py:78-92.
#include <stdlib.h>
#include "lf_common.cuh"
#include "lf_rows.cuh"
#include "lf_tc.cuh"
#include "lf_tc_ptx.cuh"

namespace lf {

constexpr int FWD_EPI_WARPS = 8;
constexpr int FWD_THREADS = 64 + 32 * FWD_EPI_WARPS;
constexpr int FWD_XCH = 16;                    // floats per thread in the pair-exchange area

struct TcFwdParams {
  int B, Bg, D, C, mode;
  int block_n;        // C rounded up to 16
  int half_n;         // first column of the second class half (multiple of 16)
  int acc_stride;     // TMEM columns between the two heads' accumulators
  int nbuf;           // accumulator pairs in TMEM (2 when 4*acc_stride <= 512)
  int tmem_cols;
  int stages;
  int ldz;
  int rows_per_cta;   // multiple of 8
  int tiles_per_cta;  // ceil(rows_per_cta / 128)
  int rem_rows;       // rows in the last tile's TMA box (== 128 when rows_per_cta % 128 == 0)
  const float* bias[2];
  const int64_t* label;
  float* z[2];
  float* avg;
  float* zdf;
  float* conf;
  float* rowstat;
  float* dz;
  float* partials;    // [gridDim.x * tiles_per_cta][stat_len]
  float* dbpart;      // [gridDim.x * tiles_per_cta][2][C]  (JLOGITS: column sums of dz)
  int prefetch;       // k-blocks of features prefetched into L2 ahead of the smem ring
  int dbg;            // LF_FWD_DBG bits (profiling experiments only): 1 skip row sweeps, 2 skip outputs, 4 skip MMAs, 8 skip W loads
};

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
tc_heads_forward_kernel(const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapF1,
                        const __grid_constant__ CUtensorMap mapR0, const __grid_constant__ CUtensorMap mapR1,
                        const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                        TcFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  uint8_t* tail = smem + (size_t)stages * stage_bytes;
  uint64_t* full_bar = (uint64_t*)tail;
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;          // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty_bar + 2);
  float* s_bias = (float*)(tail + 256);                                  // [2][block_n]
  float* s_colsum = s_bias + 2 * p.block_n;                              // [8 warps][2][block_n]
  float* s_dzsum = s_colsum + FWD_EPI_WARPS * 2 * p.block_n;             // [8 warps][block_n]
  float* s_stat = s_dzsum + FWD_EPI_WARPS * p.block_n;                   // [4][16]
  float* s_rows = s_stat + 64;                                           // [8 warps][2][32]
  float* s_xch = s_rows + FWD_EPI_WARPS * 64;                            // [8 warps][FWD_XCH][32]
  float* s_tile = s_xch + FWD_EPI_WARPS * FWD_XCH * 32;                  // [8 warps][2][32][17]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_kb = (p.D + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int C = p.C;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(p.B, r_begin + p.rows_per_cta);
  const int nt = p.tiles_per_cta;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapF0); tma_prefetch_desc(&mapF1); tma_prefetch_desc(&mapR0); tma_prefetch_desc(&mapR1);
    tma_prefetch_desc(&mapW0); tma_prefetch_desc(&mapW1);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], FWD_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int i = threadIdx.x; i < 2 * p.block_n; i += blockDim.x) {
    const int m = i / p.block_n, c = i % p.block_n;
    s_bias[i] = c < C ? p.bias[m][c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // k-blocks are numbered it = (t * 2 + m) * num_kb + kb.  The feature boxes are prefetched into L2
    // `p.prefetch` k-blocks ahead of the smem ring: HBM latency under load is ~2.5 us, and the ring alone
    // (stages x 16 KB) cannot keep enough bytes in flight to cover it.
    if (lane == 0) {
      const uint32_t n_it = (uint32_t)nt * 2 * num_kb;
      auto feat_map = [&](uint32_t i, int* row0, int* kb) -> const CUtensorMap* {
        const int t = i / (2 * num_kb), m = (i / num_kb) & 1;
        *kb = i % num_kb; *row0 = r_begin + t * TC_BLOCK_M;
        const bool last = (t == nt - 1) && p.rem_rows != TC_BLOCK_M;
        return m == 0 ? (last ? &mapR0 : &mapF0) : (last ? &mapR1 : &mapF1);
      };
      int row0, kb;
      for (uint32_t i = 0; i < (uint32_t)p.prefetch && i < n_it; ++i) {
        const CUtensorMap* mf = feat_map(i, &row0, &kb);
        tma_prefetch_l2_2d(mf, kb * TC_BLOCK_K, row0);
      }
      for (uint32_t it = 0; it < n_it; ++it) {
        if (p.prefetch > 0 && it + p.prefetch < n_it) {
          const CUtensorMap* mf = feat_map(it + p.prefetch, &row0, &kb);
          tma_prefetch_l2_2d(mf, kb * TC_BLOCK_K, row0);
        }
        const CUtensorMap* mf = feat_map(it, &row0, &kb);
        const int t = it / (2 * num_kb), m = (it / num_kb) & 1;
        const bool last = (t == nt - 1) && p.rem_rows != TC_BLOCK_M;
        const uint32_t fa = last ? (uint32_t)p.rem_rows * TC_BLOCK_K * 4 : a_bytes;
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], fa + ((p.dbg & 8) ? 0u : b_bytes));
        tma_load_2d(mf, &full_bar[s], sa, kb * TC_BLOCK_K, row0);
        if (!(p.dbg & 8)) tma_load_2d(m == 0 ? &mapW0 : &mapW1, &full_bar[s], sa + a_bytes, kb * TC_BLOCK_K, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, p.block_n, 0, 0);
      uint32_t it = 0;
      for (int t = 0; t < nt; ++t) {
        const int buf = p.nbuf == 2 ? (t & 1) : 0;
        const uint32_t use = (uint32_t)(t / p.nbuf);                 // how many times this buffer was used before
        mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        for (int m = 0; m < 2; ++m) {
          const uint32_t acc = tmem_base + (uint32_t)((buf * 2 + m) * p.acc_stride);
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % stages;
            const uint32_t ph = (it / stages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + a_bytes;
            const int krem = p.D - kb * TC_BLOCK_K;
            const int ksteps = krem >= TC_BLOCK_K ? TC_BLOCK_K / TC_UMMA_K : (krem + TC_UMMA_K - 1) / TC_UMMA_K;
            for (int k = 0; k < ((p.dbg & 4) ? 0 : ksteps); ++k)
              umma_tf32(acc, make_smem_desc(sa + k * 32, 16, 1024, 2), make_smem_desc(sb + k * 32, 16, 1024, 2), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            if (p.dbg & 16) mbar_arrive_local(&empty_bar[s]); else umma_commit(&empty_bar[s]);
          }
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ===================== epilogue: thread per sample, class range split over a warp pair ==========
    const int ew = warp - 2;                                  // 0..7
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int h = ew >> 2;                                    // class half
    const int et = threadIdx.x - 64;                          // 0..255
    const int cb = h ? p.half_n : 0;
    const int ce_full = h ? p.block_n : min(p.half_n, p.block_n);
    const int ce = (p.dbg & 1) ? cb : ce_full;
    // partner = the other warp with the same lane quarter
    const int pw = ew ^ 4;
    float* xs = s_xch + ew * (FWD_XCH * 32);                  // what I publish
    const float* xr = s_xch + pw * (FWD_XCH * 32);            // what my partner published
    float* tile = s_tile + ew * (2 * 32 * 17);
    float* tile2 = tile + 32 * 17;
    float* colsum = s_colsum + ew * 2 * p.block_n;
    float* dzsum = s_dzsum + ew * p.block_n;
    float* rowA = s_rows + ew * 64;
    float* rowB = rowA + 32;
    const float NEG = -INFINITY;
    constexpr float kLog2e = 1.4426950408889634f;
    const float dz_scale = 0.5f / (float)p.Bg;
    const int ld3 = (p.mode == LF_MODE_QMF) ? C : p.ldz;

    for (int t = 0; t < nt; ++t) {
      const int buf = p.nbuf == 2 ? (t & 1) : 0;
      const uint32_t use = (uint32_t)(t / p.nbuf);
      for (int c = lane; c < 2 * p.block_n; c += 32) colsum[c] = 0.f;
      for (int c = lane; c < p.block_n; c += 32) dzsum[c] = 0.f;
      const int row0 = r_begin + t * TC_BLOCK_M + q * 32;
      const int b = row0 + lane;
      const bool live = b < r_end;
      const int y = live ? (int)p.label[b] : -1;
      const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * p.acc_stride);
      const uint32_t t2 = t1 + (uint32_t)p.acc_stride;

      mbar_wait_warp(&tmem_full_bar[buf], use & 1);
      tc_fence_after();

      // ---- sweep A: maxima, argmax (first index on ties, like torch.argmax), z[y]
      float m1 = NEG, m2 = NEG, ma = NEG, zy1 = 0.f, zy2 = 0.f;
      int i1 = 0x7fffffff, i2 = 0x7fffffff, ia = 0x7fffffff;
      for (int c0 = cb; c0 < ce; c0 += 16) {
        float v1[16], v2[16];
        tmem_ld16(t1 + c0, v1);
        tmem_ld16(t2 + c0, v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          const bool ok = c < C;
          const float a1 = ok ? v1[i] + s_bias[c] : NEG, a2 = ok ? v2[i] + s_bias[p.block_n + c] : NEG;
          const float av = (a1 + a2) / 2.f;
          const bool g1 = a1 > m1, g2 = a2 > m2, ga = av > ma;
          m1 = g1 ? a1 : m1; i1 = g1 ? c : i1;
          m2 = g2 ? a2 : m2; i2 = g2 ? c : i2;
          ma = ga ? av : ma; ia = ga ? c : ia;
          zy1 = (c == y) ? a1 : zy1;
          zy2 = (c == y) ? a2 : zy2;
        }
      }
      xs[0 * 32 + lane] = m1; xs[1 * 32 + lane] = __int_as_float(i1);
      xs[2 * 32 + lane] = m2; xs[3 * 32 + lane] = __int_as_float(i2);
      xs[4 * 32 + lane] = ma; xs[5 * 32 + lane] = __int_as_float(ia);
      xs[6 * 32 + lane] = zy1; xs[7 * 32 + lane] = zy2;
      named_bar_sync(2 + q, 64);
      {
        const float pm1 = xr[0 * 32 + lane], pm2 = xr[2 * 32 + lane], pma = xr[4 * 32 + lane];
        const int pi1 = __float_as_int(xr[1 * 32 + lane]), pi2 = __float_as_int(xr[3 * 32 + lane]),
                  pia = __float_as_int(xr[5 * 32 + lane]);
        if (pm1 > m1 || (pm1 == m1 && pi1 < i1)) { m1 = pm1; i1 = pi1; }
        if (pm2 > m2 || (pm2 == m2 && pi2 < i2)) { m2 = pm2; i2 = pi2; }
        if (pma > ma || (pma == ma && pia < ia)) { ma = pma; ia = pia; }
        zy1 += xr[6 * 32 + lane]; zy2 += xr[7 * 32 + lane];       // exactly one half holds class y, the other 0
      }
      // ---- sweep B: sum exp(z - max), two independent partial sums per quantity
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, saa = 0.f, sab = 0.f;
      const float k1 = m1 * kLog2e, k2 = m2 * kLog2e, ka = ma * kLog2e;
      for (int c0 = cb; c0 < ce; c0 += 16) {
        float v1[16], v2[16];
        tmem_ld16(t1 + c0, v1);
        tmem_ld16(t2 + c0, v2);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = c0 + i + j;
            const bool ok = c < C;
            const float a1 = ok ? v1[i + j] + s_bias[c] : NEG, a2 = ok ? v2[i + j] + s_bias[p.block_n + c] : NEG;
            const float av = (a1 + a2) / 2.f;
            const float e1 = exp2f(fmaf(a1, kLog2e, -k1)), e2 = exp2f(fmaf(a2, kLog2e, -k2)), ea = exp2f(fmaf(av, kLog2e, -ka));
            if (j == 0) { s1a += e1; s2a += e2; saa += ea; } else { s1b += e1; s2b += e2; sab += ea; }
          }
        }
      }
      float s1 = s1a + s1b, s2 = s2a + s2b, sa = saa + sab;
      xs[8 * 32 + lane] = s1; xs[9 * 32 + lane] = s2; xs[10 * 32 + lane] = sa;
      named_bar_sync(2 + q, 64);
      {
        // fixed order (lower class half first) so both warps of the pair get bit-identical sums
        const float o1 = xr[8 * 32 + lane], o2 = xr[9 * 32 + lane], oa = xr[10 * 32 + lane];
        s1 = h ? o1 + s1 : s1 + o1; s2 = h ? o2 + s2 : s2 + o2; sa = h ? oa + sa : sa + oa;
      }
      const float lse1 = m1 + logf(s1), lse2 = m2 + logf(s2), lsea = ma + logf(sa);
      // QMF energy = log(sum(exp z)) un-stabilised in the reference (QMF.py:113): identical to the
      // stabilised value unless the plain sum overflows fp32, where the reference yields +inf
      const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;
      const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;

      // ---- sweeps C, D (QMF): z_df = c1 z1 + c2 z2 needs the finished row energies
      float md = NEG, zyd = 0.f, lsed = 0.f;
      int idf = 0x7fffffff;
      if (p.mode == LF_MODE_QMF) {
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v1[16], v2[16];
          tmem_ld16(t1 + c0, v1);
          tmem_ld16(t2 + c0, v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            const float vd = (c < C) ? (v1[i] + s_bias[c]) * c1 + (v2[i] + s_bias[p.block_n + c]) * c2 : NEG;
            const bool g = vd > md;
            md = g ? vd : md; idf = g ? c : idf;
            zyd = (c == y) ? vd : zyd;
          }
        }
        xs[11 * 32 + lane] = md; xs[12 * 32 + lane] = __int_as_float(idf); xs[13 * 32 + lane] = zyd;
        named_bar_sync(2 + q, 64);
        {
          const float pmd = xr[11 * 32 + lane];
          const int pid = __float_as_int(xr[12 * 32 + lane]);
          if (pmd > md || (pmd == md && pid < idf)) { md = pmd; idf = pid; }
          zyd += xr[13 * 32 + lane];
        }
        float sda = 0.f, sdb = 0.f;
        const float kd = md * kLog2e;
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v1[16], v2[16];
          tmem_ld16(t1 + c0, v1);
          tmem_ld16(t2 + c0, v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            const float vd = (c < C) ? (v1[i] + s_bias[c]) * c1 + (v2[i] + s_bias[p.block_n + c]) * c2 : NEG;
            const float e = exp2f(fmaf(vd, kLog2e, -kd));
            if (i & 1) sdb += e; else sda += e;
          }
        }
        float sd = sda + sdb;
        xs[14 * 32 + lane] = sd;
        named_bar_sync(2 + q, 64);
        const float od = xr[14 * 32 + lane];
        sd = h ? od + sd : sd + od;
        lsed = md + logf(sd);
      }

      // ---- outputs.  Each warp transposes 32 rows x 16 classes of z1 and z2 through two padded smem
      // tiles; in the read-back loop a half-warp covers one row's 16 classes, so z1, z2, avg and
      // z_df / dL/dz leave as 64-byte contiguous runs and the EMA column sums accumulate on the way.
      // Per-row scalars the transposed loop needs (c1,c2 or lse(avg),label) are read back as broadcasts.
      rowA[lane] = (p.mode == LF_MODE_QMF) ? c1 : lsea;
      rowB[lane] = (p.mode == LF_MODE_QMF) ? c2 : __int_as_float(y);
      const int nrows = max(0, min(32, r_end - row0));
      const int sub = lane >> 4, cl = lane & 15;
      for (int c0 = cb; c0 < ((p.dbg & 2) ? cb : ce_full); c0 += 16) {
        __syncwarp();
        {
          float v[16];
          tmem_ld16(t1 + c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tile[lane * 17 + i] = (c0 + i < C) ? v[i] + s_bias[c0 + i] : 0.f;
          tmem_ld16(t2 + c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tile2[lane * 17 + i] = (c0 + i < C) ? v[i] + s_bias[p.block_n + c0 + i] : 0.f;
        }
        __syncwarp();
        const int col = c0 + cl;
        const bool col_ok = col < C;
        float cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
        float* o0 = p.z[0] + (size_t)row0 * C + col;
        float* o1 = p.z[1] + (size_t)row0 * C + col;
        float* o2 = p.avg + (size_t)row0 * C + col;
        float* o3 = (p.mode == LF_MODE_QMF) ? p.zdf + (size_t)row0 * C + col : p.dz + (size_t)row0 * p.ldz + col;
#pragma unroll 4
        for (int r = sub; r < nrows; r += 2) {
          const float a1 = tile[r * 17 + cl], a2 = tile2[r * 17 + cl];
          const float av = (a1 + a2) / 2.f;
          cs1 += a1; cs2 += a2;
          const float ra = rowA[r], rb = rowB[r];
          float fourth;
          if (p.mode == LF_MODE_QMF) fourth = a1 * ra + a2 * rb;
          else fourth = (exp2f((av - ra) * kLog2e) - (col == __float_as_int(rb) ? 1.f : 0.f)) * dz_scale;
          if (col_ok) {
            cs3 += fourth;
            o0[(size_t)r * C] = a1;
            o1[(size_t)r * C] = a2;
            o2[(size_t)r * C] = av;
            o3[(size_t)r * ld3] = fourth;
          }
        }
        // even rows (lanes 0-15) + odd rows (lanes 16-31), fixed order
        const float e1 = __shfl_down_sync(kFull, cs1, 16), e2 = __shfl_down_sync(kFull, cs2, 16), e3 = __shfl_down_sync(kFull, cs3, 16);
        if (sub == 0 && col_ok) { colsum[col] += cs1 + e1; colsum[p.block_n + col] += cs2 + e2; dzsum[col] += cs3 + e3; }
      }
      // all TMEM reads of this tile are done: hand the accumulator pair back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&tmem_empty_bar[buf]);

      // ---- per-sample scalars and this tile's partial statistics (class-half 0 warps own the rows)
      if (h == 0) {
        float st[9];
        float ce_joint;
        int cnt_df = 0;
        if (p.mode == LF_MODE_QMF) {
          ce_joint = lsed - zyd;
          cnt_df = (idf == y);
          if (live) {
            p.conf[b] = c1;
            p.conf[p.B + b] = c2;
            *reinterpret_cast<float4*>(p.rowstat + (size_t)b * 4) = make_float4(lse1, lse2, lsed, 0.f);
          }
        } else {
          ce_joint = lsea - 0.5f * (zy1 + zy2);
        }
        st[LF_STAT_CE_JOINT] = live ? ce_joint : 0.f;
        st[LF_STAT_CE_X1] = live ? lse1 - zy1 : 0.f;
        st[LF_STAT_CE_X2] = live ? lse2 - zy2 : 0.f;
        st[LF_STAT_SCORE_X1] = live ? __expf(zy1 - lse1) : 0.f;
        st[LF_STAT_SCORE_X2] = live ? __expf(zy2 - lse2) : 0.f;
        st[LF_STAT_CNT_X1] = (live && i1 == y) ? 1.f : 0.f;
        st[LF_STAT_CNT_X2] = (live && i2 == y) ? 1.f : 0.f;
        st[LF_STAT_CNT_JOINT] = (live && ia == y) ? 1.f : 0.f;
        st[LF_STAT_CNT_DF] = (live && cnt_df) ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float s = warp_sum(st[i]);
          if (lane == 0) s_stat[q * 16 + i] = s;
        }
      }
      named_bar_sync(1, 32 * FWD_EPI_WARPS);                  // the epilogue warps only
      const size_t prow = (size_t)blockIdx.x * nt + t;
      float* out = p.partials + prow * stat_len_dev(C);
      if (et < LF_STATS_HEADER) {
        float s = 0.f;
        if (et < 9) s = (s_stat[et] + s_stat[16 + et]) + (s_stat[32 + et] + s_stat[48 + et]);
        out[et] = s;
      }
      for (int i = et; i < 2 * C; i += 32 * FWD_EPI_WARPS) {
        const int m = i / C, c = i % C;
        const int o = m * p.block_n + c;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < FWD_EPI_WARPS; ++w) s += s_colsum[w * 2 * p.block_n + o];
        out[LF_STATS_HEADER + i] = s;
      }
      if (p.mode == LF_MODE_JLOGITS)
        for (int c = et; c < C; c += 32 * FWD_EPI_WARPS) {
          float d = 0.f;
#pragma unroll
          for (int w = 0; w < FWD_EPI_WARPS; ++w) d += s_dzsum[w * p.block_n + c];
          p.dbpart[prow * 2 * C + c] = d;                        // dz1 == dz2 -> db1 == db2
          p.dbpart[prow * 2 * C + C + c] = d;
        }
      named_bar_sync(1, 32 * FWD_EPI_WARPS);                  // s_colsum / s_stat are re-zeroed by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

void finalize_forward_stats(const float* partials, int nblocks, int C, double* stats, cudaStream_t s);  // lf_rows.cu

// (#CTAs, rows per CTA) of the balanced row partition; #partial rows = grid * ceil(rows_per_cta / 128)
static void fwd_partition(int B, int* grid, int* rows_per_cta) {
  int g = div_up(B, TC_BLOCK_M);
  if (g > 148) g = 148;
  const int rpc = div_up(div_up(B, g), 8) * 8;
  *grid = div_up(B, rpc);
  *rows_per_cta = rpc;
}
int tc_forward_parts(int B) {
  int grid, rpc;
  fwd_partition(B, &grid, &rpc);
  return grid * div_up(rpc, TC_BLOCK_M);
}

// Fused forward for 32 <= C <= 256, LF_PREC_TF32.  Returns LF_ERR_UNSUPPORTED when the shape does not fit.
// Writes tc_forward_parts(B) per-tile partial rows (and dbpart rows in JLOGITS mode).
int tc_heads_forward(const LfHeadsArgs* a, float* partials, float* dbpart, float* rowstat, cudaStream_t s) {
  const int C = a->classes;
  if (C > 256) return LF_ERR_UNSUPPORTED;
  TcFwdParams p;
  p.B = a->batch; p.Bg = a->batch_global; p.D = a->dim; p.C = C; p.mode = a->mode;
  p.block_n = div_up(C, 16) * 16;
  p.half_n = div_up(div_up(p.block_n, 16), 2) * 16;
  p.acc_stride = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;
  p.nbuf = (4 * p.acc_stride <= 512) ? 2 : 1;
  p.tmem_cols = 2 * p.acc_stride * p.nbuf;
  if (p.tmem_cols < 32) p.tmem_cols = 32;
  p.ldz = a->ld_dlogits > 0 ? a->ld_dlogits : C;
  for (int m = 0; m < 2; ++m) { p.bias[m] = a->bias[m]; p.z[m] = a->logits[m]; }
  p.label = a->label; p.avg = a->avg_logits; p.zdf = a->logits_df; p.conf = a->conf;
  p.rowstat = rowstat; p.dz = a->dlogits[0]; p.partials = partials; p.dbpart = dbpart;

  // one contiguous row range per CTA, as even as 8-row granularity allows
  int grid;
  fwd_partition(a->batch, &grid, &p.rows_per_cta);
  p.tiles_per_cta = div_up(p.rows_per_cta, TC_BLOCK_M);
  p.rem_rows = p.rows_per_cta - (p.tiles_per_cta - 1) * TC_BLOCK_M;

  CUtensorMap mF[2], mR[2], mW[2];
  for (int m = 0; m < 2; ++m) {
    int rc = make_map(&mF[m], a->feat[m], a->dim, a->batch, a->dim, TC_BLOCK_K, TC_BLOCK_M, false);
    if (rc) return rc;
    rc = make_map(&mR[m], a->feat[m], a->dim, a->batch, a->dim, TC_BLOCK_K, p.rem_rows, false);
    if (rc) return rc;
    rc = make_map(&mW[m], a->weight[m], a->dim, C, a->dim, TC_BLOCK_K, p.block_n, false);
    if (rc) return rc;
  }
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = TC_BLOCK_M * TC_BLOCK_K * 4 + ((b_bytes + 1023) & ~1023u);
  const size_t tail = 256 + sizeof(float) * ((size_t)2 * p.block_n + (size_t)FWD_EPI_WARPS * 3 * p.block_n + 64 +
                                              FWD_EPI_WARPS * 64 + FWD_EPI_WARPS * FWD_XCH * 32 +
                                              FWD_EPI_WARPS * 2 * 32 * 17);
  const size_t cap = 226 * 1024;
  int stages = getenv("LF_FWD_STAGES") ? atoi(getenv("LF_FWD_STAGES")) : 8;
  while (stages > 2 && (size_t)stages * stage_bytes + tail + 1024 > cap) --stages;
  if ((size_t)stages * stage_bytes + tail + 1024 > cap) return LF_ERR_UNSUPPORTED;
  p.stages = stages;
  p.prefetch = getenv("LF_FWD_PREFETCH") ? atoi(getenv("LF_FWD_PREFETCH")) : 0;
  p.dbg = getenv("LF_FWD_DBG") ? atoi(getenv("LF_FWD_DBG")) : 0;
  const size_t smem = (size_t)stages * stage_bytes + tail + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_heads_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  LF_LAUNCH("tc_heads_forward", s, (tc_heads_forward_kernel<<<grid, FWD_THREADS, smem, s>>>(mF[0], mF[1], mR[0], mR[1], mW[0], mW[1], p)));
  int rc = check_launch("tc_heads_forward");
  if (rc) return rc;
  const int nparts = grid * p.tiles_per_cta;
  finalize_forward_stats(partials, nparts, C, a->stats, s);
  return check_launch("finalize_stats");
}

}  // namespace lf
