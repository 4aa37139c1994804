// Fused forward kernel for wide heads (32 <= C <= 256) on the tensor pipe:
//   both heads' logits (tcgen05 kind::tf32, TMA-fed)  ->  per-sample softmax statistics, mean / QMF
//   fusion, cross-entropy terms, OGM-GE scores, accuracy counts, EMA column sums, and (JLOGITS) dL/dz
// all in ONE pass over the features: the logits never round-trip through HBM before the row math.
//
// One CTA per 128 samples.  The two heads' accumulators live side by side in TMEM
// (acc1 at column 0, acc2 at column acc_stride); the epilogue is thread-per-row (TMEM lane = sample), so
// every per-sample reduction over classes is a private register loop — no shuffles.  Outputs are
// transposed 32x32 through padded shared memory so every global store instruction writes 128 contiguous
// bytes, and the same read-back loop yields the per-class column sums the EMA needs for free.
//
// Reference arithmetic: cremad/joint_model_qmf.py:57-75 (QMF), cremad/joint_model_ogm_ge.py:50-58
// (mean fusion), existing_algos/QMF.py:113-117, existing_algos/OGM_GE.py:21-22, utils/BaseModel.py:78-92.
#include <stdlib.h>
#include "lf_common.cuh"
#include "lf_rows.cuh"
#include "lf_tc.cuh"
#include "lf_tc_ptx.cuh"

namespace lf {

constexpr int FWD_THREADS = 192;

struct TcFwdParams {
  int B, Bg, D, C, mode;
  int block_n;       // C rounded up to 16
  int acc_stride;    // TMEM columns between the two accumulators
  int tmem_cols;
  int stages;
  int ldz;
  const float* bias[2];
  const int64_t* label;
  float* z[2];
  float* avg;
  float* zdf;
  float* conf;
  float* rowstat;
  float* dz;
  float* partials;
  float* dbpart;     // [gridDim.x][2][C]  (JLOGITS: column sums of dz)
  int dbg;
};

struct OnlineLse {
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add(float v) {
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }
    else s += __expf(v - m);
  }
  __device__ __forceinline__ float lse() const { return m + logf(s); }
};

__global__ void __launch_bounds__(FWD_THREADS, 2)
tc_heads_forward_kernel(const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapW0,
                        const __grid_constant__ CUtensorMap mapF1, const __grid_constant__ CUtensorMap mapW1,
                        TcFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  uint8_t* tail = smem + (size_t)stages * stage_bytes;
  uint64_t* full_bar = (uint64_t*)tail;
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);
  float* s_bias = (float*)(tail + 256);                 // [2][block_n]
  float* s_colsum = s_bias + 2 * p.block_n;             // [4 warps][2][block_n]
  float* s_stat = s_colsum + 8 * p.block_n;             // [4 warps][16]
  float* s_rows = s_stat + 64;                          // [4 warps][2][32] per-row scalars for the transposed loop
  float* s_dzsum = s_rows + 256;                        // [4 warps][block_n] column sums of dz (JLOGITS)
  float* s_tile = (float*)smem;                         // [4 warps][2][32][33], aliases the drained stages 0..1

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * TC_BLOCK_M;
  const int num_kb = (p.D + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int C = p.C;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapF0); tma_prefetch_desc(&mapW0); tma_prefetch_desc(&mapF1); tma_prefetch_desc(&mapW1);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
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
    // ===================== TMA producer: modality 0's k-blocks, then modality 1's =====================
    if (lane == 0) {
      for (int it = 0; it < 2 * num_kb; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        const int m = it / num_kb, kb = it % num_kb;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
        tma_load_2d(m == 0 ? &mapF0 : &mapF1, &full_bar[s], sa, kb * TC_BLOCK_K, m0);
        tma_load_2d(m == 0 ? &mapW0 : &mapW1, &full_bar[s], sa + a_bytes, kb * TC_BLOCK_K, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, p.block_n, 0, 0);
      for (int it = 0; it < 2 * num_kb; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        const int m = it / num_kb, kb = it % num_kb;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        const int krem = p.D - kb * TC_BLOCK_K;
        const int ksteps = krem >= TC_BLOCK_K ? TC_BLOCK_K / TC_UMMA_K : (krem + TC_UMMA_K - 1) / TC_UMMA_K;
        for (int k = 0; k < ksteps; ++k)
          umma_tf32(tmem_base + (uint32_t)(m * p.acc_stride), make_smem_desc(sa + k * 32, 16, 1024, 2),
                    make_smem_desc(sb + k * 32, 16, 1024, 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue: thread per sample =====================
    const int q = warp & 3;
    const int ew = warp - 2;                                  // 0..3, index for per-warp smem
    float* tile = s_tile + q * (2 * 32 * 33);
    float* tile2 = tile + 32 * 33;
    float* colsum = s_colsum + ew * 2 * p.block_n;
    float* dzsum = s_dzsum + ew * p.block_n;
    for (int c = lane; c < 2 * p.block_n; c += 32) colsum[c] = 0.f;
    for (int c = lane; c < p.block_n; c += 32) dzsum[c] = 0.f;
    const int row0 = m0 + q * 32;
    const int b = row0 + lane;
    const bool live = b < p.B;
    const int y = live ? (int)p.label[b] : -1;
    const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t2 = t1 + (uint32_t)p.acc_stride;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    // The row statistics are written branch-free and in separate max / sum-exp sweeps over TMEM (reads are
    // cheap, 16 columns per tcgen05.ld): an online-softmax update is a serial, divergent dependency
    // chain per element, and with one epilogue warp per scheduler nothing else can hide that latency.
    const float NEG = -INFINITY;
    constexpr float kLog2e = 1.4426950408889634f;

    // ---- sweep A: maxima, argmax (first index on ties, like torch.argmax), z[y]
    float m1 = NEG, m2 = NEG, ma = NEG, zy1 = 0.f, zy2 = 0.f;
    int i1 = 0, i2 = 0, ia = 0;
    const int ncol = (p.dbg & 1) ? 0 : p.block_n;
    for (int c0 = 0; c0 < ncol; c0 += 16) {
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
    // ---- sweep B: sum exp(z - max), two independent partial sums per quantity
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, saa = 0.f, sab = 0.f;
    const float k1 = m1 * kLog2e, k2 = m2 * kLog2e, ka = ma * kLog2e;
    for (int c0 = 0; c0 < ncol; c0 += 16) {
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
    const float lse1 = m1 + logf(s1a + s1b), lse2 = m2 + logf(s2a + s2b), lsea = ma + logf(saa + sab);
    // QMF energy = log(sum(exp z)) un-stabilised in the reference (QMF.py:113): identical to the
    // stabilised value unless the plain sum overflows fp32, where the reference yields +inf
    const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;
    const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;
    const float dz_scale = 0.5f / (float)p.Bg;

    // ---- sweeps C, D (QMF): z_df = c1 z1 + c2 z2 needs the finished row energies
    float md = NEG, zyd = 0.f, lsed = 0.f;
    int idf = 0;
    if (p.mode == LF_MODE_QMF) {
      for (int c0 = 0; c0 < ncol; c0 += 16) {
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
      float sda = 0.f, sdb = 0.f;
      const float kd = md * kLog2e;
      for (int c0 = 0; c0 < ncol; c0 += 16) {
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
      lsed = md + logf(sda + sdb);
    }

    // ---- outputs.  Each warp transposes its 32 rows x 32 classes of z1 and z2 through two padded smem
    // tiles; the read-back loop runs with lane = class, so z1, z2, avg and z_df / dL/dz leave as 128-byte
    // contiguous stores and the EMA column sums accumulate on the way.  Per-row scalars the transposed
    // loop needs (c1,c2 or lse(avg),label) are parked in smem and read back as broadcasts.
    float* rowA = s_rows + ew * 64;
    float* rowB = rowA + 32;
    rowA[lane] = (p.mode == LF_MODE_QMF) ? c1 : lsea;
    rowB[lane] = (p.mode == LF_MODE_QMF) ? c2 : __int_as_float(y);
    const int nrows = min(32, p.B - row0);
    for (int c0 = 0; c0 < ((p.dbg & 2) ? 0 : p.block_n); c0 += 32) {
      __syncwarp();
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        float v[32];
        const uint32_t t = which == 0 ? t1 : t2;
        tmem_ld16(t + c0, v);
        if (c0 + 16 < p.block_n) tmem_ld16(t + c0 + 16, v + 16);
        float* tl = which == 0 ? tile : tile2;
        const float* bs = s_bias + which * p.block_n + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i) tl[lane * 33 + i] = (c0 + i < C) ? v[i] + bs[i] : 0.f;
      }
      __syncwarp();
      const int col = c0 + lane;
      const bool col_ok = col < C;
      float cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
      float* o0 = p.z[0] + (size_t)row0 * C + col;
      float* o1 = p.z[1] + (size_t)row0 * C + col;
      float* o2 = p.avg + (size_t)row0 * C + col;
      float* o3 = (p.mode == LF_MODE_QMF) ? p.zdf + (size_t)row0 * C + col : p.dz + (size_t)row0 * p.ldz + col;
      const int ld3 = (p.mode == LF_MODE_QMF) ? C : p.ldz;
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) {
        const float a1 = tile[r * 33 + lane], a2 = tile2[r * 33 + lane];
        const float av = (a1 + a2) / 2.f;
        cs1 += a1; cs2 += a2;
        const float ra = rowA[r], rb = rowB[r];
        float fourth;
        if (p.mode == LF_MODE_QMF) fourth = a1 * ra + a2 * rb;
        else fourth = (exp2f((av - ra) * kLog2e) - (col == __float_as_int(rb) ? 1.f : 0.f)) * dz_scale;
        cs3 += fourth;
        if (col_ok) {
          o0[(size_t)r * C] = a1;
          o1[(size_t)r * C] = a2;
          o2[(size_t)r * C] = av;
          o3[(size_t)r * ld3] = fourth;
        }
      }
      if (col_ok) { colsum[col] += cs1; colsum[p.block_n + col] += cs2; dzsum[col] += cs3; }
    }

    // ---- per-sample scalars and the CTA's partial statistics
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
      if (lane == 0) s_stat[ew * 16 + i] = s;
    }
    tc_fence_before();
    named_bar_sync(1, 128);                                  // the four epilogue warps only
    float* out = p.partials + (size_t)blockIdx.x * stat_len_dev(C);
    const int et = threadIdx.x - 64;                         // 0..127
    if (et < LF_STATS_HEADER) {
      float s = 0.f;
      if (et < 9) s = (s_stat[et] + s_stat[16 + et]) + (s_stat[32 + et] + s_stat[48 + et]);
      out[et] = s;
    }
    for (int i = et; i < 2 * C; i += 128) {
      const int m = i / C, c = i % C;
      const int o = m * p.block_n + c;
      out[LF_STATS_HEADER + i] = (s_colsum[o] + s_colsum[2 * p.block_n + o]) +
                                 (s_colsum[4 * p.block_n + o] + s_colsum[6 * p.block_n + o]);
    }
    if (p.mode == LF_MODE_JLOGITS)
      for (int c = et; c < C; c += 128) {
        const float d = (s_dzsum[c] + s_dzsum[p.block_n + c]) + (s_dzsum[2 * p.block_n + c] + s_dzsum[3 * p.block_n + c]);
        p.dbpart[(size_t)blockIdx.x * 2 * C + c] = d;            // dz1 == dz2 -> db1 == db2
        p.dbpart[(size_t)blockIdx.x * 2 * C + C + c] = d;
      }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

void finalize_forward_stats(const float* partials, int nblocks, int C, double* stats, cudaStream_t s);  // lf_rows.cu

// Fused forward for 32 <= C <= 256, LF_PREC_TF32.  Returns LF_ERR_UNSUPPORTED when the shape does not fit.
int tc_heads_forward(const LfHeadsArgs* a, float* partials, float* dbpart, float* rowstat, cudaStream_t s) {
  const int C = a->classes;
  if (C > 256) return LF_ERR_UNSUPPORTED;
  TcFwdParams p;
  p.B = a->batch; p.Bg = a->batch_global; p.D = a->dim; p.C = C; p.mode = a->mode;
  p.block_n = div_up(C, 16) * 16;
  p.acc_stride = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;
  p.tmem_cols = 2 * p.acc_stride;
  p.ldz = a->ld_dlogits > 0 ? a->ld_dlogits : C;
  for (int m = 0; m < 2; ++m) { p.bias[m] = a->bias[m]; p.z[m] = a->logits[m]; }
  p.label = a->label; p.avg = a->avg_logits; p.zdf = a->logits_df; p.conf = a->conf;
  p.rowstat = rowstat; p.dz = a->dlogits[0]; p.partials = partials; p.dbpart = dbpart;
  CUtensorMap mF[2], mW[2];
  for (int m = 0; m < 2; ++m) {
    int rc = make_map(&mF[m], a->feat[m], a->dim, a->batch, a->dim, TC_BLOCK_K, TC_BLOCK_M, false);
    if (rc) return rc;
    rc = make_map(&mW[m], a->weight[m], a->dim, C, a->dim, TC_BLOCK_K, p.block_n, false);
    if (rc) return rc;
  }
  const uint32_t stage_bytes = TC_BLOCK_M * TC_BLOCK_K * 4 + p.block_n * TC_BLOCK_K * 4;
  const size_t tail = 256 + (size_t)(2 + 8 + 4) * p.block_n * 4 + 4 * 16 * 4 + 4 * 64 * 4;
  // two CTAs per SM when TMEM allows it (2 x tmem_cols <= 512): one CTA's epilogue hides behind the
  // other's loads, and 256-CTA grids stop paying a second-wave tail
  const size_t cap = ((p.tmem_cols <= 256 && !getenv("LF_FWD_ONECTA")) ? 112 : 224) * 1024;
  int stages = getenv("LF_FWD_STAGES") ? atoi(getenv("LF_FWD_STAGES")) : 6;
  p.dbg = getenv("LF_FWD_DBG") ? atoi(getenv("LF_FWD_DBG")) : 0;
  while (stages > 2 && (size_t)stages * stage_bytes + tail + 1024 > cap) --stages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + tail + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_heads_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    attr_set = true;
  }
  const int nblocks = div_up(a->batch, TC_BLOCK_M);
  LF_LAUNCH("tc_heads_forward", s, (tc_heads_forward_kernel<<<nblocks, FWD_THREADS, smem, s>>>(mF[0], mW[0], mF[1], mW[1], p)));
  int rc = check_launch("tc_heads_forward");
  if (rc) return rc;
  finalize_forward_stats(partials, nblocks, C, a->stats, s);
  return check_launch("finalize_stats");
}

}  // namespace lf
