// Fused step for NARROW heads (C <= 32; Crema-D 6-way, Enrico 20-way): exact fp32 FMA, HBM-bound
// (4-19 FLOP/B), so the design goal is one pass over the features with everything
// else on chip.
//
//   JLOGITS (mean fusion / OGM-GE), PASS_BOTH : logits -> softmax/CE/scores/counts -> dL/dz -> dfeat -> dW, db
//                                               in ONE kernel: features are read once, dfeat written once.
//   QMF, PASS_FWD : logits -> energies/conf/z_df/CE/scores/counts           (first read of the features)
//   QMF, PASS_BWD : dL/dz from the stored logits + ranking gradient -> dfeat -> dW, db, calibrated counts
//                   (second read; forced by the mid-step global dependency of QMF)
//
// One persistent CTA per SM owns a contiguous range of samples and walks it in tiles of S samples:
//   * the tile's two feature slabs are contiguous in memory (S x D fp32 each) and arrive by
//     cp.async.bulk (TMA 1-D) into a double-buffered shared-memory stage, tracked by an mbarrier;
//   * W of both heads stays resident in shared memory (24 KB for C = 6, D = 512);
//   * warp-per-sample: lanes hold the feature row as float4 registers (128-bit LDS), the C dot products
//     per head are reduced with a transposing butterfly (~C shuffles instead of 5C), after which lane c
//     holds class c and the softmax / CE / argmax row math is the same warp code as lf_rows_reg.cu;
//   * dfeat = dz W leaves as coalesced float4 stores straight from registers;
//   * dW += dz^T f accumulates per CTA in shared memory with thread-owns-column FMAs (no atomics), is
//     written as one partial per CTA and reduced in fixed order afterwards (deterministic).
#include "lf_common.cuh"
#include "lf_rows.cuh"
#include "lf_rowmath.cuh"
#include "lf_tc_ptx.cuh"

namespace lf {

constexpr int PASS_BOTH = 0, PASS_FWD = 1, PASS_BWD = 2;
constexpr int kNarrowMaxCtas = 148;

struct NarrowParams {
  int B, Bg, D, C, S;            // S = samples per tile (<= 16)
  int rows_per_cta;
  int need_dfeat;
  int ldz;
  const float* feat[2];
  const float* weight[2];
  const float* bias[2];
  const int64_t* label;
  float* z[2];
  float* avg;
  float* zdf;
  float* conf;
  float* rowstat;
  float* dz[2];                  // (B, ldz) also kept in HBM (lf_heads_backward contract / tests)
  float* dfeat[2];
  const float* qmf_g;
  const float* ema_off;
  float* partials;               // [grid][stat_len]
  float* dbpart;                 // [grid][2][C]
  float* calpart;                // [grid][2]
  float* dwpart;                 // [2][kNarrowMaxCtas][C*D]
  int nb_total;                  // partial rows the finalize kernels will read
  float w_joint, w_uni;          // 1 or 0: QMF loss-term ablations (LF_LOSS_* bits)
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// NARROW_THREADS: 512 for the small-class variants (one sample per warp and one dW column per thread per tile
// at D = 512; twice the warps to hide latency), 256 where the register budget of the wide variants needs it.
template <int MODE, int PASS, int CMAX, int KV, int NARROW_THREADS>
__global__ void __launch_bounds__(NARROW_THREADS, 1) narrow_kernel(NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int C = p.C, D = p.D, S = p.S;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = NARROW_THREADS / 32;
  const size_t cd = (size_t)C * D;
  float* Ws = (float*)smem_raw;                                   // [2][C][D]
  float* dWs = Ws + 2 * cd;                                       // [2][C][D]   (PASS != FWD)
  float* tiles = dWs + (PASS == PASS_FWD ? 0 : 2 * cd);           // [2 bufs][2 mods][S][D]
  float* dzs = tiles + (size_t)4 * S * D;                         // [S][2][CMAX]
  float* red = dzs + (size_t)S * 2 * CMAX;                        // [16 warps][3*CMAX + 12] end-of-kernel reduction
  uint64_t* bars = (uint64_t*)(red + nwarp * (3 * CMAX + 12));    // [2]

  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(p.B, r_begin + p.rows_per_cta);
  const int ntiles = r_end > r_begin ? (r_end - r_begin + S - 1) / S : 0;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  for (size_t i = threadIdx.x; i < 2 * cd; i += NARROW_THREADS) {
    Ws[i] = i < cd ? p.weight[0][i] : p.weight[1][i - cd];
    if (PASS != PASS_FWD) dWs[i] = 0.f;
  }
  __syncthreads();

  auto issue = [&](int t) {       // one elected thread: both modalities' slabs of tile t into buffer t & 1
    const int row0 = r_begin + t * S;
    const uint32_t bytes = (uint32_t)min(S, r_end - row0) * D * 4;
    float* dst = tiles + (size_t)(t & 1) * 2 * S * D;
    mbar_expect_tx(&bars[t & 1], 2 * bytes);
    bulk_load_1d(dst, p.feat[0] + (size_t)row0 * D, bytes, &bars[t & 1]);
    bulk_load_1d(dst + (size_t)S * D, p.feat[1] + (size_t)row0 * D, bytes, &bars[t & 1]);
  };
  if (threadIdx.x == 0 && ntiles > 0) issue(0);
#ifdef LF_NARROW_DEBUG
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("narrow: issued tile0 ntiles=%d S=%d C=%d D=%d rpc=%d\n", ntiles, S, C, D, p.rows_per_cta);
#endif

  // per-lane constants: lane c < C owns class c
  const bool cls_ok = lane < C;
  const float b1 = cls_ok ? p.bias[0][lane] : 0.f, b2 = cls_ok ? p.bias[1][lane] : 0.f;
  const float off1 = (PASS == PASS_BWD && cls_ok) ? p.ema_off[lane] : 0.f;
  const float off2 = (PASS == PASS_BWD && cls_ok) ? p.ema_off[C + lane] : 0.f;
  const float dz_scale = 0.5f / (float)p.Bg, invB = 1.f / (float)p.Bg;
  float cs1 = 0.f, cs2 = 0.f, cd1 = 0.f, cd2 = 0.f, cal1 = 0.f, cal2 = 0.f;
  float st[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.f;

  for (int t = 0; t < ntiles; ++t) {
    if (threadIdx.x == 0 && t + 1 < ntiles) issue(t + 1);           // buffer (t+1)&1 was released by the barrier below
    mbar_wait_warp(&bars[t & 1], (t >> 1) & 1);
    const float* tf = tiles + (size_t)(t & 1) * 2 * S * D;
#ifdef LF_NARROW_DEBUG
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("narrow: tile %d arrived\n", t);
#endif
    const int row0 = r_begin + t * S;
    const int rows = min(S, r_end - row0);

    for (int s = warp; s < rows; s += nwarp) {
      const int b = row0 + s;
      const int y = (int)p.label[b];
      const float4* f1 = reinterpret_cast<const float4*>(tf + (size_t)s * D);
      const float4* f2 = reinterpret_cast<const float4*>(tf + (size_t)(S + s) * D);
      float v1[1], v2[1];
      if (PASS != PASS_BWD) {
        // ---- logits: C dot products per head, lanes over D as float4
        float4 r1[KV], r2[KV];
#pragma unroll
        for (int k = 0; k < KV; ++k) {
          const bool ok = (lane + 32 * k) * 4 < D;
          r1[k] = ok ? f1[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
          r2[k] = ok ? f2[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float pa[CMAX], pb[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          float a1 = 0.f, a2 = 0.f;
          if (c < C) {
            const float4* w1 = reinterpret_cast<const float4*>(Ws + (size_t)c * D);
            const float4* w2 = reinterpret_cast<const float4*>(Ws + cd + (size_t)c * D);
#pragma unroll
            for (int k = 0; k < KV; ++k)
              if ((lane + 32 * k) * 4 < D) {
                const float4 x = w1[lane + 32 * k], yv = w2[lane + 32 * k];
                a1 = fmaf(r1[k].x, x.x, a1); a1 = fmaf(r1[k].y, x.y, a1); a1 = fmaf(r1[k].z, x.z, a1); a1 = fmaf(r1[k].w, x.w, a1);
                a2 = fmaf(r2[k].x, yv.x, a2); a2 = fmaf(r2[k].y, yv.y, a2); a2 = fmaf(r2[k].z, yv.z, a2); a2 = fmaf(r2[k].w, yv.w, a2);
              }
          }
          pa[c] = a1; pb[c] = a2;
        }
        const float t1 = transpose_reduce<CMAX>(pa, lane), t2 = transpose_reduce<CMAX>(pb, lane);
        const int src = (lane * (32 / CMAX)) & 31;                 // lane c <- a lane that holds class c
        const float g1v = __shfl_sync(kFull, t1, src), g2v = __shfl_sync(kFull, t2, src);   // all lanes take part
        v1[0] = cls_ok ? g1v + b1 : -INFINITY;
        v2[0] = cls_ok ? g2v + b2 : -INFINITY;
        if (cls_ok) { p.z[0][(size_t)b * C + lane] = v1[0]; p.z[1][(size_t)b * C + lane] = v2[0]; }
      } else {
        v1[0] = cls_ok ? p.z[0][(size_t)b * C + lane] : -INFINITY;
        v2[0] = cls_ok ? p.z[1][(size_t)b * C + lane] : -INFINITY;
      }

      float d1 = 0.f, d2 = 0.f;                                    // dL/dz of class `lane`, heads 1 / 2
      if (PASS != PASS_BWD) {
        // ---- forward row math (same formulas as lf_rows_reg.cu)
        float av[1] = {(v1[0] + v2[0]) / 2.f};
        if (cls_ok) { p.avg[(size_t)b * C + lane] = av[0]; cs1 += v1[0]; cs2 += v2[0]; }
        float m1, m2, ma; int i1, i2, ia;
        warp_max_arg<1>(v1, lane, m1, i1);
        warp_max_arg<1>(v2, lane, m2, i2);
        warp_max_arg<1>(av, lane, ma, ia);
        float s1 = exp_sub(v1[0], m1 * 1.4426950408889634f), s2 = exp_sub(v2[0], m2 * 1.4426950408889634f),
              sa = exp_sub(av[0], ma * 1.4426950408889634f);
        warp_sum3(s1, s2, sa);
        const float lse1 = m1 + __logf(s1), lse2 = m2 + __logf(s2), lsea = ma + __logf(sa);
        const bool yok = (unsigned)y < (unsigned)C;
        const float sy1 = __shfl_sync(kFull, v1[0], y & 31), sy2 = __shfl_sync(kFull, v2[0], y & 31);
        const float zy1 = yok ? sy1 : 0.f, zy2 = yok ? sy2 : 0.f;
        float ce_joint;
        int cnt_df = 0;
        if (MODE == LF_MODE_QMF) {
          const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;      // un-stabilised energy, QMF.py:113
          const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;
          float vd[1] = {cls_ok ? v1[0] * c1 + v2[0] * c2 : -INFINITY};
          if (cls_ok) p.zdf[(size_t)b * C + lane] = vd[0];
          float md; int idf;
          warp_max_arg<1>(vd, lane, md, idf);
          const float sd = warp_sum(exp_sub(vd[0], md * 1.4426950408889634f));
          const float lsed = md + __logf(sd);
          const float syd = __shfl_sync(kFull, vd[0], y & 31);
          const float zyd = yok ? syd : 0.f;
          ce_joint = lsed - zyd;
          cnt_df = (idf == y);
          if (lane == 0) {
            p.conf[b] = c1; p.conf[p.B + b] = c2;
            *reinterpret_cast<float4*>(p.rowstat + (size_t)b * 4) = make_float4(lse1, lse2, lsed, 0.f);
          }
        } else {
          ce_joint = lsea - 0.5f * (zy1 + zy2);
          if (cls_ok) {
            d1 = d2 = (exp_sub(av[0], lsea * 1.4426950408889634f) - (lane == y ? 1.f : 0.f)) * dz_scale;
            p.dz[0][(size_t)b * p.ldz + lane] = d1;
            cd1 += d1;
          }
        }
        st[LF_STAT_CE_JOINT] += ce_joint;
        st[LF_STAT_CE_X1] += lse1 - zy1;
        st[LF_STAT_CE_X2] += lse2 - zy2;
        st[LF_STAT_SCORE_X1] += __expf(zy1 - lse1);
        st[LF_STAT_SCORE_X2] += __expf(zy2 - lse2);
        st[LF_STAT_CNT_X1] += (i1 == y);
        st[LF_STAT_CNT_X2] += (i2 == y);
        st[LF_STAT_CNT_JOINT] += (ia == y);
        st[LF_STAT_CNT_DF] += cnt_df;
      } else {
        // ---- QMF backward row math + calibrated counts (utils/BaseModel.py:84-89)
        const float c1 = p.conf[b], c2 = p.conf[p.B + b];
        const float4 rs = *reinterpret_cast<const float4*>(p.rowstat + (size_t)b * 4);
        const float g1 = p.qmf_g[b] / 10.f, g2 = p.qmf_g[p.B + b] / 10.f;
        if (cls_ok) {
          const float oh = (lane == y) ? 1.f : 0.f;
          const float p1 = exp_sub(v1[0], rs.x * 1.4426950408889634f), p2 = exp_sub(v2[0], rs.y * 1.4426950408889634f);
          const float pd = exp_sub(v1[0] * c1 + v2[0] * c2, rs.z * 1.4426950408889634f) - oh;
          d1 = (p.w_uni * (p1 - oh) + (p.w_joint * c1) * pd) * invB + g1 * p1;
          d2 = (p.w_uni * (p2 - oh) + (p.w_joint * c2) * pd) * invB + g2 * p2;
          p.dz[0][(size_t)b * p.ldz + lane] = d1;
          p.dz[1][(size_t)b * p.ldz + lane] = d2;
          cd1 += d1; cd2 += d2;
        }
        float w1[1] = {v1[0] + off1}, w2[1] = {v2[0] + off2};
        float m; int i1, i2;
        warp_max_arg<1>(w1, lane, m, i1);
        warp_max_arg<1>(w2, lane, m, i2);
        cal1 += (i1 == y); cal2 += (i2 == y);
      }

      if (PASS != PASS_FWD) {
        // dz of this sample to smem for the dW phase
        if (lane < CMAX) { dzs[(s * 2 + 0) * CMAX + lane] = d1; dzs[(s * 2 + 1) * CMAX + lane] = d2; }
        if (p.need_dfeat) {
          // ---- dfeat = dz W : every lane needs all C coefficients, then float4 FMAs against resident W
          float4 o1[KV], o2[KV];
#pragma unroll
          for (int k = 0; k < KV; ++k) { o1[k] = make_float4(0.f, 0.f, 0.f, 0.f); o2[k] = o1[k]; }
#pragma unroll
          for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
              const float e1 = __shfl_sync(kFull, d1, c), e2 = __shfl_sync(kFull, d2, c);
              const float4* w1 = reinterpret_cast<const float4*>(Ws + (size_t)c * D);
              const float4* w2 = reinterpret_cast<const float4*>(Ws + cd + (size_t)c * D);
#pragma unroll
              for (int k = 0; k < KV; ++k)
                if ((lane + 32 * k) * 4 < D) {
                  const float4 x = w1[lane + 32 * k], yv = w2[lane + 32 * k];
                  o1[k].x = fmaf(e1, x.x, o1[k].x); o1[k].y = fmaf(e1, x.y, o1[k].y); o1[k].z = fmaf(e1, x.z, o1[k].z); o1[k].w = fmaf(e1, x.w, o1[k].w);
                  o2[k].x = fmaf(e2, yv.x, o2[k].x); o2[k].y = fmaf(e2, yv.y, o2[k].y); o2[k].z = fmaf(e2, yv.z, o2[k].z); o2[k].w = fmaf(e2, yv.w, o2[k].w);
                }
            }
          }
          float4* g1p = reinterpret_cast<float4*>(p.dfeat[0] + (size_t)b * D);
          float4* g2p = reinterpret_cast<float4*>(p.dfeat[1] + (size_t)b * D);
#pragma unroll
          for (int k = 0; k < KV; ++k)
            if ((lane + 32 * k) * 4 < D) { stg_stream(g1p + lane + 32 * k, o1[k]); stg_stream(g2p + lane + 32 * k, o2[k]); }
        }
      }
    }
#ifdef LF_NARROW_DEBUG
    if (lane == 0 && blockIdx.x == 0) printf("narrow: warp %d done P1 tile %d\n", warp, t);
#endif
    __syncthreads();                                                // dzs complete; all warps done with P1 reads of the tile

    if (PASS != PASS_FWD) {
      // ---- dW += dz^T f : thread owns columns d = tid, tid + 256, ...; dz broadcast from smem
      for (int d = threadIdx.x; d < D; d += NARROW_THREADS) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          float acc[CMAX];
#pragma unroll
          for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
          const float* fm = tf + (size_t)m * S * D + d;
          for (int s = 0; s < rows; ++s) {
            const float fv = fm[(size_t)s * D];
            const float4* dzv = reinterpret_cast<const float4*>(dzs + (s * 2 + m) * CMAX);
#pragma unroll
            for (int c4 = 0; c4 < CMAX / 4; ++c4) {
              if (c4 * 4 < C) {
                const float4 q = dzv[c4];
                acc[c4 * 4 + 0] = fmaf(q.x, fv, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(q.y, fv, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(q.z, fv, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(q.w, fv, acc[c4 * 4 + 3]);
              }
            }
          }
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) dWs[(size_t)m * cd + (size_t)c * D + d] += acc[c];
        }
      }
    }
    __syncthreads();                                                // tile buffer t&1 and dzs may be overwritten
  }

  // ---- per-CTA partials, fixed order
  float* my = red + warp * (3 * CMAX + 12);
  if (lane < CMAX) { my[lane] = cs1; my[CMAX + lane] = cs2; my[2 * CMAX + lane] = (PASS == PASS_BWD) ? cd2 : cd1; }
  if (PASS == PASS_BWD && lane < CMAX) my[lane] = cd1;              // backward: [dz1 sums | - | dz2 sums]
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) my[3 * CMAX + i] = st[i];
    my[3 * CMAX + 9] = cal1; my[3 * CMAX + 10] = cal2;
  }
  __syncthreads();
  auto rsum = [&](int off) { float s = 0.f; for (int w = 0; w < nwarp; ++w) s += red[w * (3 * CMAX + 12) + off]; return s; };
  if (PASS != PASS_BWD) {
    float* out = p.partials + (size_t)blockIdx.x * stat_len_dev(C);
    if (threadIdx.x < LF_STATS_HEADER) out[threadIdx.x] = threadIdx.x < 9 ? rsum(3 * CMAX + threadIdx.x) : 0.f;
    for (int i = threadIdx.x; i < 2 * C; i += NARROW_THREADS) out[LF_STATS_HEADER + i] = rsum((i / C) * CMAX + (i % C));
    if (PASS == PASS_BOTH)
      for (int c = threadIdx.x; c < C; c += NARROW_THREADS) {
        const float s = rsum(2 * CMAX + c);
        p.dbpart[(size_t)blockIdx.x * 2 * C + c] = s; p.dbpart[(size_t)blockIdx.x * 2 * C + C + c] = s;   // dz1 == dz2
      }
  } else {
    for (int i = threadIdx.x; i < 2 * C; i += NARROW_THREADS)
      p.dbpart[(size_t)blockIdx.x * 2 * C + i] = rsum((i / C) * 2 * CMAX + (i % C));
    if (threadIdx.x < 2) p.calpart[(size_t)blockIdx.x * 2 + threadIdx.x] = rsum(3 * CMAX + 9 + threadIdx.x);
  }
  if (PASS != PASS_FWD)
    for (size_t i = threadIdx.x; i < 2 * cd; i += NARROW_THREADS) {
      const int m = i >= cd;
      p.dwpart[((size_t)m * kNarrowMaxCtas + blockIdx.x) * cd + (i - m * cd)] = dWs[i];
    }
  // partial rows no CTA owns (the finalize kernels sum nb_total rows)
  for (int r = blockIdx.x + gridDim.x; r < p.nb_total; r += gridDim.x) {
    if (PASS != PASS_BWD)
      for (int c = threadIdx.x; c < stat_len_dev(C); c += NARROW_THREADS) p.partials[(size_t)r * stat_len_dev(C) + c] = 0.f;
    if (PASS != PASS_FWD)
      for (int c = threadIdx.x; c < 2 * C; c += NARROW_THREADS) p.dbpart[(size_t)r * 2 * C + c] = 0.f;
    if (PASS == PASS_BWD && threadIdx.x < 2) p.calpart[(size_t)r * 2 + threadIdx.x] = 0.f;
  }
}

// ---------------------------------------------------------------------------------- host side
static size_t narrow_smem(int C, int D, int S, int cmax, bool fwd_only) {
  const size_t cd = (size_t)C * D;
  return sizeof(float) * ((fwd_only ? 2 : 4) * cd + (size_t)4 * S * D + (size_t)S * 2 * cmax + 16 * (3 * cmax + 12)) + 64;
}

// Largest tile (<= 16 samples) that fits next to the resident weights; 0 = shape not supported here.
int narrow_tile(int C, int D, bool fwd_only) {
  if (C > 32 || D > 1024 || D % 4) return 0;
  const int cmax = C <= 8 ? 8 : C <= 16 ? 16 : 32;
  for (int S = 16; S >= 2; S >>= 1)
    if (narrow_smem(C, D, S, cmax, fwd_only) <= 220 * 1024) return S;
  return 0;
}
size_t narrow_dw_floats(int C, int D) { return (size_t)2 * kNarrowMaxCtas * C * D; }

template <int MODE, int PASS>
static int narrow_launch(const NarrowParams& p, int grid, size_t smem, cudaStream_t s, const char* name) {
  const int kv = p.D <= 512 ? 4 : 8;
#define LF_NARROW_GO(CM, KVV)                                                                                     \
  do {                                                                                                            \
    static bool attr = false;                                                                                     \
    constexpr int NT = (CM <= 16 && KVV == 4) ? 512 : 256;                                                       \
    if (!attr) { cudaFuncSetAttribute(narrow_kernel<MODE, PASS, CM, KVV, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; } \
    LF_LAUNCH(name, s, (narrow_kernel<MODE, PASS, CM, KVV, NT><<<grid, NT, smem, s>>>(p)));                       \
    return check_launch(name);                                                                                    \
  } while (0)
  if (p.C <= 8) { if (kv == 4) LF_NARROW_GO(8, 4); else LF_NARROW_GO(8, 8); }
  if (p.C <= 16) { if (kv == 4) LF_NARROW_GO(16, 4); else LF_NARROW_GO(16, 8); }
  if (kv == 4) LF_NARROW_GO(32, 4); else LF_NARROW_GO(32, 8);
#undef LF_NARROW_GO
}

// pass: 0 = JLOGITS fwd+bwd fused, 1 = QMF forward, 2 = QMF backward.  Returns the grid size in *grid_out.
int narrow_run(const LfHeadsArgs* a, int pass, float* partials, float* dbpart, float* calpart, float* dwpart,
               float* rowstat, int nb_total, int* grid_out, cudaStream_t s) {
  const bool fwd_only = pass == PASS_FWD;
  NarrowParams p;
  p.B = a->batch; p.Bg = a->batch_global; p.D = a->dim; p.C = a->classes;
  p.S = narrow_tile(a->classes, a->dim, fwd_only);
  if (p.S == 0) { set_error("narrow heads: shape C=%d D=%d not supported", a->classes, a->dim); return LF_ERR_UNSUPPORTED; }
  int grid = div_up(a->batch, p.S);
  if (grid > kNarrowMaxCtas) grid = kNarrowMaxCtas;
  p.rows_per_cta = div_up(a->batch, grid);          // any row count: the last tile of a range is partial
  grid = div_up(a->batch, p.rows_per_cta);
  p.need_dfeat = a->need_dfeat;
  p.ldz = a->ld_dlogits > 0 ? a->ld_dlogits : a->classes;
  for (int m = 0; m < 2; ++m) {
    p.feat[m] = a->feat[m]; p.weight[m] = a->weight[m]; p.bias[m] = a->bias[m]; p.z[m] = a->logits[m];
    p.dz[m] = a->dlogits[m]; p.dfeat[m] = a->dfeat[m];
  }
  p.label = a->label; p.avg = a->avg_logits; p.zdf = a->logits_df; p.conf = a->conf; p.rowstat = rowstat;
  p.qmf_g = a->qmf_g; p.ema_off = a->ema_offset;
  p.w_joint = (a->loss_terms & LF_LOSS_NO_JOINT) ? 0.f : 1.f;
  p.w_uni = (a->loss_terms & LF_LOSS_NO_UNI) ? 0.f : 1.f;
  p.partials = partials; p.dbpart = dbpart; p.calpart = calpart; p.dwpart = dwpart; p.nb_total = nb_total;
  const int cmax = p.C <= 8 ? 8 : p.C <= 16 ? 16 : 32;
  const size_t smem = narrow_smem(p.C, p.D, p.S, cmax, fwd_only);
  if (grid_out) *grid_out = grid;
  if (pass == PASS_BOTH) return narrow_launch<LF_MODE_JLOGITS, PASS_BOTH>(p, grid, smem, s, "narrow_step_jlogits");
  if (pass == PASS_FWD && a->mode == LF_MODE_JLOGITS) return narrow_launch<LF_MODE_JLOGITS, PASS_FWD>(p, grid, smem, s, "narrow_forward_jlogits");
  if (pass == PASS_FWD) return narrow_launch<LF_MODE_QMF, PASS_FWD>(p, grid, smem, s, "narrow_forward_qmf");
  return narrow_launch<LF_MODE_QMF, PASS_BWD>(p, grid, smem, s, "narrow_backward_qmf");
}

}  // namespace lf
