// Vectorised per-sample ("row") kernels for logits with a 16-byte row pitch (the tensor-pipe path: the
// GEMM epilogue TMA-stores z1/z2 with ld = ceil4(C)).  Same arithmetic and the same partial-statistics
// layouts as lf_rows_reg.cu; what changes is the mapping to the machine:
//
//   * G lanes per sample (G = 8 for C <= 128, 16 for C <= 256; wider heads stay on lf_rows_reg.cu), each lane owning NK float4
//     chunks of the row (columns 4(l + G k) .. +3), so a warp works on 32/G samples per iteration.  The
//     per-sample scalar work (log, energy, CE terms, counts) is shared by 32/G samples per instruction
//     instead of being repeated in 32 lanes for one sample, and the reductions over classes take
//     log2(G) shuffle steps instead of 5.  (The one-warp-per-sample kernels were issue-bound: ~650
//     warp instructions per sample at C = 101; this mapping needs ~160.)
//   * every access is a 128-bit load or store: z1/z2 in, avg / z_df (when their pitch is padded too) and
//     dz (fp32 x4 or bf16 x4) out.
//   * no software prefetch: 16 resident warps x 8 independent 512-byte loads cover the latency (tried:
//     prefetch.global.L1 of the next iteration's rows made both kernels ~30 % slower).
//
// Reference arithmetic: see lf_rows.cu (formulas and reference line numbers).
#include "lf_common.cuh"
#include "lf_rows.cuh"
#include "lf_rowmath.cuh"
#include "lf_rowvec.cuh"

namespace lf {

using namespace rowvec;

namespace {

// dL/dz chunk store: fp32 x4 (16 bytes) or bf16 x4 (8 bytes); the pitch is a multiple of 4 elements
template <int G, int NK>
__device__ __forceinline__ void store_dz_chunks(const RowsArgs& a, int m, size_t row, const float (&d)[NK * 4], int l, int C) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c0 = 4 * (l + G * k);
    if (c0 >= C) continue;
    if (a.dz_bf16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(d[4 * k], d[4 * k + 1]);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(d[4 * k + 2], d[4 * k + 3]);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.dz[m]) + row * a.ldz + c0) = u;
    } else {
      *reinterpret_cast<float4*>(a.dz[m] + row * a.ldz + c0) = make_float4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    }
  }
}

}  // namespace

template <int MODE, int G, int NK>
__global__ void __launch_bounds__(256, 2) rows_forward_vec_kernel(RowsArgs a) {
  extern __shared__ float smem[];                       // [8 warps][3][C] column sums, then [8][9] stats
  constexpr int SPW = 32 / G, NE = NK * 4;
  const int C = a.C, B = a.B, ld = a.ld_z, nq = ld / 4;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  const int l = lane % G, gi = lane / G;
  const float dz_scale = 0.5f / (float)a.B_global;
  const bool vec_out = a.ld_f % 4 == 0;

  pdl_wait();
  pdl_trigger();
  float cs1[NE], cs2[NE], cs3[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) { cs1[i] = 0.f; cs2[i] = 0.f; cs3[i] = 0.f; }
  float st[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.f;

  const int nquads = (B + SPW - 1) / SPW;
  for (int qd = blockIdx.x * nwarp + warp; qd < nquads; qd += gridDim.x * nwarp) {
    const int bs = qd * SPW + gi;
    const bool valid = bs < B;
    const int b = valid ? bs : B - 1;
    const size_t off = (size_t)b * ld;
    Row<G, NK> r1, r2;
    r1.load(a.z[0] + off, l, nq);
    r2.load(a.z[1] + off, l, nq);
    const int y = (int)a.label[b];
    const bool yok = (unsigned)y < (unsigned)C;
    const float zy1 = yok ? __ldg(a.z[0] + off + y) : 0.f, zy2 = yok ? __ldg(a.z[1] + off + y) : 0.f;
    mask_cols<G, NK>(r1.v, l, C, 0.f);                   // pitch padding is uninitialised memory
    mask_cols<G, NK>(r2.v, l, C, 0.f);
    if (valid) {
#pragma unroll
      for (int i = 0; i < NE; ++i) { cs1[i] += r1.v[i]; cs2[i] += r2.v[i]; }
    }
    mask_cols<G, NK>(r1.v, l, C, -INFINITY);             // ... and must stay out of the maxima / sums of exp
    mask_cols<G, NK>(r2.v, l, C, -INFINITY);

    float av[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) av[i] = (r1.v[i] + r2.v[i]) / 2.f;
    if (valid) store_chunks_f32<G, NK>(a.avg + (size_t)b * a.ld_f, av, l, C, vec_out);
    float m1, m2, ma; int i1, i2, ia;
    row_max_arg<G, NK>(r1.v, l, m1, i1);
    row_max_arg<G, NK>(r2.v, l, m2, i2);
    row_max_arg<G, NK>(av, l, ma, ia);
    float s1 = 0.f, s2 = 0.f, sa = 0.f;
    {
      const float k1 = m1 * kLog2e, k2 = m2 * kLog2e, ka = ma * kLog2e;
#pragma unroll
      for (int i = 0; i < NE; ++i) { s1 += exp_sub(r1.v[i], k1); s2 += exp_sub(r2.v[i], k2); sa += exp_sub(av[i], ka); }
    }
    s1 = group_sum<G>(s1); s2 = group_sum<G>(s2); sa = group_sum<G>(sa);
    const float lse1 = m1 + __logf(s1), lse2 = m2 + __logf(s2), lsea = ma + __logf(sa);

    float ce_joint;
    int cnt_df = 0;
    if (MODE == LF_MODE_QMF) {
      // energy = log(sum(exp z)) is NOT stabilised in the reference (QMF.py:113): same value unless the
      // plain fp32 sum overflows, where the reference yields +inf
      const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;
      const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;
      float vd[NE];
#pragma unroll
      for (int i = 0; i < NE; ++i) vd[i] = r1.v[i] * c1 + r2.v[i] * c2;
      mask_cols<G, NK>(vd, l, C, -INFINITY);             // (-inf * c) may be +inf / NaN
      if (valid) store_chunks_f32<G, NK>(a.zdf + (size_t)b * a.ld_f, vd, l, C, vec_out);
      float md; int idf;
      row_max_arg<G, NK>(vd, l, md, idf);
      float sd = 0.f;
      const float kd = md * kLog2e;
#pragma unroll
      for (int i = 0; i < NE; ++i) sd += exp_sub(vd[i], kd);
      sd = group_sum<G>(sd);
      const float lsed = md + __logf(sd);
      const float zyd = yok ? zy1 * c1 + zy2 * c2 : 0.f;
      ce_joint = lsed - zyd;
      cnt_df = (idf == y);
      if (valid && l == 0) {
        a.conf[b] = c1;
        a.conf[B + b] = c2;
        *reinterpret_cast<float4*>(a.rowstat + (size_t)b * 4) = make_float4(lse1, lse2, lsed, 0.f);
      }
    } else {
      ce_joint = lsea - 0.5f * (zy1 + zy2);
      float d[NE];
      const float kl = lsea * kLog2e;
#pragma unroll
      for (int i = 0; i < NE; ++i) {            // dL/dz1 = dL/dz2 = (softmax(avg) - onehot) / (2 Bg)
        const int c = col_of<G, NK>(l, i);
        d[i] = (exp_sub(av[i], kl) - (c == y ? 1.f : 0.f)) * dz_scale;
      }
      mask_cols<G, NK>(d, l, C, 0.f);
      if (valid) {
        store_dz_chunks<G, NK>(a, 0, (size_t)b, d, l, C);
#pragma unroll
        for (int i = 0; i < NE; ++i) cs3[i] += d[i];
      }
    }
    if (valid) {
      st[LF_STAT_CE_JOINT] += ce_joint;
      st[LF_STAT_CE_X1] += lse1 - zy1;
      st[LF_STAT_CE_X2] += lse2 - zy2;
      st[LF_STAT_SCORE_X1] += __expf(zy1 - lse1);
      st[LF_STAT_SCORE_X2] += __expf(zy2 - lse2);
      st[LF_STAT_CNT_X1] += (i1 == y);
      st[LF_STAT_CNT_X2] += (i2 == y);
      st[LF_STAT_CNT_JOINT] += (ia == y);
      st[LF_STAT_CNT_DF] += cnt_df;
    }
  }

  // ---- warp: fold the 32/G lane groups; CTA: fixed-order sum over warps -> one partial row per CTA
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    cs1[i] = across_groups_sum<G>(cs1[i]); cs2[i] = across_groups_sum<G>(cs2[i]);
    if (MODE == LF_MODE_JLOGITS) cs3[i] = across_groups_sum<G>(cs3[i]);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = across_groups_sum<G>(st[i]);     // every lane of a group holds the same value
  float* colsum = smem + (size_t)warp * 3 * C;
  if (gi == 0) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int c = col_of<G, NK>(l, i);
      if (c < C) { colsum[c] = cs1[i]; colsum[C + c] = cs2[i]; colsum[2 * C + c] = cs3[i]; }
    }
  }
  float* sst = smem + (size_t)nwarp * 3 * C;
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 9; ++i) sst[warp * 9 + i] = st[i];
  __syncthreads();
  float* out = a.partials + (size_t)blockIdx.x * stat_len_dev(C);
  if (threadIdx.x < LF_STATS_HEADER) {
    float s = 0.f;
    if (threadIdx.x < 9)
      for (int w = 0; w < nwarp; ++w) s += sst[w * 9 + threadIdx.x];
    out[threadIdx.x] = s;
  }
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += smem[(size_t)w * 3 * C + c];
    out[LF_STATS_HEADER + c] = s;
  }
  if (MODE == LF_MODE_JLOGITS)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += smem[(size_t)w * 3 * C + 2 * C + c];
      a.dbpart[(size_t)blockIdx.x * 2 * C + c] = s;          // dz1 == dz2 -> db1 == db2
      a.dbpart[(size_t)blockIdx.x * 2 * C + C + c] = s;
    }
  for (int r = blockIdx.x + gridDim.x; r < a.nb_total; r += gridDim.x) {     // rows no CTA owns
    for (int c = threadIdx.x; c < stat_len_dev(C); c += blockDim.x) a.partials[(size_t)r * stat_len_dev(C) + c] = 0.f;
    if (MODE == LF_MODE_JLOGITS)
      for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) a.dbpart[(size_t)r * 2 * C + c] = 0.f;
  }
}

template <int MODE, int G, int NK>
__global__ void __launch_bounds__(256, 2) rows_backward_vec_kernel(RowsArgs a) {
  extern __shared__ float smem[];                       // [8 warps][2][C] column sums of dz (QMF)
  constexpr int SPW = 32 / G, NE = NK * 4;
  const int C = a.C, B = a.B, ld = a.ld_z, nq = ld / 4;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  const int l = lane % G, gi = lane / G;
  const float invB = 1.f / (float)a.B_global;

  // EMA offsets of this step, padded with -inf (keeps padded columns out of the calibrated argmax), in
  // shared memory after the column-sum area: registers are the scarce resource here
  pdl_wait();
  pdl_trigger();
  constexpr int LDP = 4 * G * NK;
  float* soff = smem + (MODE == LF_MODE_QMF ? (size_t)8 * 2 * C : 0);
  for (int c = threadIdx.x; c < 2 * LDP; c += blockDim.x) {
    const int m = c / LDP, cc = c - m * LDP;
    soff[c] = cc < C ? a.ema_off[m * C + cc] : -INFINITY;
  }
  __syncthreads();
  float d1s[NE], d2s[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) { d1s[i] = 0.f; d2s[i] = 0.f; }
  float cal1 = 0.f, cal2 = 0.f;

  const int nquads = (B + SPW - 1) / SPW;
  for (int qd = blockIdx.x * nwarp + warp; qd < nquads; qd += gridDim.x * nwarp) {
    const int bs = qd * SPW + gi;
    const bool valid = bs < B;
    const int b = valid ? bs : B - 1;
    const size_t off = (size_t)b * ld;
    Row<G, NK> r1, r2;
    r1.load(a.z[0] + off, l, nq);
    r2.load(a.z[1] + off, l, nq);
    mask_cols<G, NK>(r1.v, l, C, 0.f);                   // pitch padding is uninitialised memory
    mask_cols<G, NK>(r2.v, l, C, 0.f);
    const int y = (int)a.label[b];
    if (MODE == LF_MODE_QMF) {
      const float c1 = a.conf[b], c2 = a.conf[B + b];
      const float4 rs = *reinterpret_cast<const float4*>(a.rowstat + (size_t)b * 4);          // lse1, lse2, lse(z_df)
      const float g1 = a.qmf_g[b] / 10.f, g2 = a.qmf_g[B + b] / 10.f;
      const float l1 = rs.x * kLog2e, l2 = rs.y * kLog2e, ldf = rs.z * kLog2e;
      const float wu = a.w_uni, cj1 = a.w_joint * c1, cj2 = a.w_joint * c2;      // loss-term ablations (1 or 0)
      float d1[NE], d2[NE];
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int c = col_of<G, NK>(l, i);
        const float oh = (c == y) ? 1.f : 0.f;
        const float p1 = exp_sub(r1.v[i], l1), p2 = exp_sub(r2.v[i], l2);
        const float pd = exp_sub(r1.v[i] * c1 + r2.v[i] * c2, ldf) - oh;
        d1[i] = (wu * (p1 - oh) + cj1 * pd) * invB + g1 * p1;
        d2[i] = (wu * (p2 - oh) + cj2 * pd) * invB + g2 * p2;
      }
      mask_cols<G, NK>(d1, l, C, 0.f);                   // padded columns hold z = 0: their dz is not a gradient of anything
      mask_cols<G, NK>(d2, l, C, 0.f);
      if (valid) {
        store_dz_chunks<G, NK>(a, 0, (size_t)b, d1, l, C);
        store_dz_chunks<G, NK>(a, 1, (size_t)b, d2, l, C);
#pragma unroll
        for (int i = 0; i < NE; ++i) { d1s[i] += d1[i]; d2s[i] += d2[i]; }
      }
    }
    float w1[NE], w2[NE];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const float4 o1 = *reinterpret_cast<const float4*>(soff + 4 * (l + G * k));
      const float4 o2 = *reinterpret_cast<const float4*>(soff + LDP + 4 * (l + G * k));
      w1[4 * k] = r1.v[4 * k] + o1.x; w1[4 * k + 1] = r1.v[4 * k + 1] + o1.y; w1[4 * k + 2] = r1.v[4 * k + 2] + o1.z; w1[4 * k + 3] = r1.v[4 * k + 3] + o1.w;
      w2[4 * k] = r2.v[4 * k] + o2.x; w2[4 * k + 1] = r2.v[4 * k + 1] + o2.y; w2[4 * k + 2] = r2.v[4 * k + 2] + o2.z; w2[4 * k + 3] = r2.v[4 * k + 3] + o2.w;
    }
    float m; int i1, i2;
    row_max_arg<G, NK>(w1, l, m, i1);
    row_max_arg<G, NK>(w2, l, m, i2);
    if (valid) { cal1 += (i1 == y); cal2 += (i2 == y); }
  }
  cal1 = across_groups_sum<G>(cal1); cal2 = across_groups_sum<G>(cal2);
  __shared__ float s1[8], s2[8];
  if (lane == 0) { s1[warp] = cal1; s2[warp] = cal2; }
  if (MODE == LF_MODE_QMF) {
#pragma unroll
    for (int i = 0; i < NE; ++i) { d1s[i] = across_groups_sum<G>(d1s[i]); d2s[i] = across_groups_sum<G>(d2s[i]); }
    float* dsum = smem + (size_t)warp * 2 * C;
    if (gi == 0) {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int c = col_of<G, NK>(l, i);
        if (c < C) { dsum[c] = d1s[i]; dsum[C + c] = d2s[i]; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += threadIdx.x == 0 ? s1[w] : s2[w];
    a.calpart[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
  if (MODE == LF_MODE_QMF)
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += smem[(size_t)w * 2 * C + c];
      a.dbpart[(size_t)blockIdx.x * 2 * C + c] = s;
    }
  for (int r = blockIdx.x + gridDim.x; r < a.nb_total; r += gridDim.x) {     // rows no CTA owns
    if (threadIdx.x < 2) a.calpart[(size_t)r * 2 + threadIdx.x] = 0.f;
    if (MODE == LF_MODE_QMF)
      for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) a.dbpart[(size_t)r * 2 * C + c] = 0.f;
  }
}

// one full wave: never more CTAs than are resident at once (a partial second wave costs a whole pass)
template <class K>
static int one_wave_vec(K kernel, size_t smem, int nb) {
  static int occ = 0;                      // per template instantiation
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 1;
  }
  const int cap = 148 * occ;
  return nb < cap ? nb : cap;
}

template <int MODE, int G, int NK>
static int launch_fwd_vec(const RowsArgs& a, int nb, cudaStream_t s) {
  const size_t sm = (size_t)8 * (3 * a.C + 9) * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(rows_forward_vec_kernel<MODE, G, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  nb = one_wave_vec(rows_forward_vec_kernel<MODE, G, NK>, sm, nb);
  LF_LAUNCH(MODE == LF_MODE_QMF ? "rows_forward_qmf" : "rows_forward_jlogits", s,
            launch_pdl(rows_forward_vec_kernel<MODE, G, NK>, dim3(nb), dim3(256), sm, s, a));
  return check_launch("rows_forward_vec_kernel");
}
template <int MODE, int G, int NK>
static int launch_bwd_vec(const RowsArgs& a, int nb, cudaStream_t s) {
  const size_t sm = ((MODE == LF_MODE_QMF ? (size_t)8 * 2 * a.C : 0) + 2 * 4 * G * NK) * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(rows_backward_vec_kernel<MODE, G, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  nb = one_wave_vec(rows_backward_vec_kernel<MODE, G, NK>, sm, nb);
  LF_LAUNCH(MODE == LF_MODE_QMF ? "rows_backward_qmf" : "rows_calibrated", s,
            launch_pdl(rows_backward_vec_kernel<MODE, G, NK>, dim3(nb), dim3(256), sm, s, a));
  return check_launch("rows_backward_vec_kernel");
}

#define LF_DISPATCH_VEC(FN, MODE_)                                    \
  do {                                                                \
    const int nq = a.ld_z / 4;                                        \
    if (nq <= 16) return FN<MODE_, 8, 2>(a, nb, s);                   \
    if (nq <= 32) return FN<MODE_, 8, 4>(a, nb, s);                   \
    return FN<MODE_, 16, 4>(a, nb, s);                                \
  } while (0)

// 16-byte row pitch and base alignment of z (and of dz where this pass writes it); C <= 512.
bool rows_vec_supported(const RowsArgs& a, bool writes_dz) {
  // up to 256 classes (G = 8 or 16: several samples per warp).  Above that a warp holds one sample either way and
  // the register-prefetching one-warp-per-sample kernels measured ~10 % faster (K5, C = 309: 133 vs 149 us).
  if (a.C > 256 || a.ld_z % 4 || a.ld_z < a.C || a.ld_z > 256) return false;
  if (((uintptr_t)a.z[0] | (uintptr_t)a.z[1]) & 15) return false;
  if (a.rowstat && ((uintptr_t)a.rowstat & 15)) return false;
  if (a.ld_f % 4 == 0 && ((((uintptr_t)a.avg) | ((uintptr_t)a.zdf)) & 15)) return false;
  if (writes_dz) {
    if (a.ldz % 4 || a.ldz < a.C) return false;
    const uintptr_t al = a.dz_bf16 ? 7 : 15;
    if (((uintptr_t)a.dz[0] & al) || (a.dz[1] && ((uintptr_t)a.dz[1] & al))) return false;
  }
  return true;
}
int rows_forward_vec(const RowsArgs& a, int mode, int nb, cudaStream_t s) {
  if (mode == LF_MODE_QMF) LF_DISPATCH_VEC(launch_fwd_vec, LF_MODE_QMF);
  LF_DISPATCH_VEC(launch_fwd_vec, LF_MODE_JLOGITS);
}
int rows_backward_vec(const RowsArgs& a, int mode, int nb, cudaStream_t s) {
  if (mode == LF_MODE_QMF) LF_DISPATCH_VEC(launch_bwd_vec, LF_MODE_QMF);
  LF_DISPATCH_VEC(launch_bwd_vec, LF_MODE_JLOGITS);
}

}  // namespace lf
