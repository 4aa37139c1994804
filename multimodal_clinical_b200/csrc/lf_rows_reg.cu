// Register-resident per-sample ("row") kernels for C <= 512: one warp per sample, lane = class, the whole
// row of both heads' logits lives in registers (NCH = ceil(C/32) values per lane per head), so every
// statistic of the step is computed from ONE coalesced read of z1/z2:
//   forward : avg, QMF energy/conf/z_df, CE terms, OGM-GE scores, argmax counts, EMA column sums,
//             (JLOGITS) dL/dz and its column sums
//   backward: QMF dL/dz_m and column sums (db), calibrated argmax counts (z_m + ema_offset_m)
// Reductions over classes: maxima and first-index argmax go through the integer REDUX unit on an
// order-preserving float->uint map (one instruction each); sums are xor-butterflies (fixed order, so the
// step stays bit-reproducible).  The next sample's logits are loaded while the current one is reduced.
// Column sums accumulate in registers (a lane owns the same classes for every sample) and are combined
// across the CTA's warps once, at the end.
//
// Reference arithmetic: see lf_rows.cu (same formulas; this file only changes the mapping to the machine).
#include "lf_common.cuh"
#include "lf_rows.cuh"
#include "lf_rowmath.cuh"

namespace lf {

template <int NCH>
__device__ __forceinline__ void load_row(const float* __restrict__ z1, const float* __restrict__ z2, size_t off, int C,
                                         int lane, float (&v1)[NCH], float (&v2)[NCH]) {
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = lane + 32 * k;
    v1[k] = c < C ? z1[off + c] : -INFINITY;
    v2[k] = c < C ? z2[off + c] : -INFINITY;
  }
}

template <int MODE, int NCH>
__global__ void __launch_bounds__(256) rows_forward_reg_kernel(RowsArgs a) {
  extern __shared__ float smem[];                       // [8 warps][3][C] column sums, then [8][9] stats
  const int C = a.C, B = a.B;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  const int stride = gridDim.x * nwarp;
  const float dz_scale = 0.5f / (float)a.B_global;
  float cs1[NCH], cs2[NCH], cs3[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) { cs1[k] = 0.f; cs2[k] = 0.f; cs3[k] = 0.f; }
  float st[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.f;

  int b = blockIdx.x * nwarp + warp;
  float n1[NCH], n2[NCH];
  int ny = 0;
  if (b < B) { load_row<NCH>(a.z[0], a.z[1], (size_t)b * a.ld_z, C, lane, n1, n2); ny = (int)a.label[b]; }
  for (; b < B; b += stride) {
    float v1[NCH], v2[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { v1[k] = n1[k]; v2[k] = n2[k]; }
    const int y = ny;
    if (b + stride < B) { load_row<NCH>(a.z[0], a.z[1], (size_t)(b + stride) * a.ld_z, C, lane, n1, n2); ny = (int)a.label[b + stride]; }

    float av[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      av[k] = (v1[k] + v2[k]) / 2.f;
      const int c = lane + 32 * k;
      if (c < C) { a.avg[(size_t)b * a.ld_f + c] = av[k]; cs1[k] += v1[k]; cs2[k] += v2[k]; }
    }
    float m1, m2, ma; int i1, i2, ia;
    warp_max_arg<NCH>(v1, lane, m1, i1);
    warp_max_arg<NCH>(v2, lane, m2, i2);
    warp_max_arg<NCH>(av, lane, ma, ia);
    float s1 = 0.f, s2 = 0.f, sa = 0.f;
#pragma unroll
    {
      const float k1 = m1 * 1.4426950408889634f, k2 = m2 * 1.4426950408889634f, ka = ma * 1.4426950408889634f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) { s1 += exp_sub(v1[k], k1); s2 += exp_sub(v2[k], k2); sa += exp_sub(av[k], ka); }
    }
    warp_sum3(s1, s2, sa);
    const float lse1 = m1 + __logf(s1), lse2 = m2 + __logf(s2), lsea = ma + __logf(sa);
    const bool yok = (unsigned)y < (unsigned)C;
    const float py1 = pick_class<NCH>(v1, y), py2 = pick_class<NCH>(v2, y);
    const float zy1 = yok ? py1 : 0.f, zy2 = yok ? py2 : 0.f;

    float ce_joint;
    int cnt_df = 0;
    if (MODE == LF_MODE_QMF) {
      // energy = log(sum(exp z)) is NOT stabilised in the reference (QMF.py:113): same value unless the
      // plain fp32 sum overflows, where the reference yields +inf
      const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;
      const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;
      float vd[NCH];
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        vd[k] = c < C ? v1[k] * c1 + v2[k] * c2 : -INFINITY;
        if (c < C) a.zdf[(size_t)b * a.ld_f + c] = vd[k];
      }
      float md; int idf;
      warp_max_arg<NCH>(vd, lane, md, idf);
      float sd = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) sd += exp_sub(vd[k], md * 1.4426950408889634f);
      sd = warp_sum(sd);
      const float lsed = md + __logf(sd);
      const float pyd = pick_class<NCH>(vd, y);
      const float zyd = yok ? pyd : 0.f;
      ce_joint = lsed - zyd;
      cnt_df = (idf == y);
      if (lane == 0) {
        a.conf[b] = c1;
        a.conf[B + b] = c2;
        *reinterpret_cast<float4*>(a.rowstat + (size_t)b * 4) = make_float4(lse1, lse2, lsed, 0.f);
      }
    } else {
      ce_joint = lsea - 0.5f * (zy1 + zy2);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {          // dL/dz1 = dL/dz2 = (softmax(avg) - onehot) / (2 Bg)
        const int c = lane + 32 * k;
        if (c < C) {
          const float d = (exp_sub(av[k], lsea * 1.4426950408889634f) - (c == y ? 1.f : 0.f)) * dz_scale;
          store_dz(a, 0, (size_t)b * a.ldz + c, d);
          cs3[k] += d;
        }
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
  }

  // ---- CTA reduction in fixed order -> one partial row per CTA
  float* colsum = smem + (size_t)warp * 3 * C;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = lane + 32 * k;
    if (c < C) { colsum[c] = cs1[k]; colsum[C + c] = cs2[k]; colsum[2 * C + c] = cs3[k]; }
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

template <int MODE, int NCH>
__global__ void __launch_bounds__(256) rows_backward_reg_kernel(RowsArgs a) {
  extern __shared__ float smem[];                       // [8 warps][2][C] column sums of dz (QMF)
  const int C = a.C, B = a.B;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  const int stride = gridDim.x * nwarp;
  const float invB = 1.f / (float)a.B_global;
  float off1[NCH], off2[NCH], d1s[NCH], d2s[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = lane + 32 * k;
    off1[k] = c < C ? a.ema_off[c] : 0.f;
    off2[k] = c < C ? a.ema_off[C + c] : 0.f;
    d1s[k] = 0.f; d2s[k] = 0.f;
  }
  float cal1 = 0.f, cal2 = 0.f;

  int b = blockIdx.x * nwarp + warp;
  float n1[NCH], n2[NCH];
  // the next sample's logits AND its per-row scalars are fetched while the current sample is processed
  int ny = 0;
  float nc1 = 0.f, nc2 = 0.f, ng1 = 0.f, ng2 = 0.f;
  float4 nrs = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch_scalars = [&](int r) {
    ny = (int)a.label[r];
    if (MODE == LF_MODE_QMF) {
      nc1 = a.conf[r]; nc2 = a.conf[B + r];
      nrs = *reinterpret_cast<const float4*>(a.rowstat + (size_t)r * 4);               // lse1, lse2, lse(z_df)
      ng1 = a.qmf_g[r]; ng2 = a.qmf_g[B + r];
    }
  };
  if (b < B) { load_row<NCH>(a.z[0], a.z[1], (size_t)b * a.ld_z, C, lane, n1, n2); fetch_scalars(b); }
  for (; b < B; b += stride) {
    float v1[NCH], v2[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { v1[k] = n1[k]; v2[k] = n2[k]; }
    const int y = ny;
    const float c1 = nc1, c2 = nc2, g1 = ng1 / 10.f, g2 = ng2 / 10.f;
    const float4 rs = nrs;
    if (b + stride < B) { load_row<NCH>(a.z[0], a.z[1], (size_t)(b + stride) * a.ld_z, C, lane, n1, n2); fetch_scalars(b + stride); }
    if (MODE == LF_MODE_QMF) {
      const float l1 = rs.x * 1.4426950408889634f, l2 = rs.y * 1.4426950408889634f, ld = rs.z * 1.4426950408889634f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
          const float oh = (c == y) ? 1.f : 0.f;
          const float p1 = exp_sub(v1[k], l1), p2 = exp_sub(v2[k], l2);
          const float pd = exp_sub(v1[k] * c1 + v2[k] * c2, ld) - oh;
          const float d1 = (a.w_uni * (p1 - oh) + (a.w_joint * c1) * pd) * invB + g1 * p1;
          const float d2 = (a.w_uni * (p2 - oh) + (a.w_joint * c2) * pd) * invB + g2 * p2;
          store_dz(a, 0, (size_t)b * a.ldz + c, d1);
          store_dz(a, 1, (size_t)b * a.ldz + c, d2);
          d1s[k] += d1; d2s[k] += d2;
        }
      }
    }
    float w1[NCH], w2[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { w1[k] = v1[k] + off1[k]; w2[k] = v2[k] + off2[k]; }
    float m; int i1, i2;
    warp_max_arg<NCH>(w1, lane, m, i1);
    warp_max_arg<NCH>(w2, lane, m, i2);
    cal1 += (i1 == y); cal2 += (i2 == y);
  }
  __shared__ float s1[8], s2[8];
  if (lane == 0) { s1[warp] = cal1; s2[warp] = cal2; }
  if (MODE == LF_MODE_QMF) {
    float* dsum = smem + (size_t)warp * 2 * C;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = lane + 32 * k;
      if (c < C) { dsum[c] = d1s[k]; dsum[C + c] = d2s[k]; }
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
static int one_wave(K kernel, size_t smem, int nb) {
  static int occ = 0;                      // per template instantiation
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 1;
  }
  const int cap = 148 * occ;
  return nb < cap ? nb : cap;
}

template <int MODE, int NCH>
static int launch_fwd(const RowsArgs& a, int nb, cudaStream_t s) {
  const size_t sm = (size_t)8 * (3 * a.C + 9) * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(rows_forward_reg_kernel<MODE, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  nb = one_wave(rows_forward_reg_kernel<MODE, NCH>, sm, nb);
  LF_LAUNCH(MODE == LF_MODE_QMF ? "rows_forward_qmf" : "rows_forward_jlogits", s,
            (rows_forward_reg_kernel<MODE, NCH><<<nb, 256, sm, s>>>(a)));
  return check_launch("rows_forward_reg_kernel");
}
template <int MODE, int NCH>
static int launch_bwd(const RowsArgs& a, int nb, cudaStream_t s) {
  const size_t sm = MODE == LF_MODE_QMF ? (size_t)8 * 2 * a.C * sizeof(float) : 0;
  if (sm > 48 * 1024) cudaFuncSetAttribute(rows_backward_reg_kernel<MODE, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  nb = one_wave(rows_backward_reg_kernel<MODE, NCH>, sm, nb);
  LF_LAUNCH(MODE == LF_MODE_QMF ? "rows_backward_qmf" : "rows_calibrated", s,
            (rows_backward_reg_kernel<MODE, NCH><<<nb, 256, sm, s>>>(a)));
  return check_launch("rows_backward_reg_kernel");
}

#define LF_DISPATCH_NCH(FN, MODE_)                                   \
  do {                                                               \
    const int nch = div_up(a.C, 32);                                 \
    if (nch <= 1) return FN<MODE_, 1>(a, nb, s);                     \
    if (nch <= 2) return FN<MODE_, 2>(a, nb, s);                     \
    if (nch <= 4) return FN<MODE_, 4>(a, nb, s);                     \
    if (nch <= 7) return FN<MODE_, 7>(a, nb, s);                     \
    if (nch <= 10) return FN<MODE_, 10>(a, nb, s);                   \
    return FN<MODE_, 16>(a, nb, s);                                  \
  } while (0)

// C <= 512.  Same outputs / partial layouts as the generic kernels in lf_rows.cu.
bool rows_reg_supported(int C) { return C <= 512; }
int rows_forward_reg(const RowsArgs& a, int mode, int nb, cudaStream_t s) {
  if (mode == LF_MODE_QMF) LF_DISPATCH_NCH(launch_fwd, LF_MODE_QMF);
  LF_DISPATCH_NCH(launch_fwd, LF_MODE_JLOGITS);
}
int rows_backward_reg(const RowsArgs& a, int mode, int nb, cudaStream_t s) {
  if (mode == LF_MODE_QMF) LF_DISPATCH_NCH(launch_bwd, LF_MODE_QMF);
  LF_DISPATCH_NCH(launch_bwd, LF_MODE_JLOGITS);
}

}  // namespace lf
