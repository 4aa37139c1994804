// Per-sample ("row") math of the late-fusion step on logits already in HBM:
// softmax / log-sum-exp statistics, cross-entropy terms, QMF energy confidence and fused logits,
// OGM-GE scores, accuracy counts, EMA logit sums, and dL/dlogits.   One warp per sample, lanes
// strided over classes (coalesced for any C), online log-sum-exp, warp-shuffle reductions,
// deterministic two-stage reduction of the batch statistics.
//
// Reference arithmetic restated here (paths relative to the reference tree):
//   avg = (z1+z2)/2, CE(avg)                cremad/joint_model_ogm_ge.py:54-56
//   energy/conf/z_df                        existing_algos/QMF.py:113-117
//   CE(z_m), CE(z_df)                       cremad/joint_model_qmf.py:64,68
//   scores                                  existing_algos/OGM_GE.py:21-22
//   argmax accuracies                       utils/BaseModel.py:78-92, 961
//   dL/dz (analytic)
#include "lf_common.cuh"
#include "lf_rows.cuh"

namespace lf {

struct Lse {  // online log-sum-exp accumulator
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add(float v) {
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }
    else s += __expf(v - m);
  }
  // warp-combine; result (same in all lanes) = log sum exp over everything added by the warp
  __device__ __forceinline__ float finish() {
    const float M = warp_max(m);
    const float part = (m == -INFINITY) ? 0.f : s * __expf(m - M);
    return M + logf(warp_sum(part));
  }
};

// NOTE on intrinsics: __expf/__logf are the fast SFU forms (abs err ~2 ulp in the ranges that occur
// here); the fp32 parity budget is 1e-5 relative and tests/test_parity_gpu.py checks it holds.
// The QMF energy deliberately uses the NON-stabilised sum of exp like the reference (QMF.py:113).

template <int MODE>
__global__ void __launch_bounds__(256) rows_forward_kernel(RowsArgs a) {
  extern __shared__ float smem[];
  const int C = a.C, B = a.B;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  float* colsum = smem + (size_t)warp * 3 * C;  // [z1 | z2 | dz][C] owned by this warp
  for (int c = lane; c < 3 * C; c += 32) colsum[c] = 0.f;
  __syncwarp();

  float st[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.f;

  const float dz_scale = 0.5f / (float)a.B_global;

  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    const float* __restrict__ z1 = a.z[0] + (size_t)b * a.ld_z;
    const float* __restrict__ z2 = a.z[1] + (size_t)b * a.ld_z;
    const int y = (int)a.label[b];
    Lse l1, l2, la;
    l1.init(); l2.init(); la.init();
    float e1 = 0.f, e2 = 0.f;                       // plain sum exp (QMF energy)
    float m1 = -INFINITY, m2 = -INFINITY, ma = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff, ia = 0x7fffffff;
    float zy1 = 0.f, zy2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v1 = z1[c], v2 = z2[c];
      const float av = (v1 + v2) / 2.f;
      a.avg[(size_t)b * a.ld_f + c] = av;
      l1.add(v1); l2.add(v2); la.add(av);
      if (MODE == LF_MODE_QMF) { e1 += expf(v1); e2 += expf(v2); }
      if (v1 > m1) { m1 = v1; i1 = c; }
      if (v2 > m2) { m2 = v2; i2 = c; }
      if (av > ma) { ma = av; ia = c; }
      if (c == y) { zy1 = v1; zy2 = v2; }
      colsum[c] += v1;
      colsum[C + c] += v2;
    }
    const float lse1 = l1.finish(), lse2 = l2.finish(), lsea = la.finish();
    warp_argmax(m1, i1); warp_argmax(m2, i2); warp_argmax(ma, ia);
    zy1 = __shfl_sync(kFull, zy1, y & 31);
    zy2 = __shfl_sync(kFull, zy2, y & 31);

    float ce_joint, lse_joint;
    int cnt_df = 0;
    if (MODE == LF_MODE_QMF) {
      const float c1 = logf(warp_sum(e1)) / 10.f;
      const float c2 = logf(warp_sum(e2)) / 10.f;
      Lse ld; ld.init();
      float md = -INFINITY; int idf = 0x7fffffff; float zyd = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float vd = z1[c] * c1 + z2[c] * c2;
        a.zdf[(size_t)b * a.ld_f + c] = vd;
        ld.add(vd);
        if (vd > md) { md = vd; idf = c; }
        if (c == y) zyd = vd;
      }
      lse_joint = ld.finish();
      warp_argmax(md, idf);
      zyd = __shfl_sync(kFull, zyd, y & 31);
      ce_joint = lse_joint - zyd;
      cnt_df = (idf == y);
      if (lane == 0) {
        a.conf[b] = c1;
        a.conf[B + b] = c2;
        a.rowstat[(size_t)b * 4 + 0] = lse1;
        a.rowstat[(size_t)b * 4 + 1] = lse2;
        a.rowstat[(size_t)b * 4 + 2] = lse_joint;
      }
    } else {
      lse_joint = lsea;
      ce_joint = lsea - 0.5f * (zy1 + zy2);
      // dL/dz1 = dL/dz2 = (softmax(avg) - onehot) / (2 Bg)
      for (int c = lane; c < C; c += 32) {
        const float av = (z1[c] + z2[c]) / 2.f;
        const float p = __expf(av - lsea);
        const float d = (p - (c == y ? 1.f : 0.f)) * dz_scale;
        store_dz(a, 0, (size_t)b * a.ldz + c, d);
        colsum[2 * C + c] += d;
      }
    }
    if (lane == 0) {
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

  // ---- block reduction in fixed order -> one partial row per block
  __shared__ float sst[8][9];
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 9; ++i) sst[warp][i] = st[i];
  __syncthreads();
  float* out = a.partials + (size_t)blockIdx.x * stat_len_dev(C);
  if (threadIdx.x < LF_STATS_HEADER) {
    float s = 0.f;
    if (threadIdx.x < 9)
      for (int w = 0; w < nwarp; ++w) s += sst[w][threadIdx.x];
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
}

// stats[i] = sum over blocks (fp64, fixed order).  Entries [lo, hi) plus, when with_cols, the 2C tail.
// One CTA per 32 statistics: 32 columns x 8 row groups, coalesced 128-byte reads, fixed summation order.
__global__ void __launch_bounds__(1024) finalize_stats_kernel(const float* __restrict__ partials, int nblocks, int len, int lo,
                                                              int hi, int with_cols, double* __restrict__ stats) {
  // 8 columns x 128 row groups: with <= 640 partial rows every thread's loads are issued at once (one memory round
  // trip for the whole kernel); fixed-order two-level sum over the groups
  const int tx = threadIdx.x % 8, ty = threadIdx.x / 8;
  const int i = blockIdx.x * 8 + tx;
  pdl_wait();
  pdl_trigger();
  double s = 0.0;
  if (i < len) {
    int b = ty;
    for (; b + 512 < nblocks; b += 640) {
      float v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) v[u] = partials[(size_t)(b + 128 * u) * len + i];
#pragma unroll
      for (int u = 0; u < 5; ++u) s += (double)v[u];
    }
    for (; b < nblocks; b += 128) s += (double)partials[(size_t)b * len + i];
  }
  __shared__ double sm[128][9];
  __shared__ double sm2[8][9];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty < 8) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += sm[ty * 16 + w][tx];
    sm2[ty][tx] = t;
  }
  __syncthreads();
  if (ty != 0 || i >= len) return;
  const bool header = i < LF_STATS_HEADER;
  if (header && (i < lo || i >= hi)) return;
  if (!header && !with_cols) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sm2[w][tx];
  stats[i] = t;
}

// Backward rows: QMF dL/dz_m and, for both modes, the calibrated counts
// argmax(z_m + offset_m) == y (utils/BaseModel.py:84-89).
template <int MODE>
__global__ void __launch_bounds__(256) rows_backward_kernel(RowsArgs a) {
  extern __shared__ float smem[];
  const int C = a.C, B = a.B;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nwarp = blockDim.x / 32;
  float* dsum = smem + (size_t)warp * 2 * C;     // [2][C] column sums of dz owned by this warp (QMF)
  if (MODE == LF_MODE_QMF) {
    for (int c = lane; c < 2 * C; c += 32) dsum[c] = 0.f;
    __syncwarp();
  }
  const float invB = 1.f / (float)a.B_global;
  float cal1 = 0.f, cal2 = 0.f;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    const float* __restrict__ z1 = a.z[0] + (size_t)b * a.ld_z;
    const float* __restrict__ z2 = a.z[1] + (size_t)b * a.ld_z;
    const int y = (int)a.label[b];
    float c1 = 0.f, c2 = 0.f, lse1 = 0.f, lse2 = 0.f, lsed = 0.f, g1 = 0.f, g2 = 0.f;
    if (MODE == LF_MODE_QMF) {
      c1 = a.conf[b]; c2 = a.conf[B + b];
      lse1 = a.rowstat[(size_t)b * 4 + 0];
      lse2 = a.rowstat[(size_t)b * 4 + 1];
      lsed = a.rowstat[(size_t)b * 4 + 2];
      g1 = a.qmf_g[b] / 10.f;
      g2 = a.qmf_g[B + b] / 10.f;
    }
    float m1 = -INFINITY, m2 = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v1 = z1[c], v2 = z2[c];
      const float w1 = v1 + a.ema_off[c], w2 = v2 + a.ema_off[C + c];
      if (w1 > m1) { m1 = w1; i1 = c; }
      if (w2 > m2) { m2 = w2; i2 = c; }
      if (MODE == LF_MODE_QMF) {
        const float oh = (c == y) ? 1.f : 0.f;
        const float p1 = __expf(v1 - lse1), p2 = __expf(v2 - lse2);
        const float pd = __expf((v1 * c1 + v2 * c2) - lsed) - oh;
        const float d1 = (a.w_uni * (p1 - oh) + (a.w_joint * c1) * pd) * invB + g1 * p1;
        const float d2 = (a.w_uni * (p2 - oh) + (a.w_joint * c2) * pd) * invB + g2 * p2;
        store_dz(a, 0, (size_t)b * a.ldz + c, d1);
        store_dz(a, 1, (size_t)b * a.ldz + c, d2);
        dsum[c] += d1;
        dsum[C + c] += d2;
      }
    }
    warp_argmax(m1, i1); warp_argmax(m2, i2);
    cal1 += (i1 == y); cal2 += (i2 == y);
  }
  __shared__ float s1[8], s2[8];
  if (lane == 0) { s1[warp] = cal1; s2[warp] = cal2; }
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
}

// dbias / calibrated counts: 32 columns x 8 row groups per CTA, fixed summation order
__global__ void __launch_bounds__(1024) finalize_db_cal_kernel(const float* __restrict__ dbpart, int nb_db, int C,
                                                               const float* __restrict__ calpart, int nb_cal,
                                                               float* __restrict__ db0, float* __restrict__ db1,
                                                               double* __restrict__ stats) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;      // 32 columns x 32 row groups
  const int i = blockIdx.x * 32 + tx;             // [0, 2C): db columns; [2C, 2C+2): calibrated counts
  const bool is_db = i < 2 * C;
  const float* __restrict__ src = is_db ? dbpart + i : calpart + (i - 2 * C);
  const size_t pitch = is_db ? (size_t)2 * C : 2;
  const int nb = is_db ? nb_db : nb_cal;
  double s = 0.0;
  if (i < 2 * C + 2) {
    int b = ty;
    for (; b + 96 < nb; b += 128) {
      const float v0 = src[(size_t)b * pitch], v1 = src[(size_t)(b + 32) * pitch];
      const float v2 = src[(size_t)(b + 64) * pitch], v3 = src[(size_t)(b + 96) * pitch];
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; b < nb; b += 32) s += (double)src[(size_t)b * pitch];
  }
  __shared__ double sm[32][33];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || i >= 2 * C + 2) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 32; ++w) t += sm[w][tx];
  if (i < C) db0[i] = (float)t;
  else if (i < 2 * C) db1[i - C] = (float)t;
  else stats[LF_STAT_CNT_X1_CAL + (i - 2 * C)] = t;
}

int finalize_db_cal(const float* dbpart, int nb_db, int C, const float* calpart, int nb_cal, float* db0, float* db1,
                    double* stats, cudaStream_t s) {
  LF_LAUNCH("finalize_db_cal", s, (finalize_db_cal_kernel<<<div_up(2 * C + 2, 32), 1024, 0, s>>>(dbpart, nb_db, C, calpart, nb_cal, db0, db1, stats)));
  return check_launch("finalize_db_cal");
}

// The tail of the backward pass in ONE launch: dW_m = fixed-order sum of the split-K partials (CTAs [0, 2 nblk))
// and db_m / calibrated counts = column sums of the per-CTA partials (remaining CTAs; 8 columns x 32 row groups).
__global__ void __launch_bounds__(256) finalize_grads_kernel(const float* __restrict__ part, float* __restrict__ dw0,
                                                             float* __restrict__ dw1, int splits, int max_splits, size_t n,
                                                             int nblk, const float* __restrict__ dbpart, int nb_db, int C,
                                                             const float* __restrict__ calpart, int nb_cal,
                                                             float* __restrict__ db0, float* __restrict__ db1,
                                                             double* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  if ((int)blockIdx.x < 2 * nblk) {
    // 32 float4 columns x 8 split groups: every thread's loads are independent (one DRAM round trip instead of a
    // chain of `splits`), then a fixed-order sum over the groups
    const int m = (int)blockIdx.x / nblk, blk = (int)blockIdx.x - m * nblk;
    const size_t i = ((size_t)blk * 32 + tx) * 4;                     // n is a multiple of 4 (D % 4 == 0)
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      const float* p = part + (size_t)m * max_splits * n + i;
#pragma unroll 4
      for (int k = ty; k < splits; k += 8) {
        const float4 v = *reinterpret_cast<const float4*>(p + (size_t)k * n);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    __shared__ float4 sm4[8][33];
    sm4[ty][tx] = s;
    __syncthreads();
    if (ty != 0 || i >= n) return;
    float4 t = sm4[0][tx];
#pragma unroll
    for (int g = 1; g < 8; ++g) { const float4 v = sm4[g][tx]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    float* o = (m == 0 ? dw0 : dw1) + i;           // the flat gradient buffer packs dW2 after db1: not always 16-byte aligned
    if (((uintptr_t)o & 15) == 0) *reinterpret_cast<float4*>(o) = t;
    else { o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
    return;
  }
  // db / calibrated counts: 8 columns x 32 row groups per CTA, eight independent loads in flight per thread
  const int cx = threadIdx.x % 8, ry = threadIdx.x / 8;
  const int i = ((int)blockIdx.x - 2 * nblk) * 8 + cx;              // [0, 2C): db columns; [2C, 2C+2): calibrated counts
  const bool is_db = i < 2 * C;
  const float* __restrict__ src = is_db ? dbpart + i : calpart + (i - 2 * C);
  const size_t pitch = is_db ? (size_t)2 * C : 2;
  const int nb = is_db ? nb_db : nb_cal;
  double s = 0.0;
  if (i < 2 * C + 2) {
    int b = ry;
    for (; b + 224 < nb; b += 256) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(b + 32 * u) * pitch];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; b < nb; b += 32) s += (double)src[(size_t)b * pitch];
  }
  __shared__ double sm[32][9];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry != 0 || i >= 2 * C + 2) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 32; ++w) t += sm[w][cx];
  if (i < C) db0[i] = (float)t;
  else if (i < 2 * C) db1[i - C] = (float)t;
  else stats[LF_STAT_CNT_X1_CAL + (i - 2 * C)] = t;
}

bool finalize_grads_supported(const float* part, const float* dw0, const float* dw1, int splits, size_t n) {
  (void)dw0; (void)dw1;
  return splits <= 32 && n % 4 == 0 && (((uintptr_t)part) & 15) == 0;
}
int finalize_grads(const float* part, float* dw0, float* dw1, int splits, int max_splits, size_t n, const float* dbpart,
                   int nb_db, int C, const float* calpart, int nb_cal, float* db0, float* db1, double* stats, cudaStream_t s) {
  const int nblk = div_up((long long)(n / 4), 32);
  LF_LAUNCH("finalize_grads", s, launch_pdl(finalize_grads_kernel, dim3(2 * nblk + div_up(2 * C + 2, 8)), dim3(256), 0, s,
      part, dw0, dw1, splits, max_splits, n, nblk, dbpart, nb_db, C, calpart, nb_cal, db0, db1, stats));
  return check_launch("finalize_grads");
}

int row_blocks(int B) {
  int nb = div_up(B, 8);
  return nb < kMaxRowBlocks ? (nb < 1 ? 1 : nb) : kMaxRowBlocks;
}

bool rows_reg_supported(int C);                                                   // lf_rows_reg.cu
int rows_forward_reg(const RowsArgs& a, int mode, int nb, cudaStream_t s);
int rows_backward_reg(const RowsArgs& a, int mode, int nb, cudaStream_t s);
bool rows_vec_supported(const RowsArgs& a, bool writes_dz);                       // lf_rows_vec.cu
int rows_forward_vec(const RowsArgs& a, int mode, int nb, cudaStream_t s);
int rows_backward_vec(const RowsArgs& a, int mode, int nb, cudaStream_t s);

int rows_forward(const RowsArgs& a, int mode, cudaStream_t s) {
  const int nb = row_blocks(a.B);
  if (rows_reg_supported(a.C)) {
    // 16-byte pitched logits (tensor-pipe path): G lanes per sample, 128-bit accesses; else one warp per sample
    int rc = rows_vec_supported(a, mode == LF_MODE_JLOGITS) ? rows_forward_vec(a, mode, nb, s) : rows_forward_reg(a, mode, nb, s);
    if (rc) return rc;
    LF_LAUNCH("finalize_stats", s, launch_pdl(finalize_stats_kernel, dim3(div_up(stat_len(a.C), 8)), dim3(1024), 0, s, (const float*)a.partials, nb, stat_len(a.C), 0, LF_STATS_HEADER, 1, a.stats));
    return check_launch("finalize_stats_kernel");
  }
  const size_t sm = (size_t)8 * 3 * a.C * sizeof(float);
  if (sm > 200 * 1024) { set_error("classes=%d too wide for rows_forward shared memory", a.C); return LF_ERR_UNSUPPORTED; }
  if (mode == LF_MODE_QMF) {
    if (sm > 48 * 1024) cudaFuncSetAttribute(rows_forward_kernel<LF_MODE_QMF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    LF_LAUNCH("rows_forward_qmf", s, (rows_forward_kernel<LF_MODE_QMF><<<nb, 256, sm, s>>>(a)));
  } else {
    if (sm > 48 * 1024) cudaFuncSetAttribute(rows_forward_kernel<LF_MODE_JLOGITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    LF_LAUNCH("rows_forward_jlogits", s, (rows_forward_kernel<LF_MODE_JLOGITS><<<nb, 256, sm, s>>>(a)));
  }
  int rc = check_launch("rows_forward_kernel");
  if (rc) return rc;
  LF_LAUNCH("finalize_stats", s, launch_pdl(finalize_stats_kernel, dim3(div_up(stat_len(a.C), 8)), dim3(1024), 0, s, (const float*)a.partials, nb, stat_len(a.C), 0, LF_STATS_HEADER, 1, a.stats));
  return check_launch("finalize_stats_kernel");
}

void finalize_forward_stats(const float* partials, int nblocks, int C, double* stats, cudaStream_t s) {
  LF_LAUNCH("finalize_stats", s, launch_pdl(finalize_stats_kernel, dim3(div_up(stat_len(C), 8)), dim3(1024), 0, s, partials, nblocks, stat_len(C), 0, LF_STATS_HEADER, 1, stats));
}

int rows_backward(const RowsArgs& a, int mode, cudaStream_t s) {
  const int nb = row_blocks(a.B);
  if (rows_reg_supported(a.C))
    return rows_vec_supported(a, mode == LF_MODE_QMF) ? rows_backward_vec(a, mode, nb, s) : rows_backward_reg(a, mode, nb, s);
  const size_t sm = (size_t)8 * 2 * a.C * sizeof(float);
  if (mode == LF_MODE_QMF) {
    if (sm > 48 * 1024) cudaFuncSetAttribute(rows_backward_kernel<LF_MODE_QMF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    LF_LAUNCH("rows_backward_qmf", s, (rows_backward_kernel<LF_MODE_QMF><<<nb, 256, sm, s>>>(a)));
  } else {
    LF_LAUNCH("rows_calibrated", s, (rows_backward_kernel<LF_MODE_JLOGITS><<<nb, 256, 0, s>>>(a)));
  }
  return check_launch("rows_backward_kernel");
}

// ---- tiny scalar kernels ---------------------------------------------------------------------
__global__ void loss_finalize_kernel(const double* __restrict__ stats, int mode, int Bg, float* out) {
  const double inv = 1.0 / (double)Bg;
  // each term is a separate fp32 mean in the reference; summed in fp32 (cremad/joint_model_qmf.py:70)
  float loss = (float)(stats[LF_STAT_CE_JOINT] * inv);
  if (mode == LF_MODE_QMF) {
    const float uni = (float)(stats[LF_STAT_CE_X1] * inv) + (float)(stats[LF_STAT_CE_X2] * inv);
    loss = loss + uni + (float)(stats[LF_STAT_REG_SUM] * inv);
  }
  out[0] = loss;
}

__global__ void ema_update_kernel(float* __restrict__ x, float* __restrict__ off, const double* __restrict__ stats,
                                  int C, int Bg, float beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean1 = (float)(stats[LF_STATS_HEADER + c] / (double)Bg);
  const float mean2 = (float)(stats[LF_STATS_HEADER + C + c] / (double)Bg);
  const float x1 = ema_mix(mean1, x[c], beta);                  // utils/EMA.py:33
  const float x2 = ema_mix(mean2, x[C + c], beta);
  x[c] = x1; x[C + c] = x2;
  const float mu = (x1 + x2) / 2.f;                             // utils/EMA.py:38
  off[c] = mu - x1;
  off[C + c] = mu - x2;
}

__global__ void ogm_coeff_kernel(const double* __restrict__ stats, float alpha, float* __restrict__ coeff) {
  const float s1 = (float)stats[LF_STAT_SCORE_X1], s2 = (float)stats[LF_STAT_SCORE_X2];
  const float r1 = s1 / s2;                 // existing_algos/OGM_GE.py:24
  const float r2 = 1.f / r1;                // :25
  float k1 = 1.f, k2 = 1.f;
  if (r1 > 1.f) k1 = 1.f - tanhf(alpha * fmaxf(r1, 0.f));   // :35-37
  else k2 = 1.f - tanhf(alpha * fmaxf(r2, 0.f));            // :38-40
  coeff[0] = k1; coeff[1] = k2;
}

int loss_finalize(const double* stats, int mode, int Bg, float* out, cudaStream_t s) {
  LF_LAUNCH("loss_finalize", s, (loss_finalize_kernel<<<1, 1, 0, s>>>(stats, mode, Bg, out)));
  return check_launch("loss_finalize_kernel");
}
int ema_update(float* x, float* off, const double* stats, int C, int Bg, float beta, cudaStream_t s) {
  LF_LAUNCH("ema_update", s, (ema_update_kernel<<<div_up(C, 128), 128, 0, s>>>(x, off, stats, C, Bg, beta)));
  return check_launch("ema_update_kernel");
}
int ogm_coeff(const double* stats, float alpha, float* coeff, cudaStream_t s) {
  LF_LAUNCH("ogm_coeff", s, (ogm_coeff_kernel<<<1, 1, 0, s>>>(stats, alpha, coeff)));
  return check_launch("ogm_coeff_kernel");
}

}  // namespace lf
