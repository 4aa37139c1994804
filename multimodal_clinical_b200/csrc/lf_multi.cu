// Mean fusion of M = 2..4 narrow heads with per-modality feature widths:
//
//   z_m = f_m W_m^T + b_m   (f_m: (B, D_m), W_m: (C, D_m))          mustard/joint_model.py:72-74, avmnist/joint_model.py:128-131
//   avg = (z_1 + ... + z_M) / M ;  L = CE(avg, y)                    mustard/joint_model.py:77-80, avmnist/joint_model.py:134-136
//   backward of the heads: dz = (softmax(avg) - onehot(y)) / (M B), dW_m = dz^T f_m, db_m = sum dz, df_m = dz W_m
//
// (mustard: three LstmClassifier.fc3 heads of width 100, two classes; avmnist: widths 48 and 192, ten classes.)  The loss
// has no grid-wide dependency, so forward AND backward are ONE pass over the features: a CTA walks tiles of S samples;
// per tile the features of all modalities are staged side by side in shared memory ([S][sum D_m]), the heads are resident
// in shared memory for the whole kernel, and four phases run on the tile:
//   A  coalesced loads of the tile                          B  one thread per (sample, modality): all C dot products together
//   C  one thread per sample: mean, softmax, CE, argmaxes, dz
//   D  one thread per feature column: df (written once, coalesced) and the dW column, accumulated in a shared-memory
//      column the thread owns (no atomics); db by one thread per class
// Per-CTA partials (dW, db, CE sum, counts) are summed in CTA order by a second small kernel, so the step is
// bit-reproducible.  Exact fp32 (parity class 1e-5).  HBM-bound: 4 (sum D_m) B read + the same written per sample.
#include "lf_common.cuh"

namespace lf {

constexpr int kMultiThreads = 256;
constexpr int kMultiMaxGrid = 592;

struct MultiParams {
  int M, B, C, Dtot, S, ldw;
  int D[LF_MAX_MODALITIES], off[LF_MAX_MODALITIES];
  const float* feat[LF_MAX_MODALITIES];
  const float* W[LF_MAX_MODALITIES];
  const float* bias[LF_MAX_MODALITIES];
  const int64_t* label;
  float* logits[LF_MAX_MODALITIES];
  float* avg;
  float* dfeat[LF_MAX_MODALITIES];
  float* dw_part;    // [grid][C][Dtot]
  float* db_part;    // [grid][C]
  double* st_part;   // [grid][2 + M]: CE sum, joint hits, per-modality hits
  int need_dfeat;
  float inv_mb;      // 1 / (M B)
};

__device__ __forceinline__ int modality_of(const MultiParams& p, int j) {
  int m = 0;
#pragma unroll
  for (int k = 1; k < LF_MAX_MODALITIES; ++k) if (k < p.M && j >= p.off[k]) m = k;
  return m;
}

template <int CPAD>
__global__ void __launch_bounds__(kMultiThreads) multi_heads_kernel(MultiParams p) {
  extern __shared__ __align__(16) float sm[];
  const int C = p.C, M = p.M, S = p.S, Dtot = p.Dtot, ldw = p.ldw;
  float* Wt = sm;                               // [Dtot][CPAD]  heads side by side, TRANSPOSED: the C weights of a feature are one
                                                //               16-byte-aligned row, read with broadcast LDS.128
  float* dzs = Wt + (size_t)Dtot * CPAD;        // [S][CPAD]     avg, then dz (same reason for the pitch)
  float* dWs = dzs + (size_t)S * CPAD;          // [C][Dtot]     this CTA's dW, column j owned by one thread
  float* fs = dWs + (size_t)C * Dtot;           // [S][ldw]      feature tile; ldw odd: samples on different banks
  float* zs = fs + (size_t)S * ldw;             // [S][M*C]
  float* bs = zs + (size_t)S * M * C;           // [M*C]
  double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(bs + M * C) + 7) & ~(uintptr_t)7);   // [S][2 + M] end-of-kernel reduction
  const int tid = threadIdx.x;

  for (int i = tid; i < Dtot * CPAD; i += kMultiThreads) Wt[i] = 0.f;
  __syncthreads();
  for (int m = 0; m < M; ++m) {
    const int Dm = p.D[m], o = p.off[m];
    for (int i = tid; i < C * Dm; i += kMultiThreads) { const int c = i / Dm, d = i - c * Dm; Wt[(o + d) * CPAD + c] = p.W[m][i]; }
    for (int c = tid; c < C; c += kMultiThreads) bs[m * C + c] = p.bias[m][c];
  }
  for (int i = tid; i < C * Dtot; i += kMultiThreads) dWs[i] = 0.f;
  for (int i = tid; i < S * CPAD; i += kMultiThreads) dzs[i] = 0.f;      // the pad columns stay zero
  double ce = 0.0;                               // thread s < S: sums over the samples it owned
  int hits_joint = 0, hits_m[LF_MAX_MODALITIES] = {0, 0, 0, 0};
  float db_acc = 0.f;                            // thread c < C
  __syncthreads();

  const int tiles = (p.B + S - 1) / S;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int row0 = tile * S, rows = min(S, p.B - row0);
    // ---- A: the tile's features, every modality's rows contiguous in global memory
    for (int m = 0; m < M; ++m) {
      const int Dm = p.D[m], o = p.off[m];
      const float* src = p.feat[m] + (size_t)row0 * Dm;
      if ((Dm & 3) == 0) {
        const int q4 = Dm >> 2;
        for (int i = tid; i < rows * q4; i += kMultiThreads) {
          const int s = i / q4, q = i - s * q4;
          const float4 v = ldg_stream(reinterpret_cast<const float4*>(src) + i);
          float* dst = fs + s * ldw + o + 4 * q;
          dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
      } else {
        for (int i = tid; i < rows * Dm; i += kMultiThreads) { const int s = i / Dm, d = i - s * Dm; fs[s * ldw + o + d] = __ldg(src + i); }
      }
    }
    __syncthreads();
    // ---- B: logits.  One (modality, sample) per thread, modality-major so that a warp shares the weight rows it reads;
    // all C dot products of the pair advance together: per feature one LDS of f and CPAD / 4 broadcast LDS.128 of weights
    for (int i = tid; i < rows * M; i += kMultiThreads) {
      const int m = i / rows, s = i - m * rows;
      const float* fr = fs + s * ldw + p.off[m];
      const float4* wr = reinterpret_cast<const float4*>(Wt + (size_t)p.off[m] * CPAD);
      float acc[CPAD];
#pragma unroll
      for (int c = 0; c < CPAD; ++c) acc[c] = 0.f;
      for (int d = 0; d < p.D[m]; ++d) {
        const float fv = fr[d];
#pragma unroll
        for (int q = 0; q < CPAD / 4; ++q) {
          const float4 w4 = wr[d * (CPAD / 4) + q];
          acc[4 * q + 0] = fmaf(fv, w4.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(fv, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(fv, w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(fv, w4.w, acc[4 * q + 3]);
        }
      }
      float* zr = zs + (s * M + m) * C;
      float* out = p.logits[m] + (size_t)(row0 + s) * C;
#pragma unroll
      for (int c = 0; c < CPAD; ++c)
        if (c < C) { const float z = acc[c] + bs[m * C + c]; zr[c] = z; out[c] = z; }
    }
    __syncthreads();
    // ---- C: row math, one sample per thread
    for (int s = tid; s < rows; s += kMultiThreads) {
      const float* zr = zs + s * M * C;
      float* ar = dzs + s * CPAD;
      const long long y = (long long)p.label[row0 + s];
      float mx = -INFINITY; int arg = 0;
      for (int c = 0; c < C; ++c) {
        float a = zr[c];
        for (int m = 1; m < M; ++m) a += zr[m * C + c];
        a = __fdiv_rn(a, (float)M);
        ar[c] = a;
        p.avg[(size_t)(row0 + s) * C + c] = a;
        if (!(mx != mx) && (a > mx || a != a)) { mx = a; arg = c; }          // first maximum; a NaN wins and stays (torch.argmax)
      }
      hits_joint += (arg == y);
      for (int m = 0; m < M; ++m) {
        float bm = -INFINITY; int am = 0;
        for (int c = 0; c < C; ++c) { const float v = zr[m * C + c]; if (!(bm != bm) && (v > bm || v != v)) { bm = v; am = c; } }
        hits_m[m] += (am == y);
      }
      float den = 0.f;
      for (int c = 0; c < C; ++c) den += expf(ar[c] - mx);
      const float lse = mx + logf(den);
      const float ay = (y >= 0 && y < C) ? ar[y] : 0.f;
      ce += (double)(lse - ay);
      const float inv = 1.f / den;
      for (int c = 0; c < C; ++c) ar[c] = (expf(ar[c] - mx) * inv - (c == y ? 1.f : 0.f)) * p.inv_mb;
    }
    __syncthreads();
    // ---- D: df and dW, one feature column per thread and trip; db by the first C threads
    for (int j = tid; j < Dtot; j += kMultiThreads) {
      const int m = modality_of(p, j);
      const int Dm = p.D[m], d = j - p.off[m];
      float w[CPAD], acc[CPAD];
#pragma unroll
      for (int q = 0; q < CPAD / 4; ++q) {
        const float4 w4 = reinterpret_cast<const float4*>(Wt + (size_t)j * CPAD)[q];
        w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
      }
#pragma unroll
      for (int c = 0; c < CPAD; ++c) acc[c] = 0.f;
      float* out = p.dfeat[m] + (size_t)row0 * Dm + d;
      for (int s = 0; s < rows; ++s) {
        const float fv = fs[s * ldw + j];
        const float4* dz4 = reinterpret_cast<const float4*>(dzs + s * CPAD);
        float g = 0.f;
#pragma unroll
        for (int q = 0; q < CPAD / 4; ++q) {                    // pad columns of dz and w are zero
          const float4 v = dz4[q];
          g = fmaf(v.x, w[4 * q], g); g = fmaf(v.y, w[4 * q + 1], g); g = fmaf(v.z, w[4 * q + 2], g); g = fmaf(v.w, w[4 * q + 3], g);
          acc[4 * q] = fmaf(v.x, fv, acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, fv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v.z, fv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, fv, acc[4 * q + 3]);
        }
        if (p.need_dfeat) out[(size_t)s * Dm] = g;
      }
#pragma unroll
      for (int c = 0; c < CPAD; ++c) if (c < C) dWs[c * Dtot + j] += acc[c];
    }
    if (tid < C) for (int s = 0; s < rows; ++s) db_acc += dzs[s * CPAD + tid];
    __syncthreads();
  }
  // ---- this CTA's partials
  float* dwp = p.dw_part + (size_t)blockIdx.x * C * Dtot;
  for (int i = tid; i < C * Dtot; i += kMultiThreads) dwp[i] = dWs[i];
  if (tid < C) p.db_part[(size_t)blockIdx.x * C + tid] = db_acc;
  const int owners = min(S, kMultiThreads);
  if (tid < owners) {
    red[tid * (2 + M) + 0] = ce; red[tid * (2 + M) + 1] = (double)hits_joint;
    for (int m = 0; m < M; ++m) red[tid * (2 + M) + 2 + m] = (double)hits_m[m];
  }
  __syncthreads();
  if (tid < 2 + M) {
    double s = 0.0;
    for (int t = 0; t < owners; ++t) s += red[t * (2 + M) + tid];
    p.st_part[(size_t)blockIdx.x * (2 + M) + tid] = s;
  }
}

// Sums of the per-CTA partials: dW (scattered back to the per-modality tensors), db (the same vector for every head: dz
// does not depend on the modality), statistics and the batch-mean loss.  One WARP per output: lane l adds partials l,
// l + 32, ... in order and a fixed butterfly adds the lanes -- the same order on every launch, and ~nparts / 32 dependent
// loads per lane instead of nparts (the first version's one thread per output took 50 us for 592 partials).
__global__ void __launch_bounds__(256) multi_finalize_kernel(MultiParams p, int nparts, float* dweight0, float* dweight1, float* dweight2,
                                                            float* dweight3, float* dbias0, float* dbias1, float* dbias2, float* dbias3,
                                                            double* stats, float* loss_out) {
  float* dweight[LF_MAX_MODALITIES] = {dweight0, dweight1, dweight2, dweight3};
  float* dbias[LF_MAX_MODALITIES] = {dbias0, dbias1, dbias2, dbias3};
  const int n = p.C * p.Dtot, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  const int extra = p.C + 2 + p.M;                     // db entries and statistics follow the dW elements
  for (int i = warp; i < n + extra; i += nwarps) {
    if (i < n) {
      float s = 0.f;
      for (int g = lane; g < nparts; g += 32) s += p.dw_part[(size_t)g * n + i];
      s = warp_sum(s);
      const int c = i / p.Dtot, j = i - c * p.Dtot;
      const int m = modality_of(p, j);
      if (lane == 0) dweight[m][(size_t)c * p.D[m] + (j - p.off[m])] = s;
    } else if (i < n + p.C) {
      const int c = i - n;
      float s = 0.f;
      for (int g = lane; g < nparts; g += 32) s += p.db_part[(size_t)g * p.C + c];
      s = warp_sum(s);
      if (lane == 0) for (int m = 0; m < p.M; ++m) dbias[m][c] = s;
    } else {
      const int k = i - n - p.C;
      double s = 0.0;
      for (int g = lane; g < nparts; g += 32) s += p.st_part[(size_t)g * (2 + p.M) + k];
      s = warp_sum(s);
      if (lane == 0) { stats[k] = s; if (k == 0) loss_out[0] = (float)(s / (double)p.B); }
    }
  }
}

static int multi_cpad(int C) { return C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32; }
static size_t multi_smem_bytes(int M, int C, int Dtot, int S) {
  const int ldw = Dtot | 1, cpad = multi_cpad(C);
  size_t f = (size_t)Dtot * cpad + (size_t)S * cpad + (size_t)C * Dtot + (size_t)S * ldw + (size_t)S * M * C + (size_t)M * C;
  return f * sizeof(float) + 8 + (size_t)S * (2 + M) * sizeof(double);
}

static int multi_plan(const LfMultiHeadsArgs* a, int* S_out, size_t* smem_out, int* grid_out, int* Dtot_out) {
  int Dtot = 0;
  for (int m = 0; m < a->modalities; ++m) Dtot += a->dim[m];
  int S = 64;
  while (S > 8 && multi_smem_bytes(a->modalities, a->classes, Dtot, S) > 100 * 1024) S >>= 1;
  const size_t smem = multi_smem_bytes(a->modalities, a->classes, Dtot, S);
  if (smem > 220 * 1024) return LF_ERR_UNSUPPORTED;
  int per_sm = (int)((size_t)(220 * 1024) / smem);
  if (per_sm > 4) per_sm = 4;
  int grid = div_up(a->batch, S);
  if (grid > 148 * per_sm) grid = 148 * per_sm;
  *S_out = S; *smem_out = smem; *grid_out = grid; *Dtot_out = Dtot;
  return LF_OK;
}

static int multi_check(const LfMultiHeadsArgs* a) {
  if (!a) { set_error("null LfMultiHeadsArgs"); return LF_ERR_BAD_ARG; }
  if (a->modalities < 2 || a->modalities > LF_MAX_MODALITIES || a->batch < 1 || a->classes < 1 || a->classes > 32) {
    set_error("lf_multi_heads_step: modalities 2..%d, classes 1..32, batch >= 1 (got M=%d C=%d B=%d)", LF_MAX_MODALITIES, a->modalities,
              a->classes, a->batch);
    return a && a->classes > 32 ? LF_ERR_UNSUPPORTED : LF_ERR_BAD_ARG;
  }
  for (int m = 0; m < a->modalities; ++m) {
    if (a->dim[m] < 1) { set_error("lf_multi_heads_step: dim[%d] = %d", m, a->dim[m]); return LF_ERR_BAD_ARG; }
    if (!a->feat[m] || !a->weight[m] || !a->bias[m] || !a->logits[m] || !a->dweight[m] || !a->dbias[m] || (a->need_dfeat && !a->dfeat[m])) {
      set_error("lf_multi_heads_step: null per-modality pointer (modality %d)", m);
      return LF_ERR_BAD_ARG;
    }
    if ((a->dim[m] & 3) == 0 && ((uintptr_t)a->feat[m] & 15)) { set_error("lf_multi_heads_step: feat[%d] must be 16-byte aligned", m); return LF_ERR_BAD_ARG; }
  }
  if (!a->label || !a->avg_logits || !a->loss_out || !a->stats || !a->workspace) { set_error("lf_multi_heads_step: null pointer argument"); return LF_ERR_BAD_ARG; }
  return LF_OK;
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_multi_heads_workspace_bytes(int32_t modalities, int32_t classes, int32_t dim_total) {
  if (modalities < 2 || modalities > LF_MAX_MODALITIES || classes < 1 || dim_total < 1) return 0;
  return align_up((size_t)kMultiMaxGrid * classes * dim_total * sizeof(float), 256) + align_up((size_t)kMultiMaxGrid * classes * sizeof(float), 256) +
         align_up((size_t)kMultiMaxGrid * (2 + modalities) * sizeof(double), 256);
}

extern "C" int lf_multi_heads_step(const LfMultiHeadsArgs* a, void* stream) {
  int rc = multi_check(a);
  if (rc) return rc;
  int S, grid, Dtot; size_t smem;
  if (multi_plan(a, &S, &smem, &grid, &Dtot)) {
    set_error("lf_multi_heads_step: heads of %d classes x %d features do not fit shared memory", a->classes, Dtot);
    return LF_ERR_UNSUPPORTED;
  }
  if (a->workspace_bytes < lf_multi_heads_workspace_bytes(a->modalities, a->classes, Dtot)) {
    set_error("lf_multi_heads_step: workspace too small");
    return LF_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  MultiParams p = {};
  p.M = a->modalities; p.B = a->batch; p.C = a->classes; p.Dtot = Dtot; p.S = S; p.ldw = Dtot | 1;
  int off = 0;
  for (int m = 0; m < LF_MAX_MODALITIES; ++m) {
    const int k = m < a->modalities ? m : 0;
    p.D[m] = a->dim[k]; p.off[m] = m < a->modalities ? off : 0x7fffffff;
    if (m < a->modalities) off += a->dim[m];
    p.feat[m] = a->feat[k]; p.W[m] = a->weight[k]; p.bias[m] = a->bias[k]; p.logits[m] = a->logits[k]; p.dfeat[m] = a->dfeat[k];
  }
  p.label = a->label; p.avg = a->avg_logits; p.need_dfeat = a->need_dfeat;
  p.inv_mb = 1.0f / ((float)a->modalities * (float)a->batch);
  char* ws = (char*)a->workspace;
  p.dw_part = (float*)ws; ws += align_up((size_t)kMultiMaxGrid * p.C * Dtot * sizeof(float), 256);
  p.db_part = (float*)ws; ws += align_up((size_t)kMultiMaxGrid * p.C * sizeof(float), 256);
  p.st_part = (double*)ws;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    LF_LAUNCH("multi_heads_step", s, (kernel<<<grid, kMultiThreads, smem, s>>>(p)));
  };
  if (p.C <= 4) launch(multi_heads_kernel<4>);
  else if (p.C <= 8) launch(multi_heads_kernel<8>);
  else if (p.C <= 16) launch(multi_heads_kernel<16>);
  else launch(multi_heads_kernel<32>);
  rc = check_launch("multi_heads_kernel");
  if (rc) return rc;
  const int fgrid = min(592, div_up((long long)p.C * Dtot + p.C + 2 + p.M, 8));
  float* dw[LF_MAX_MODALITIES]; float* db[LF_MAX_MODALITIES];
  for (int m = 0; m < LF_MAX_MODALITIES; ++m) { const int k = m < a->modalities ? m : 0; dw[m] = a->dweight[k]; db[m] = a->dbias[k]; }
  LF_LAUNCH("multi_heads_finalize", s, (multi_finalize_kernel<<<fgrid, 256, 0, s>>>(p, grid, dw[0], dw[1], dw[2], dw[3], db[0], db[1], db[2], db[3],
                                                                                     a->stats, a->loss_out)));
  return check_launch("multi_finalize_kernel");
}
