// Hidden layers of the Food101 per-modality MLPs (food101/joint_model_qmf.py:12-26):
//
//   forward    H = dropout_p(relu(X W^T + b))                      nn.Linear -> nn.ReLU -> nn.Dropout(0.2)
//   backward   dP = dH * [H > 0] / (1 - p)                         (H > 0 exactly where the unit was kept AND active)
//              dX = dP W,   dW = dP^T X,   db = sum_rows dP
//
// for the two modalities' layers of one shape in one launch each.  The GEMMs are the tensor-pipe kernel of lf_tc.cu
// (bf16 / TF32 / 3xTF32 by LfHiddenArgs.precision); bias, ReLU and the Philox dropout mask are applied in the forward
// GEMM's epilogue straight from TMEM, so the pre-activation never exists in HBM and no mask is stored: the backward
// re-derives it from H.  dP is formed by one streaming kernel that also accumulates db per CTA; the split-K partials of
// dW and the db partials are summed in a fixed order (bit-reproducible for a given seed / offset).
#include <cuda_bf16.h>
#include "lf_common.cuh"
#include "lf_gemm.cuh"
#include "lf_tc.cuh"

namespace lf {

int cast_weights_bf16(const float* w0, const float* w1, void* out, size_t n, cudaStream_t s);   // lf_gemm.cu

constexpr int kHidRowsPerCta = 64;

template <class T> __device__ __forceinline__ float hid_ld(const T* p);
template <> __device__ __forceinline__ float hid_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float hid_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <class T> __device__ __forceinline__ void hid_st(T* p, float v);
template <> __device__ __forceinline__ void hid_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void hid_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// dP = dH * [H > 0] * scale; per-CTA column sums of dP (fp32) into dbpart[modality][block][Dout].
// One CTA = kHidRowsPerCta rows x all columns.  Vector path (Dout / VEC threads per row divide the CTA): a thread moves 16
// bytes of a row per access and keeps the sums of its VEC columns over the rows of its row group; the row groups are then
// added in order through shared memory.  Scalar path for other widths: a thread owns whole columns.
template <class T> struct HidVec;
template <> struct HidVec<float> { static constexpr int n = 4; };
template <> struct HidVec<__nv_bfloat16> { static constexpr int n = 8; };

template <class T>
__global__ void __launch_bounds__(256) hidden_dpre_kernel(const T* __restrict__ dh0, const T* __restrict__ dh1, const T* __restrict__ h0,
                                                          const T* __restrict__ h1, T* __restrict__ dp0, T* __restrict__ dp1,
                                                          float* __restrict__ dbpart, int B, int Dout, float scale, int nblocks) {
  constexpr int VEC = HidVec<T>::n;
  __shared__ float colsum[256 * VEC];
  const int m = blockIdx.y, blk = blockIdx.x;
  const T* dh = m ? dh1 : dh0; const T* h = m ? h1 : h0; T* dp = m ? dp1 : dp0;
  const int r0 = blk * kHidRowsPerCta, r1 = min(B, r0 + kHidRowsPerCta);
  float* dst = dbpart + ((size_t)m * nblocks + blk) * Dout;
  const int tpr = Dout / VEC;                                    // threads per row
  if (Dout % VEC == 0 && tpr <= 256 && 256 % tpr == 0) {
    const int q = threadIdx.x % tpr, rg = threadIdx.x / tpr, R = 256 / tpr;
    float s[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) s[e] = 0.f;
#pragma unroll 4
    for (int r = r0 + rg; r < r1; r += R) {
      const uint4 hv = *reinterpret_cast<const uint4*>(h + (size_t)r * Dout + q * VEC);
      const uint4 gv = *reinterpret_cast<const uint4*>(dh + (size_t)r * Dout + q * VEC);
      uint4 ov;
      const T* hp = reinterpret_cast<const T*>(&hv); const T* gp = reinterpret_cast<const T*>(&gv); T* op = reinterpret_cast<T*>(&ov);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        hid_st(op + e, hid_ld(hp + e) > 0.f ? hid_ld(gp + e) * scale : 0.f);
        s[e] += hid_ld(op + e);                                  // what the GEMMs will read (bf16-rounded in bf16 mode)
      }
      *reinterpret_cast<uint4*>(dp + (size_t)r * Dout + q * VEC) = ov;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) colsum[(rg * tpr + q) * VEC + e] = s[e];
    __syncthreads();
    for (int c = threadIdx.x; c < Dout; c += 256) {
      float t = 0.f;
      for (int g = 0; g < R; ++g) t += colsum[g * Dout + c];
      dst[c] = t;
    }
    return;
  }
  for (int c = threadIdx.x; c < Dout; c += 256) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) {
      T out;
      hid_st(&out, hid_ld(h + (size_t)r * Dout + c) > 0.f ? hid_ld(dh + (size_t)r * Dout + c) * scale : 0.f);
      dp[(size_t)r * Dout + c] = out;
      s += hid_ld(&out);
    }
    dst[c] = s;
  }
}

// db = the block partials added per column: one warp per column, lane l takes blocks l, l + 32, ... in order, fixed butterfly
__global__ void __launch_bounds__(256) hidden_db_kernel(const float* __restrict__ dbpart, float* __restrict__ db0, float* __restrict__ db1,
                                                        int Dout, int nblocks) {
  const int m = blockIdx.y, lane = threadIdx.x & 31;
  float* db = m ? db1 : db0;
  for (int c = blockIdx.x * 8 + (threadIdx.x >> 5); c < Dout; c += gridDim.x * 8) {
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += dbpart[((size_t)m * nblocks + b) * Dout + c];
    s = warp_sum(s);
    if (lane == 0) db[c] = s;
  }
}

struct HiddenWs { void* w16; float* dw_part; float* dbpart; size_t total; };
static HiddenWs hidden_carve(void* base, int B, int Din, int Dout) {
  HiddenWs w; size_t off = 0;
  auto take = [&](size_t n) { void* p = base ? (char*)base + off : nullptr; off += align_up(n, 256); return p; };
  w.w16 = take((size_t)2 * Din * Dout * 2);
  w.dw_part = (float*)take((size_t)2 * kMaxSplits * Din * Dout * sizeof(float));
  w.dbpart = (float*)take((size_t)2 * div_up(B, kHidRowsPerCta) * Dout * sizeof(float));
  w.total = off;
  return w;
}

static int hidden_check(const LfHiddenArgs* a, bool backward) {
  if (!a) { set_error("null LfHiddenArgs"); return LF_ERR_BAD_ARG; }
  if (a->batch < 1 || a->dim_in < 8 || a->dim_out < 8 || a->dim_in % 8 || a->dim_out % 8) {
    set_error("lf_hidden: batch >= 1, dim_in / dim_out multiples of 8 (got B=%d %d -> %d)", a->batch, a->dim_in, a->dim_out);
    return LF_ERR_BAD_ARG;
  }
  if (a->precision != LF_PREC_FP32 && a->precision != LF_PREC_TF32 && a->precision != LF_PREC_BF16) { set_error("lf_hidden: bad precision %d", a->precision); return LF_ERR_BAD_ARG; }
  if (a->drop_p < 0.f || a->drop_p >= 1.f) { set_error("lf_hidden: dropout probability %g", (double)a->drop_p); return LF_ERR_BAD_ARG; }
  for (int m = 0; m < 2; ++m) {
    if (!a->x[m] || !a->weight[m] || !a->bias[m] || !a->h[m]) { set_error("lf_hidden: null forward pointer (layer %d)", m); return LF_ERR_BAD_ARG; }
    if (backward && (!a->dh[m] || !a->dpre[m] || !a->dweight[m] || !a->dbias[m])) { set_error("lf_hidden: null backward pointer (layer %d)", m); return LF_ERR_BAD_ARG; }
  }
  if (!a->workspace || a->workspace_bytes < hidden_carve(nullptr, a->batch, a->dim_in, a->dim_out).total) { set_error("lf_hidden: workspace too small"); return LF_ERR_WORKSPACE; }
  return LF_OK;
}

}  // namespace lf

using namespace lf;

extern "C" size_t lf_hidden_workspace_bytes(int32_t batch, int32_t dim_in, int32_t dim_out) {
  if (batch < 1 || dim_in < 1 || dim_out < 1) return 0;
  return hidden_carve(nullptr, batch, dim_in, dim_out).total;
}

extern "C" int lf_hidden_forward(const LfHiddenArgs* a, void* stream) {
  int rc = hidden_check(a, false);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  HiddenWs w = hidden_carve(a->workspace, a->batch, a->dim_in, a->dim_out);
  const bool bf16 = a->precision == LF_PREC_BF16;
  const size_t n = (size_t)a->dim_in * a->dim_out;
  if (bf16) { rc = cast_weights_bf16(a->weight[0], a->weight[1], w.w16, n, s); if (rc) return rc; }   // autocast's per-call cast of nn.Linear.weight
  TcGemmDesc d;
  d.nbatch = 2;
  d.elem = bf16 ? 2 : 4; d.out_elem = bf16 ? 2 : 4; d.x3 = a->precision == LF_PREC_FP32;
  for (int m = 0; m < 2; ++m) {
    d.A[m] = a->x[m];
    d.B[m] = bf16 ? (const void*)((const char*)w.w16 + m * n * 2) : (const void*)a->weight[m];
    d.bias[m] = a->bias[m]; d.out[m] = a->h[m];
  }
  d.M = a->batch; d.N = a->dim_out; d.K = a->dim_in;
  d.lda = a->dim_in; d.ldb = a->dim_in; d.ld_out = a->dim_out;
  d.a_mn_major = 0; d.b_mn_major = 0;
  d.block_n = a->dim_out >= 256 ? 256 : div_up(a->dim_out, 64) * 64;
  d.splits = 1; d.split_stride = 0; d.balance_m = 0;
  d.act_relu = 1; d.drop_p = a->training ? a->drop_p : 0.f; d.seed = a->seed; d.rng_offset = a->offset;
  d.name = "hidden_forward";
  return tc_gemm(d, s);
}

extern "C" int lf_hidden_backward(const LfHiddenArgs* a, void* stream) {
  int rc = hidden_check(a, true);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  HiddenWs w = hidden_carve(a->workspace, a->batch, a->dim_in, a->dim_out);
  const bool bf16 = a->precision == LF_PREC_BF16;
  const int B = a->batch, Din = a->dim_in, Dout = a->dim_out;
  const size_t n = (size_t)Din * Dout;
  const float scale = (a->training && a->drop_p > 0.f) ? 1.f / (1.f - a->drop_p) : 1.f;
  const int nblocks = div_up(B, kHidRowsPerCta);
  // ---- dP and the db partials
  if (bf16) {
    LF_LAUNCH("hidden_dpre", s, (hidden_dpre_kernel<__nv_bfloat16><<<dim3(nblocks, 2), 256, 0, s>>>(
        (const __nv_bfloat16*)a->dh[0], (const __nv_bfloat16*)a->dh[1], (const __nv_bfloat16*)a->h[0], (const __nv_bfloat16*)a->h[1],
        (__nv_bfloat16*)a->dpre[0], (__nv_bfloat16*)a->dpre[1], w.dbpart, B, Dout, scale, nblocks)));
  } else {
    LF_LAUNCH("hidden_dpre", s, (hidden_dpre_kernel<float><<<dim3(nblocks, 2), 256, 0, s>>>(
        (const float*)a->dh[0], (const float*)a->dh[1], (const float*)a->h[0], (const float*)a->h[1], (float*)a->dpre[0], (float*)a->dpre[1],
        w.dbpart, B, Dout, scale, nblocks)));
  }
  rc = check_launch("hidden_dpre_kernel");
  if (rc) return rc;
  LF_LAUNCH("hidden_db", s, (hidden_db_kernel<<<dim3(min(296, div_up(Dout, 8)), 2), 256, 0, s>>>(w.dbpart, a->dbias[0], a->dbias[1], Dout, nblocks)));
  rc = check_launch("hidden_db_kernel");
  if (rc) return rc;
  // ---- dX = dP W   (A = dP K-major, B = W (Dout x Din) row-major = MN-major)
  if (a->dx[0] && a->dx[1]) {
    if (bf16) { rc = cast_weights_bf16(a->weight[0], a->weight[1], w.w16, n, s); if (rc) return rc; }
    TcGemmDesc d;
    d.nbatch = 2; d.elem = bf16 ? 2 : 4; d.out_elem = bf16 ? 2 : 4; d.x3 = a->precision == LF_PREC_FP32;
    for (int m = 0; m < 2; ++m) {
      d.A[m] = a->dpre[m];
      d.B[m] = bf16 ? (const void*)((const char*)w.w16 + m * n * 2) : (const void*)a->weight[m];
      d.bias[m] = nullptr; d.out[m] = a->dx[m];
    }
    d.M = B; d.N = Din; d.K = Dout; d.lda = Dout; d.ldb = Din; d.ld_out = Din;
    d.a_mn_major = 0; d.b_mn_major = 1; d.block_n = Din >= 256 ? 256 : div_up(Din, 64) * 64;
    d.splits = 1; d.split_stride = 0; d.balance_m = 0; d.name = "hidden_dx";
    rc = tc_gemm(d, s);
    if (rc) return rc;
  } else if (a->dx[0] || a->dx[1]) {
    set_error("lf_hidden_backward: dx must be given for both layers or for neither");
    return LF_ERR_BAD_ARG;
  }
  // ---- dW = dP^T X   (both MN-major, split-K over the batch; partials summed in split order)
  {
    TcGemmDesc d;
    d.nbatch = 2; d.elem = bf16 ? 2 : 4; d.out_elem = 4; d.x3 = a->precision == LF_PREC_FP32;
    for (int m = 0; m < 2; ++m) { d.A[m] = a->dpre[m]; d.B[m] = a->x[m]; d.bias[m] = nullptr; d.out[m] = w.dw_part + (size_t)m * kMaxSplits * n; }
    d.M = Dout; d.N = Din; d.K = B; d.lda = Dout; d.ldb = Din; d.ld_out = Din;
    d.a_mn_major = 1; d.b_mn_major = 1; d.block_n = Din >= 256 ? 256 : div_up(Din, 64) * 64;
    const int tiles = div_up(Dout, 128) * div_up(Din, d.block_n) * 2;
    int splits = 148 / tiles;
    const int by_rows = div_up(B, 128);
    if (splits > by_rows) splits = by_rows;
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
    d.splits = splits; d.split_stride = (long long)n; d.balance_m = 0; d.name = "hidden_dw";
    rc = tc_gemm(d, s);
    if (rc) return rc;
    rc = reduce_splits2(w.dw_part, a->dweight[0], a->dweight[1], splits, kMaxSplits, n, s);
  }
  return rc;
}
