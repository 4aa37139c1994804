#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/lf_fusion.h"

namespace lf {

struct RowsArgs {
  const float* z[2];     // (B,C) logits
  float* avg;            // (B,C)
  float* zdf;            // (B,C) QMF
  float* conf;           // (2,B) QMF
  float* rowstat;        // (B,4): lse1, lse2, lse_joint, - (QMF forward -> backward)
  float* dz[2];          // (B,C)
  const int64_t* label;  // (B)
  const float* qmf_g;    // (2,B)
  const float* ema_off;  // (2,C)
  float* partials;       // [blocks][stat_len]
  float* dbpart;         // [blocks][2][C] column sums of dz (written by whichever kernel produces dz)
  float* calpart;        // [blocks][2] calibrated-accuracy counts
  double* stats;
  int B, B_global, C;
  int ld_z;              // row pitch of the input logits z (>= C; padded to 16 B when the tensor pipe TMA-stores them)
  int ldz;               // row pitch of dz
  int ld_f;              // row pitch of avg / zdf (>= C)
  float w_joint, w_uni;  // 1 or 0: weight of CE(z_df) / of the unimodal CE terms in dL/dz (QMF loss ablations)
  int dz_bf16;           // 1: dz[] point at bf16 buffers (LF_PREC_BF16): dL/dz is stored rounded to bf16
  int nb_total;          // partial rows the finalize kernels will sum; CTAs zero the rows beyond the grid
};

// dL/dz store: fp32, or bf16 when the tensor pipe runs kind::f16 (the GEMMs then consume it directly)
__device__ __forceinline__ void store_dz(const RowsArgs& a, int m, size_t pos, float d) {
  if (a.dz_bf16) reinterpret_cast<__nv_bfloat16*>(a.dz[m])[pos] = __float2bfloat16_rn(d);
  else a.dz[m][pos] = d;
}

__host__ __device__ inline int stat_len_dev(int C) { return LF_STATS_HEADER + 2 * C; }

int rows_forward(const RowsArgs& a, int mode, cudaStream_t s);
int rows_backward(const RowsArgs& a, int mode, cudaStream_t s);
int row_blocks(int B);
// dbias[m][c] = sum over blocks of dbpart; stats[CNT_X*_CAL] = sum over blocks of calpart
int finalize_db_cal(const float* dbpart, int nb_db, int C, const float* calpart, int nb_cal, float* db0, float* db1,
                    double* stats, cudaStream_t s);
// dW (sum of split-K partials) + db + calibrated counts in one launch (splits <= 32, 16-byte aligned buffers)
bool finalize_grads_supported(const float* part, const float* dw0, const float* dw1, int splits, size_t n);
int finalize_grads(const float* part, float* dw0, float* dw1, int splits, int max_splits, size_t n, const float* dbpart,
                   int nb_db, int C, const float* calpart, int nb_cal, float* db0, float* db1, double* stats, cudaStream_t s);
int loss_finalize(const double* stats, int mode, int Bg, float* out, cudaStream_t s);
int ema_update(float* x, float* off, const double* stats, int C, int Bg, float beta, cudaStream_t s);
int ogm_coeff(const double* stats, float alpha, float* coeff, cudaStream_t s);

}  // namespace lf
