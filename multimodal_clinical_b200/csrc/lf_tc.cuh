#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lf_fusion.h"

namespace lf {

// Optional tail of the split-K dW GEMM (replaces the separate finalize_grads launch): after a grid-wide barrier every
// CTA sums a slice of the split-K partials in split order (bit-reproducible), CTA 0 sums the db / calibrated-count
// partials of the kernel that produced dz, and -- when hyper != null -- the SGD update of the heads is applied in place.
struct TcTail {
  int on;
  unsigned* sync;            // [0] arrivals, [1] departures; zero before the launch, zero after it
  float* dw[2];              // out (C*D)
  long long n;               // C*D
  const float* dbpart;       // [nb_db][2][C]
  const float* calpart;      // [nb_cal][2]
  int nb_db, nb_cal, C;
  float* db[2];              // out (C)
  double* stats;             // calibrated counts land in stats[LF_STAT_CNT_X1_CAL ..]
  const float* hyper;        // device [lr, momentum, weight_decay]; null = no optimizer step
  float* param_w[2];
  float* param_b[2];
  float* mom_w[2];
  float* mom_b[2];
  void* w16[2];              // optional bf16 copies of the updated weights
  unsigned long long* trace; // LF_DW_TRACE=1: [grid][8] %globaltimer stamps
  // Sharded steps: the gradient all-reduce runs in here.  Every CTA stores the chunk it has just reduced into slot
  // [parity][rank] of every rank's receive area (layout [dW1 | dW2 | db1 | db2 | cal x2 | reg], n_padded floats) and then
  // polls the same chunk of every rank in its LOCAL receive area against the sentinel the slots are armed with
  // (lf_peer.cuh: no fence, flag or barrier; the ranks run the same grid over the same chunks), sums the ranks' chunks
  // in rank order and re-arms the slot.
  int peer_on;
  LfPeerComm comm;
  int n_padded;
  const float* reg_local;    // QMF: this rank's ranking-loss partial (lf_step_mid), or null
  float* loss_out;           // QMF: loss without the ranking term; the all-reduced term / batch_global is added
  int batch_global;
};

// Device-side parameters of the persistent tc_gemm_kernel.
struct TcGemmParams {
  int M, N, K;             // true problem sizes (ragged edges are zero-filled by TMA / clipped on store)
  int block_n;             // N tile: multiple of 16 (32 when B is MN-major or the TMA-store epilogue is used), <= 256
  int nbatch;              // 1 or 2 (the two modalities)
  int splits;              // split-K factor
  int k_per_split;         // multiple of 32
  int stages;              // smem ring depth
  int acc_cols;            // TMEM columns of one accumulator buffer (power of two >= block_n)
  int tmem_cols;           // 2 * acc_cols (double-buffered)
  int a_mn_major, b_mn_major;
  int tma_store;           // epilogue writes through cp.async.bulk.tensor (needs a 16-byte output pitch)
  int epi_halves;          // 1 or 2 groups of four epilogue warps taking alternate 128-byte column chunks
  int elem;                // operand element size: 4 = fp32 consumed as TF32, 2 = bf16
  int chunks, k_per_chunk; // x3 split-K: each split's K range is accumulated in `chunks` pieces (one work item each, same CTA,
                           // in order); piece 0 stores its tile, the others add theirs with a TMA reduce -- the tensor core's
                           // accumulator truncates on every MMA, so a long K chain in TMEM loses ~5e-9 * K relative
  int act_relu;            // epilogue (TMA-store path): out = dropout(relu(acc + bias)) -- the hidden layers of the Food101 MLPs
  float drop_p, drop_scale;           // dropout probability (0 = off) and 1 / (1 - p)
  unsigned long long seed, rng_offset; // Philox4x32-10 key / counter high words; element (row, col) draws 16-bit half
                                       // (e & 7) of counter group e >> 3, e = row * N + col: the mask does not depend on the tiling
  int x3;                  // fp32 operands split into tf32 hi + lo in shared memory, three MMAs per k-step (exact-fp32 tier)
  int out_elem;            // output element size (TMA-store epilogue): 4 = fp32, 2 = bf16
  int kb_elems;            // K elements per stage (128 B per operand row): 32 / 64
  int umma_k;              // K elements per tcgen05.mma: 8 / 16
  int mn_box;              // MN elements per MN-major TMA box (128 B): 32 / 64
  int mn_box_bytes;        // bytes of one MN-major box: 128 B x kb_elems rows
  unsigned mn_step, mn_lbo, mn_sbo, mn_lt;   // MN-major UMMA descriptor: bytes per instruction, LBO, SBO, layout type
  int l2_last_use;         // bit 0 / 1: operand A / B is read for the last time in the step -> loads tagged evict_first
  int tile_m;              // rows per M tile (<= 128): A box rows; smaller tiles balance the item count over the SMs
  int m_tiles, n_tiles, total_items;
  void* out[2];
  const float* bias[2];
  long long ld_out;        // output row pitch (elements)
  long long split_stride;  // elements between split partials
  TcTail tail;
};

// Host-side description of one (batched x2) GEMM:  out[b] (M x N) = A[b] * B[b] (+ bias[b]).
//   a_mn_major = 0: A is (M x K) row-major with pitch lda (K contiguous)
//   a_mn_major = 1: A is stored (K x M) row-major with pitch lda (M contiguous), i.e. A = stored^T
//   b_mn_major = 0: B is stored (N x K) row-major with pitch ldb (K contiguous), i.e. out = A * stored^T
//   b_mn_major = 1: B is (K x N) row-major with pitch ldb (N contiguous)
struct TcGemmDesc {
  int nbatch;
  const void* A[2];        // fp32 or bf16 (elem)
  const void* B[2];
  const float* bias[2];
  void* out[2];            // fp32, or bf16 when out_elem == 2
  int M, N, K;
  long long lda, ldb, ld_out;
  int a_mn_major, b_mn_major;
  int block_n;
  int splits;
  long long split_stride;
  int elem = 4;            // 4 = fp32 operands (TF32 MMA), 2 = bf16 operands
  int act_relu = 0;        // TMA-store epilogue: relu after the bias
  float drop_p = 0.f;      // then inverted dropout with this probability (Philox4x32-10 keyed by seed / rng_offset)
  unsigned long long seed = 0, rng_offset = 0;
  int x3 = 0;              // elem == 4 only: 3xTF32 (hi*hi + hi*lo + lo*hi), fp32-grade products on the tensor pipe
  int out_elem = 4;        // 4 = fp32 output, 2 = bf16 output (TMA-store epilogue only)
  int max_epi_halves = 2;  // 1: never add the second group of epilogue warps (the CTA then leaves room for a concurrent kernel)
  int l2_last_use = 0;     // bit 0: A, bit 1: B are dead after this GEMM (L2 evict_first hint on their loads)
  int balance_m = 0;           // 1: pick tile_m so that the number of work items is a multiple of the SM count (K-major A, plain-store epilogue only)
  TcTail tail = {};            // tail.on: fused reduction of the split-K partials (needs all CTAs co-resident: grid <= 148)
  const char* name;
};

int tc_gemm(const TcGemmDesc& d, cudaStream_t s);

}  // namespace lf
