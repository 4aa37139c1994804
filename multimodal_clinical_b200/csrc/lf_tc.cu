// Tensor-pipe path for the wide heads (Food101 101-way, VGGSound 309-way): one persistent,
// warp-specialised tcgen05 + TMA GEMM kernel (kind::tf32, fp32 accumulators in TMEM) used for the head
// products that are not covered by the fused forward kernel (lf_tc_fwd.cu):
//
//   logits   Z  = F  W^T (+b)   M=B  N=C  K=D    A = F  K-major   B = W  K-major        (C > 256 only)
//   dfeat    dF = dZ W          M=B  N=D  K=C    A = dZ K-major   B = W  MN-major (no transpose pass)
//   dweight  dW = dZ^T F        M=C  N=D  K=B    A = dZ MN-major  B = F  MN-major, split-K over CTAs
//
// Operands stream HBM -> shared memory with TMA (cp.async.bulk.tensor) straight from the row-major fp32
// tensors — K-major operands as one [32 k x rows] SWIZZLE_128B box per stage, MN-major operands as
// [32 mn x 32 k] boxes in the 128B-swizzle / 32B-atom mode (the only MN-major layout the tensor core
// accepts for 32-bit operands).  fp32 data is consumed as TF32 (mantissa truncated in the datapath), so
// there is no conversion pass and no extra HBM traffic.
//
// Persistent: one CTA per SM walks a static list of (batch, split, m tile, n tile) items.  Roles:
//   warp 0    TMA producer, runs ahead across item boundaries through a multi-stage full/empty mbarrier ring
//   warp 1    TMEM allocator + single-thread MMA issuer; accumulators are DOUBLE-BUFFERED in TMEM
//   warps 2-9 epilogue, two halves of four warps (one warp per TMEM lane quarter in each half): the halves take
//             alternate 128-byte column chunks of the accumulator, each with its own pair of staging boxes, named
//             barrier and TMA-store issuing thread (with one half the epilogue, not HBM, bounded tc_dfeat: 10 items
//             per CTA x ~2.5 us).  tcgen05.ld -> registers -> swizzled smem box -> TMA store (cp.async.bulk.tensor
//             shared->global), so the store of tile i overlaps the loads and MMAs of tile i+1.
//             Outputs whose row pitch is not a multiple of 16 B (logits, C = 309) take a transposing path
//             with 128-byte coalesced stores instead.
// Exact-fp32 tier (TcGemmParams::x3, LF_PREC_FP32 on wide heads): 3xTF32.  Four converter warps sit between the TMA
// producer and the MMA issuer: they split every fp32 word of a landed stage into hi = tf32(x) (written back in place) and
// lo = tf32(x - hi) (a second copy of the stage with the same swizzled layout -- the split is element-wise, so it is
// layout-agnostic and serves K-major and MN-major operands alike), and the issuer runs lo*hi + hi*lo + hi*hi per k-step
// into the same fp32 accumulator.  The dropped lo*lo term and the rounding of lo are ~2^-23 of a product.
// Ragged edges: TMA zero-fills out-of-bounds loads and clips stores, so C = 101/309 and K = 101/309 need
// no padding copies.  Parity class: 2e-2 (tests/test_tc_gemm_gpu.py, tests/test_parity_gpu.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "lf_common.cuh"
#include "lf_peer.cuh"
#include "lf_philox.cuh"
#include "lf_tc.cuh"
#include "lf_tc_ptx.cuh"

namespace lf {

constexpr int kX3ChunkK = 512;                           // 3xTF32: longest K accumulated in TMEM before the tile is flushed (see TcGemmParams::chunks)
constexpr int kConvThreads = 256;                        // 3xTF32: eight converter warps
constexpr int TC_THREADS = 448;                          // producer warp, MMA warp, 2 x 4 epilogue warps; x3: 4 epilogue + 8 converter warps
// staging: two [128 rows x 128 B] store boxes per epilogue half (TcGemmParams::epi_halves = 1 or 2)

struct Item { int batch, split, m0, n0, k_begin, num_kb, k_end, chunk; };

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Tail of the split-K dW GEMM, run by every thread of every CTA after the main loop (TcTail).
__device__ __forceinline__ void gemm_tail(const TcGemmParams& p, float* s_rows) {   // s_rows: >= 384 floats of shared memory
  const TcTail& t = p.tail;
  unsigned long long* tr = t.trace ? t.trace + (size_t)blockIdx.x * 8 : nullptr;
  auto stamp = [&](int k) { if (tr && threadIdx.x == 0) { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x)); tr[k] = x; } };
  stamp(1);
  // ---- grid barrier: every CTA's partial tiles are in global memory (the TMA stores were waited for and fenced)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&t.sync[0], 1u);
    while (ld_acquire_gpu_u32(&t.sync[0]) < gridDim.x) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
  stamp(2);
  const float lr = t.hyper ? t.hyper[0] : 0.f, mom = t.hyper ? t.hyper[1] : 0.f, wd = t.hyper ? t.hyper[2] : 0.f;
  const long long n4 = t.n / 4, total4 = 2 * n4;
  const long long per_cta = (total4 + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per_cta, hi = lo + per_cta < total4 ? lo + per_cta : total4;
  const int n_entries = 2 * t.C + 2 + ((t.peer_on && t.reg_local) ? 1 : 0);      // db1, db2, calibrated counts x2, ranking-loss partial
  const bool peer = t.peer_on != 0;
  const long long epoch = peer ? t.comm.epoch[1] + 1 : 0;
  const int parity = (int)(epoch & 1), G = peer ? t.comm.n_ranks : 1;
  const size_t slot_me = ((size_t)parity * G + (peer ? t.comm.rank : 0)) * (size_t)t.n_padded;

  // SGD on one float4 of head m (torch.optim.SGD: d = g + wd p; buf = mom buf + d; p -= lr buf) + the bf16 copy
  auto sgd4 = [&](int m, long long i, const float g[4], const float4& w4, const float4& b4) {
    const float w0[4] = {w4.x, w4.y, w4.z, w4.w}, m0[4] = {b4.x, b4.y, b4.z, b4.w};
    float np[4], nb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = g[e] + wd * w0[e];
      nb[e] = mom * m0[e] + d;
      np[e] = w0[e] - lr * nb[e];
    }
    *reinterpret_cast<float4*>(t.mom_w[m] + i) = make_float4(nb[0], nb[1], nb[2], nb[3]);
    *reinterpret_cast<float4*>(t.param_w[m] + i) = make_float4(np[0], np[1], np[2], np[3]);
    if (t.w16[m]) {
      const __nv_bfloat162 lo2 = __floats2bfloat162_rn(np[0], np[1]), hi2 = __floats2bfloat162_rn(np[2], np[3]);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&lo2); u.y = *reinterpret_cast<const uint32_t*>(&hi2);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(t.w16[m]) + i) = u;
    }
  };
  // ---- dW: float4 j of [dW1 | dW2] = sum over the splits, in split order; the work is spread over the whole grid
  for (long long j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const int m = j >= n4 ? 1 : 0;
    const long long i = (j - (long long)m * n4) * 4;
    const float* part = (const float*)p.out[m] + i;
    // parameter and momentum of this float4 are fetched with the partials (one round trip for everything)
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
    if (t.hyper && !peer) {
      w4 = *reinterpret_cast<const float4*>(t.param_w[m] + i);
      b4 = *reinterpret_cast<const float4*>(t.mom_w[m] + i);
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= p.splits; k += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(part + (long long)(k + u) * p.split_stride));
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; k < p.splits; ++k) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (long long)k * p.split_stride));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (peer) {                           // this rank's reduced float4 -> every rank's receive area (NVLink stores)
      s.x = clean_f32(s.x); s.y = clean_f32(s.y); s.z = clean_f32(s.z); s.w = clean_f32(s.w);
#pragma unroll
      for (int r = 0; r < LF_MAX_RANKS; ++r)
        if (r < G) reinterpret_cast<float4*>((float*)t.comm.recv_grad[r] + slot_me)[j] = s;
      continue;
    }
    float* o = t.dw[m] + i;               // the flat gradient buffer packs dW2 after db1: not always 16-byte aligned
    const float g[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = g[e];
    if (t.hyper) sgd4(m, i, g, w4, b4);
  }
  stamp(3);
  // ---- db, the calibrated counts (and the ranking-loss partial of a sharded QMF step): column sums of the per-CTA
  // partials of the kernel that produced dz.  One output per CTA and pass: every thread fetches one partial row (a
  // single memory round trip), warp 0 adds them in a fixed order
  auto finish_entry = [&](int i, double s) {          // thread 0: final value of entry i
    if (i >= 2 * t.C + 2) {                            // ranking loss: sum of the ranks' slices
      t.stats[LF_STAT_REG_SUM] = s;
      if (t.loss_out) t.loss_out[0] = t.loss_out[0] + (float)(s / (double)t.batch_global);
      return;
    }
    if (i >= 2 * t.C) { t.stats[LF_STAT_CNT_X1_CAL + (i - 2 * t.C)] = s; return; }
    const int m = i >= t.C ? 1 : 0, c = i - m * t.C;
    const float g = (float)s;
    t.db[m][c] = g;
    if (t.hyper) {
      const float w0 = t.param_b[m][c];
      const float d = g + wd * w0;
      const float bb = mom * t.mom_b[m][c] + d;
      t.mom_b[m][c] = bb;
      t.param_b[m][c] = w0 - lr * bb;
    }
  };
  for (int i = blockIdx.x; i < n_entries; i += gridDim.x) {
    double s = 0.0;
    if (i >= 2 * t.C + 2) {
      s = (double)t.reg_local[0];
    } else {
      const bool is_db = i < 2 * t.C;
      const float* src = is_db ? t.dbpart + i : t.calpart + (i - 2 * t.C);
      const int pitch = is_db ? 2 * t.C : 2, nb = is_db ? t.nb_db : t.nb_cal;
      for (int b0 = 0; b0 < nb; b0 += 384) {
        __syncthreads();
        for (int b = threadIdx.x; b < 384 && b0 + b < nb; b += blockDim.x) s_rows[b] = src[(size_t)(b0 + b) * pitch];
        __syncthreads();
        if (threadIdx.x < 32) {
          // lane l adds rows l, l + 32, ... in order, then a fixed xor butterfly: the same order on every launch
          double q = 0.0;
          for (int b = threadIdx.x; b < 384 && b0 + b < nb; b += 32) q += (double)s_rows[b];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          s += q;
        }
      }
    }
    if (threadIdx.x != 0) continue;
    if (peer) {
#pragma unroll
      for (int r = 0; r < LF_MAX_RANKS; ++r)
        if (r < G) ((float*)t.comm.recv_grad[r] + slot_me)[4 * total4 + i] = clean_f32((float)s);
    } else {
      finish_entry(i, s);
    }
  }
  if (peer) {
    // ---- no fence, no flag: the receive slots were armed with the sentinel, every word read below validates itself
    // (lf_peer.cuh).  Sum of the ranks' chunks in rank order from LOCAL memory (bit-identical on every rank), the slot
    // is re-armed for epoch + 2, then the optimizer.
    stamp(5);
    float* rbase = (float*)t.comm.recv_grad[t.comm.rank] + (size_t)parity * G * (size_t)t.n_padded;
    const uint4 ones = make_uint4(kSentinel32, kSentinel32, kSentinel32, kSentinel32);
    for (long long j = lo + threadIdx.x; j < hi; j += blockDim.x) {
      const int m = j >= n4 ? 1 : 0;
      const long long i = (j - (long long)m * n4) * 4;
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
      if (t.hyper) {
        w4 = *reinterpret_cast<const float4*>(t.param_w[m] + i);
        b4 = *reinterpret_cast<const float4*>(t.mom_w[m] + i);
      }
      uint4 v[LF_MAX_RANKS];
      bool ok[LF_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < LF_MAX_RANKS; ++r) ok[r] = r >= G;
      PeerSpin spin;
      for (;;) {
#pragma unroll
        for (int r = 0; r < LF_MAX_RANKS; ++r)          // every missing rank's load in flight together
          if (!ok[r]) v[r] = ld_volatile_u4(reinterpret_cast<const uint4*>(rbase + (size_t)r * t.n_padded) + j);
        bool all = true;
#pragma unroll
        for (int r = 0; r < LF_MAX_RANKS; ++r) {
          if (!ok[r]) ok[r] = v[r].x != kSentinel32 && v[r].y != kSentinel32 && v[r].z != kSentinel32 && v[r].w != kSentinel32;
          all = all && ok[r];
        }
        if (all) break;
        spin.wait(t.comm.error);
      }
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < LF_MAX_RANKS; ++r)
        if (r < G) {
          s.x += __uint_as_float(v[r].x); s.y += __uint_as_float(v[r].y); s.z += __uint_as_float(v[r].z); s.w += __uint_as_float(v[r].w);
          reinterpret_cast<uint4*>(rbase + (size_t)r * t.n_padded)[j] = ones;
        }
      float* o = t.dw[m] + i;
      const float g[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = g[e];
      if (t.hyper) sgd4(m, i, g, w4, b4);
    }
    stamp(6);
    if (threadIdx.x == 0)
      for (int i = blockIdx.x; i < n_entries; i += gridDim.x) {
        double s = 0.0;
        for (int r = 0; r < G; ++r) {
          unsigned* q = reinterpret_cast<unsigned*>(rbase + (size_t)r * t.n_padded + 4 * total4 + i);
          unsigned u = ld_volatile_u32(q);
          PeerSpin spin;
          while (u == kSentinel32) { spin.wait(t.comm.error); u = ld_volatile_u32(q); }
          s += (double)__uint_as_float(u);
          *q = kSentinel32;
        }
        finish_entry(i, s);
      }
    stamp(7);
  }
  // ---- leave the counters zero for the next launch: the last CTA to depart resets them
  __syncthreads();
  stamp(4);
  if (threadIdx.x == 0) {
    if (atomicAdd(&t.sync[1], 1u) == gridDim.x - 1) {
      t.sync[0] = 0u; t.sync[1] = 0u;
      if (peer) t.comm.epoch[1] = epoch;
      __threadfence();
    }
  }
}

// Epilogue activation of the hidden layers: relu, then inverted dropout.  v holds 8 * N8 consecutive columns of output row
// `row` starting at column col0 (multiple of 8).  Element e = row * N + col draws 16 bits -- half (e & 7) of the Philox block
// with counter (e >> 3, rng_offset) -- and is kept when they are >= round(p * 65536): one Philox call per eight elements.
template <int N8>
__device__ __forceinline__ void relu_dropout(float* v, long long row, int col0, const TcGemmParams& p) {
#pragma unroll
  for (int i = 0; i < 8 * N8; ++i) v[i] = relu_nan(v[i]);
  if (p.drop_p > 0.f) {
    const uint32_t thr = (uint32_t)(p.drop_p * 65536.f + 0.5f);
    const unsigned long long g0 = ((unsigned long long)row * (unsigned long long)p.N + (unsigned long long)col0) >> 3;
#pragma unroll
    for (int g = 0; g < N8; ++g) {
      const uint4 r = dropout_words(g0 + g, p.seed, p.rng_offset);
      const uint32_t wds[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[8 * g + 2 * k] = (wds[k] & 0xffffu) >= thr ? v[8 * g + 2 * k] * p.drop_scale : 0.f;
        v[8 * g + 2 * k + 1] = (wds[k] >> 16) >= thr ? v[8 * g + 2 * k + 1] * p.drop_scale : 0.f;
      }
    }
  }
}

__device__ __forceinline__ Item decode_item(const TcGemmParams& p, int item) {
  Item it;
  const int n_t = item % p.n_tiles; int r = item / p.n_tiles;
  const int m_t = r % p.m_tiles; r /= p.m_tiles;
  it.split = r % p.splits; r /= p.splits;
  it.batch = r % p.nbatch; it.chunk = r / p.nbatch;     // chunks > 1: grid == items per chunk, so a CTA walks one tile's chunks
  it.m0 = m_t * p.tile_m; it.n0 = n_t * p.block_n;
  it.k_begin = it.split * p.k_per_split + it.chunk * p.k_per_chunk;
  it.k_end = min(min(p.K, (it.split + 1) * p.k_per_split), it.k_begin + p.k_per_chunk);
  it.num_kb = it.k_end > it.k_begin ? (it.k_end - it.k_begin + p.kb_elems - 1) / p.kb_elems : 0;
  return it;
}

// The single MMA-issuing thread.  Everything that does not change per instruction is hoisted: the two
// shared-memory descriptors are built once (address field = 0) and advanced with one 64-bit add per k-step
// (the start-address field holds addr >> 4 and never carries out of its 14 bits for a 227 KB window), so the
// issue loop stays far below the tensor pipe's time per instruction (56-128 cycles).
//   K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); 32 B of K per instruction.
//   MN-major: rows are k; fp32: 128B swizzle with 32-byte atoms, 4 k-rows per atom (SBO 512), 32-element MN
//   chunks 4096 B apart (LBO), 1024 B per instruction; bf16: plain 128B swizzle, 8 k-rows per atom (SBO 1024),
//   64-element MN chunks 8192 B apart, 2048 B per instruction.
template <bool TF32, bool X3>
__device__ __forceinline__ void mma_issue_loop(const TcGemmParams& p, uint8_t* smem, uint32_t stage_bytes, uint32_t a_bytes,
                                               uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tmem_full_bar,
                                               uint64_t* tmem_empty_bar, uint32_t tmem_base) {
  // X3: full_bar is the converters' barrier; the lo copies of both operands sit half a stage above the hi ones
  const uint64_t lo_off = (uint64_t)((stage_bytes / 2) >> 4);
  const int stages = p.stages;
  const uint32_t idesc = TF32 ? make_idesc_tf32(TC_BLOCK_M, p.block_n, p.a_mn_major, p.b_mn_major)
                              : make_idesc_bf16(TC_BLOCK_M, p.block_n, p.a_mn_major, p.b_mn_major);
  const uint64_t descA0 = p.a_mn_major ? make_smem_desc(0, p.mn_lbo, p.mn_sbo, p.mn_lt) : make_smem_desc(0, 16, 1024, 2);
  const uint64_t descB0 = p.b_mn_major ? make_smem_desc(0, p.mn_lbo, p.mn_sbo, p.mn_lt) : make_smem_desc(0, 16, 1024, 2);
  const uint64_t stepA = (p.a_mn_major ? p.mn_step : 32u) >> 4, stepB = (p.b_mn_major ? p.mn_step : 32u) >> 4;
  const uint32_t smem0 = smem_u32(smem);
  uint32_t it = 0, li = 0, s = 0, ph = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
    const Item w = decode_item(p, item);
    const uint32_t buf = li & 1;
    mbar_wait(&tmem_empty_bar[buf], ((li >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
    tc_fence_after();
    const uint32_t acc = tmem_base + buf * (uint32_t)p.acc_cols;
    int krem = w.k_end - w.k_begin;
    for (int kb = 0; kb < w.num_kb; ++kb, ++it, krem -= p.kb_elems) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem0 + s * stage_bytes;
      uint64_t da = descA0 + (uint64_t)(sa >> 4), db = descB0 + (uint64_t)((sa + a_bytes) >> 4);
      const int ksteps = krem >= p.kb_elems ? 4 : (krem + p.umma_k - 1) / p.umma_k;       // 32 B of K per instruction
      for (int k = 0; k < ksteps; ++k, da += stepA, db += stepB) {
        if (X3) {                                 // small terms first, all three into the same fp32 accumulator
          umma_tf32(acc, da + lo_off, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_tf32(acc, da, db + lo_off, idesc, 1u);
          umma_tf32(acc, da, db, idesc, 1u);
        } else if (TF32) umma_tf32(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        else umma_f16(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);               // frees the stage once the MMAs above have read it
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
    umma_commit(&tmem_full_bar[buf]);           // accumulator complete
  }
}

// X3: the 3xTF32 variant (converter warps, 448 threads); ACT: the relu / dropout epilogue of the hidden layers.  Separate
// instantiations keep the plain kernel at its own register budget and its epilogue loop free of the Philox code.
template <bool X3, bool ACT>
__global__ void __launch_bounds__(X3 ? TC_THREADS : 320, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapO0, const __grid_constant__ CUtensorMap mapO1, TcGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BLOCK_K * 4;
  const uint32_t half_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  const uint32_t stage_bytes = X3 ? 2 * half_bytes : half_bytes;   // x3: [A hi | B hi | A lo | B lo]
  uint8_t* staging = smem + (size_t)stages * stage_bytes;            // 1024-aligned (stage_bytes is)
  uint64_t* full_bar = (uint64_t*)(staging + (size_t)2 * p.epi_halves * TC_BLOCK_M * 128);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;                      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                      // [2]
  uint64_t* conv_bar = tmem_empty_bar + 2;                           // [stages], x3 only: stage split into hi / lo
  uint32_t* tmem_slot = (uint32_t*)(conv_bar + 8);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (p.tail.trace && threadIdx.x == 0) { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x)); p.tail.trace[(size_t)blockIdx.x * 8] = x; }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapB0);
    if (p.nbatch > 1) { tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB1); }
    if (p.tma_store) { tma_prefetch_desc(&mapO0); if (p.nbatch > 1) tma_prefetch_desc(&mapO1); }
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&conv_bar[s], kConvThreads / 32); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4 * p.epi_halves); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // everything above overlapped the previous kernel's tail; global memory is touched from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;                                // ring position / phase, carried across items
      const uint64_t pol_first = l2_policy_evict_first();
      const int a_boxes = TC_BLOCK_M / p.mn_box, b_boxes = p.block_n / p.mn_box;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item w = decode_item(p, item);
        const CUtensorMap* mapA = w.batch == 0 ? &mapA0 : &mapA1;
        const CUtensorMap* mapB = w.batch == 0 ? &mapB0 : &mapB1;
        for (int kb = 0; kb < w.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          mbar_expect_tx(&full_bar[s], (p.a_mn_major ? a_bytes : (uint32_t)p.tile_m * TC_BLOCK_K * 4) + b_bytes);
          const int k0 = w.k_begin + kb * p.kb_elems;
          const bool ha = (p.l2_last_use & 1) != 0, hb = (p.l2_last_use & 2) != 0;
          if (!p.a_mn_major) {
            if (ha) tma_load_2d_hint(mapA, &full_bar[s], sa, k0, w.m0, pol_first);
            else tma_load_2d(mapA, &full_bar[s], sa, k0, w.m0);                 // [32 k x 128 rows]
          } else {
#pragma unroll
            for (int i = 0; i < a_boxes; ++i) {                                 // [mn_box m x kb_elems k] boxes of 128-B rows
              if (ha) tma_load_2d_hint(mapA, &full_bar[s], sa + i * p.mn_box_bytes, w.m0 + p.mn_box * i, k0, pol_first);
              else tma_load_2d(mapA, &full_bar[s], sa + i * p.mn_box_bytes, w.m0 + p.mn_box * i, k0);
            }
          }
          if (!p.b_mn_major) {
            if (hb) tma_load_2d_hint(mapB, &full_bar[s], sb, k0, w.n0, pol_first);
            else tma_load_2d(mapB, &full_bar[s], sb, k0, w.n0);                 // [32 k x block_n rows]
          } else {
            for (int i = 0; i < b_boxes; ++i) {
              if (hb) tma_load_2d_hint(mapB, &full_bar[s], sb + i * p.mn_box_bytes, w.n0 + p.mn_box * i, k0, pol_first);
              else tma_load_2d(mapB, &full_bar[s], sb + i * p.mn_box_bytes, w.n0 + p.mn_box * i, k0);
            }
          }
          if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      if (X3) mma_issue_loop<true, true>(p, smem, stage_bytes, a_bytes, conv_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
      else if (p.elem == 4) mma_issue_loop<true, false>(p, smem, stage_bytes, a_bytes, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
      else mma_issue_loop<false, false>(p, smem, stage_bytes, a_bytes, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
    }
  } else if (warp < 2 + 4 * p.epi_halves) {
    // ===================== epilogue =====================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;               // which of the two epilogue halves
    const int et = (threadIdx.x - 64) & 127;        // 0..127 within the half
    const int row_in_tile = q * 32 + lane;
    uint32_t li = 0, cc = 0;                        // local item counter, this half's running store-chunk counter
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
      const Item w = decode_item(p, item);
      const uint32_t buf = li & 1;
      mbar_wait_warp(&tmem_full_bar[buf], (li >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * (uint32_t)p.acc_cols + ((uint32_t)(q * 32) << 16);
      const CUtensorMap* mapO = w.batch == 0 ? &mapO0 : &mapO1;
      // a later chunk adds onto what the previous one stored: that store (or reduce) has to be complete, not just read
      if (w.chunk > 0 && et == 0) tma_store_wait_all();
      if (p.tma_store) {
        const int ccols = 128 / p.out_elem;                 // output columns per 128-byte staging row: 32 fp32 / 64 bf16
        for (int c0 = half * ccols; c0 < p.block_n; c0 += p.epi_halves * ccols, ++cc) {
          uint32_t pk[32];                                   // 128 bytes of this thread's output row
          if (p.out_elem == 4) {
            float v[32];
            tmem_ld16(acc + c0, v);
            tmem_ld16(acc + c0 + 16, v + 16);
            const float* bias = w.chunk == 0 ? p.bias[w.batch] : nullptr;
            if (bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) { const int col = w.n0 + c0 + i; v[i] += col < p.N ? __ldg(bias + col) : 0.f; }
            }
            if (ACT) relu_dropout<4>(v, (long long)w.m0 + row_in_tile, w.n0 + c0, p);
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = w.num_kb == 0 ? 0u : __float_as_uint(v[i]);
          } else {
            const float* bias = w.chunk == 0 ? p.bias[w.batch] : nullptr;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              float v[16];
              tmem_ld16(acc + c0 + 16 * h, v);
              if (bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { const int col = w.n0 + c0 + 16 * h + i; v[i] += col < p.N ? __ldg(bias + col) : 0.f; }
              }
              if (ACT) relu_dropout<2>(v, (long long)w.m0 + row_in_tile, w.n0 + c0 + 16 * h, p);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                pk[8 * h + i] = w.num_kb == 0 ? 0u : *reinterpret_cast<const uint32_t*>(&b2);
              }
            }
          }
          uint8_t* box = staging + (half * 2 + (cc & 1)) * (TC_BLOCK_M * 128);
          if (et == 0) tma_store_wait_read<1>();     // the store that last used this box has read it
          named_bar_sync(1 + half, 128);
          uint8_t* rowp = box + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)                 // SWIZZLE_128B: 16-byte chunk j of row r lives at j ^ (r & 7)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async();
          named_bar_sync(1 + half, 128);
          if (et == 0) {
            if (w.chunk > 0) tma_reduce_add_3d(mapO, box, w.n0 + c0, w.m0, w.split);
            else tma_store_3d(mapO, box, w.n0 + c0, w.m0, w.split);
            tma_store_commit();
          }
        }
      } else if (half == 0) {
        // transposing path: 32x32 blocks through a padded per-warp tile, 128 contiguous bytes per store
        float* tile = reinterpret_cast<float*>(staging) + q * (32 * 33);
        const int row0 = w.m0 + q * 32;
        float* out = (float*)p.out[w.batch] + (size_t)w.split * p.split_stride;
        const float* bias = p.bias[w.batch];
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          float v[32];
          tmem_ld16(acc + c0, v);
          if (c0 + 16 < p.block_n) tmem_ld16(acc + c0 + 16, v + 16);
          if (w.num_kb == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = v[i];
          __syncwarp();
          const int col = w.n0 + c0 + lane;
          const bool col_ok = (c0 + lane < p.block_n) && col < p.N;
          const float bv = (bias && col_ok) ? bias[col] : 0.f;
          const int nrows = max(0, min(min(32, p.tile_m - q * 32), p.M - row0));
#pragma unroll 4
          for (int r = 0; r < nrows; ++r)
            if (col_ok) out[(size_t)(row0 + r) * p.ld_out + col] = tile[r * 33 + lane] + bv;
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (p.tma_store && et == 0) {
      tma_store_wait_all();
      if (p.tail.on) asm volatile("fence.proxy.async;" ::: "memory");    // async-proxy stores -> generic-proxy readers of the tail
    }
  }
  else if (X3) {
    // ===================== 3xTF32 converters (warps 6-13; x3 runs one epilogue half) =====================
    const int ct = threadIdx.x - (64 + 128 * p.epi_halves);          // 0..kConvThreads-1
    const uint32_t n16 = half_bytes >> 4;                            // 16-byte words of [A | B] in one stage
    uint32_t s = 0, ph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const Item w = decode_item(p, item);
      for (int kb = 0; kb < w.num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);               // every lane polls: measured faster than one polling lane + __syncwarp here
        uint4* hi = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes);
        uint4* lo = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + half_bytes);
        // loads of a whole batch first: the tensor core's operand reads keep the shared-memory port busy, so an LDS takes
        // hundreds of cycles here, and the compiler cannot hoist loads over the (possibly aliasing) stores by itself
        for (uint32_t i0 = ct; i0 < n16; i0 += kConvThreads * 8) {
          uint4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { const uint32_t i = i0 + j * kConvThreads; if (i < n16) x[j] = hi[i]; }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t i = i0 + j * kConvThreads;
            if (i < n16) {
              uint4 h, l;
              split_tf32(x[j].x, h.x, l.x); split_tf32(x[j].y, h.y, l.y); split_tf32(x[j].z, h.z, l.z); split_tf32(x[j].w, h.w, l.w);
              hi[i] = h; lo[i] = l;
            }
          }
        }
        fence_proxy_async();                       // generic-proxy writes -> the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[s]);
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (p.tail.on) gemm_tail(p, reinterpret_cast<float*>(staging));      // the store boxes are idle by now
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Encoded tensor maps are cached per thread, keyed by every argument of the encode call: an eager training step issues
// 16 of them over the same handful of buffers, and cuTensorMapEncodeTiled (~1-2 us of host time each) was a visible
// part of the host cost of a step that takes 150 us on the device.
struct MapKey {
  const void* base; long long d0, d1, d2, s0, s1; int b0, b1, b2, elem, swz, rank;
  bool operator==(const MapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s0 == o.s0 && s1 == o.s1 && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && elem == o.elem && swz == o.swz && rank == o.rank;
  }
};
struct MapCache {
  static constexpr int N = 64;
  MapKey key[N]; CUtensorMap val[N]; int used = 0, next = 0;
  const CUtensorMap* find(const MapKey& k) const {
    for (int i = 0; i < used; ++i) if (key[i] == k) return &val[i];
    return nullptr;
  }
  void put(const MapKey& k, const CUtensorMap& v) {
    const int i = used < N ? used++ : (next = (next + 1) % N);
    key[i] = k; val[i] = v;
  }
};
static thread_local MapCache g_maps;

// 2-D tensor map: `inner` contiguous elements per row, `outer` rows, row pitch ld elements.
int make_map(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld,
             int box_inner, int box_outer, bool mn_major, int elem, int mn_swizzle) {
  const MapKey mk{base, inner, outer, 0, ld, 0, box_inner, box_outer, 0, elem, mn_major ? 1000 + mn_swizzle : 0, 2};
  if (const CUtensorMap* hit = g_maps.find(mk)) { *map = *hit; return LF_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LF_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ld * elem) % 16) { set_error("TMA operand must be 16-byte aligned with a 16-byte row pitch"); return LF_ERR_BAD_ARG; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (mn_major) swz = mn_swizzle ? (CUtensorMapSwizzle)mn_swizzle : (elem == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return LF_ERR_CUDA; }
  g_maps.put(mk, *map);
  return LF_OK;
}

// 3-D output map for the TMA-store epilogue: (cols, rows, splits), box [32 x 128 x 1], SWIZZLE_128B.
int make_store_map(CUtensorMap* map, void* base, long long cols, long long rows, long long ld,
                          long long splits, long long split_stride, int elem, int tile_m) {
  const MapKey mk{base, cols, rows, splits, ld, split_stride, 128 / elem, tile_m, 1, elem, 0, 3};
  if (const CUtensorMap* hit = g_maps.find(mk)) { *map = *hit; return LF_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LF_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)splits};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem, (cuuint64_t)(splits > 1 ? split_stride : rows * ld) * elem};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem), (cuuint32_t)tile_m, 1};      // balanced M tiles store only their own rows
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(store) failed (%d)", (int)r); return LF_ERR_CUDA; }
  g_maps.put(mk, *map);
  return LF_OK;
}

int tc_gemm(const TcGemmDesc& d, cudaStream_t s) {
  if (d.block_n % 16 || d.block_n < 16 || d.block_n > 256 || (d.b_mn_major && d.block_n % 32)) {
    set_error("tc_gemm: bad block_n %d", d.block_n);
    return LF_ERR_BAD_ARG;
  }
  TcGemmParams p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.elem = d.elem == 2 ? 2 : 4;
  p.x3 = (d.x3 && p.elem == 4) ? 1 : 0;
  p.out_elem = d.out_elem == 2 ? 2 : 4;
  p.kb_elems = 128 / p.elem; p.umma_k = 32 / p.elem; p.mn_box = 128 / p.elem;
  p.mn_box_bytes = 128 * p.kb_elems;
  if (p.elem == 4) { p.mn_step = 1024; p.mn_lbo = 4096; p.mn_sbo = 512; p.mn_lt = 1; }     // 128B swizzle, 32-byte atoms
  else { p.mn_step = 2048; p.mn_lbo = 8192; p.mn_sbo = 1024; p.mn_lt = 2; }                // plain 128B swizzle
  const int mn_swz = 0;
  if ((d.b_mn_major && d.block_n % p.mn_box) || (p.out_elem == 2 && d.block_n % 64)) { set_error("tc_gemm: block_n %d not a multiple of the box width", d.block_n); return LF_ERR_BAD_ARG; }
  p.block_n = d.block_n;
  p.nbatch = d.nbatch;
  p.splits = d.splits < 1 ? 1 : d.splits;
  int kps = div_up(d.K, p.splits);
  kps = div_up(kps, p.kb_elems) * p.kb_elems;
  p.k_per_split = kps;
  p.chunks = 1; p.k_per_chunk = kps;
  p.a_mn_major = d.a_mn_major; p.b_mn_major = d.b_mn_major;
  p.l2_last_use = d.l2_last_use;
  for (int b = 0; b < 2; ++b) { p.out[b] = d.out[b < d.nbatch ? b : 0]; p.bias[b] = d.bias[b < d.nbatch ? b : 0]; }
  p.ld_out = d.ld_out; p.split_stride = d.split_stride;
  // TMA-store epilogue: 16-byte output pitch; the 128-byte store chunks must tile the N tile exactly unless it is
  // the only N tile (then the tensor map clips the overhang); bias only for fp32 output
  const int ccols = 128 / p.out_elem;
  bool aligned = ((d.ld_out * p.out_elem) % 16 == 0) && ((d.split_stride * p.out_elem) % 16 == 0) &&
                 (d.block_n % ccols == 0 || d.N <= d.block_n);
  for (int b = 0; b < d.nbatch; ++b) aligned = aligned && (((uintptr_t)d.out[b] & 15) == 0);
  p.tma_store = aligned ? 1 : 0;
  p.act_relu = d.act_relu; p.drop_p = d.drop_p; p.drop_scale = d.drop_p > 0.f ? 1.f / (1.f - d.drop_p) : 1.f;
  p.seed = d.seed; p.rng_offset = d.rng_offset;
  if (d.act_relu && (!p.tma_store || (d.N & 7) || d.splits > 1 || d.drop_p < 0.f || d.drop_p >= 1.f)) {
    set_error("tc_gemm: the relu / dropout epilogue needs the TMA-store path, N %% 8 == 0, no split-K and 0 <= p < 1");
    return LF_ERR_BAD_ARG;
  }
  if (p.out_elem == 2 && !p.tma_store) { set_error("tc_gemm: bf16 output needs the TMA-store epilogue (16-byte pitch, no bias)"); return LF_ERR_BAD_ARG; }
  p.acc_cols = d.block_n <= 32 ? 32 : d.block_n <= 64 ? 64 : d.block_n <= 128 ? 128 : 256;
  p.tmem_cols = 2 * p.acc_cols;
  p.tile_m = TC_BLOCK_M;
  p.n_tiles = div_up(d.N, d.block_n);
  if (d.balance_m && !d.a_mn_major) {
    // smallest number of M tiles >= ceil(M/128) that makes items a multiple of 148, if it costs < 15% extra tiles
    const int per_m = p.n_tiles * p.splits * d.nbatch;
    const int base = div_up(d.M, TC_BLOCK_M);
    for (int mt = base; mt <= base + base / 6 + 1; ++mt)
      if ((mt * per_m) % 148 == 0 || (mt * per_m) % 148 > 140) {
        const int tm = div_up(div_up(d.M, mt), 8) * 8;
        if (tm <= TC_BLOCK_M && tm >= 64) { p.tile_m = tm; break; }
      }
  }
  p.m_tiles = div_up(d.M, p.tile_m);
  p.total_items = p.m_tiles * p.n_tiles * p.splits * d.nbatch;
  int grid_fixed = 0;
  if (p.x3 && kps > kX3ChunkK && p.tma_store && p.total_items <= 148 && !d.act_relu) {     // (a nonlinear epilogue needs the whole sum)
    // 3xTF32 with a long K per work item (the split-K dW GEMM): bound the accumulation chain in TMEM
    p.chunks = div_up(kps, kX3ChunkK);
    p.k_per_chunk = div_up(div_up(kps, p.chunks), p.kb_elems) * p.kb_elems;
    grid_fixed = p.total_items;
    p.total_items *= p.chunks;
  }

  CUtensorMap mA[2], mB[2], mO[2];
  for (int b = 0; b < d.nbatch; ++b) {
    int rc;
    // K-major: inner = K, outer = MN rows, box [32 x rows].  MN-major: inner = MN, outer = K rows, box [32 x 32].
    rc = d.a_mn_major ? make_map(&mA[b], d.A[b], d.M, d.K, d.lda, p.mn_box, p.kb_elems, true, p.elem, mn_swz)
                      : make_map(&mA[b], d.A[b], d.K, d.M, d.lda, p.kb_elems, p.tile_m, false, p.elem);
    if (rc) return rc;
    rc = d.b_mn_major ? make_map(&mB[b], d.B[b], d.N, d.K, d.ldb, p.mn_box, p.kb_elems, true, p.elem, mn_swz)
                      : make_map(&mB[b], d.B[b], d.K, d.N, d.ldb, p.kb_elems, d.block_n, false, p.elem);
    if (rc) return rc;
    if (p.tma_store) {
      rc = make_store_map(&mO[b], d.out[b], d.N, d.M, d.ld_out, p.splits, d.split_stride, p.out_elem, p.tile_m);
      if (rc) return rc;
    } else {
      mO[b] = mA[b];
    }
  }
  if (d.nbatch == 1) { mA[1] = mA[0]; mB[1] = mB[0]; mO[1] = mO[0]; }

  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)d.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = (a_bytes + ((b_bytes + 1023) & ~1023u)) * (p.x3 ? 2 : 1);
  // the second epilogue half pays when the epilogue bounds the kernel: short K (tc_dfeat of a <= 128-way head: ten
  // 128 x 256 output tiles per CTA, two k-blocks each); with long K its two extra staging boxes cost a pipeline stage
  // (x3: warps 6-9 are the converters, so one half)
  // (the activation epilogue -- Philox per eight outputs -- is the longer side of an item even at K = 768: two halves too)
  p.epi_halves = (p.tma_store && (d.K <= 128 || d.act_relu) && d.max_epi_halves >= 2 && !p.x3) ? 2 : 1;
  const size_t staging_bytes = (size_t)2 * p.epi_halves * TC_BLOCK_M * 128;
  const size_t fixed = staging_bytes + 320;
  int stages = 8;
  while (stages > 2 && (size_t)stages * stage_bytes + fixed > 226 * 1024) --stages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcGemmParams);
  static const KernelFn kernels[4] = {tc_gemm_kernel<false, false>, tc_gemm_kernel<false, true>, tc_gemm_kernel<true, false>, tc_gemm_kernel<true, true>};
  const KernelFn kernel = kernels[2 * p.x3 + (p.act_relu ? 1 : 0)];
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 4; ++i) cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    attr_set = true;
  }
  const int grid = grid_fixed ? grid_fixed : (p.total_items < 148 ? p.total_items : 148);
  p.tail = d.tail;
  static unsigned long long* trace_buf = nullptr;
  static int trace_calls = 0;
  if (p.tail.on && getenv("LF_DW_TRACE")) {
    if (!trace_buf) { cudaMalloc(&trace_buf, 148 * 8 * sizeof(unsigned long long)); cudaMemset(trace_buf, 0, 148 * 8 * sizeof(unsigned long long)); }
    p.tail.trace = trace_buf;
  }
  if (p.tail.on) {
    // the tail's grid barrier needs every CTA resident at once: a cooperative launch guarantees it (or fails)
    if (!p.tma_store || p.tail.n % 4) { set_error("tc_gemm: the fused dW tail needs the TMA-store epilogue and C*D %% 4 == 0"); return LF_ERR_BAD_ARG; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 128 * p.epi_halves + kConvThreads * p.x3); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = getenv("LF_DW_NOCOOP") ? 0 : 1;
    cudaError_t e = cudaSuccess;
    LF_LAUNCH(d.name, s, (e = cudaLaunchKernelEx(&cfg, kernel, mA[0], mB[0], mA[1], mB[1], mO[0], mO[1], p)));
    if (e != cudaSuccess) { set_error("%s: %s", d.name, cudaGetErrorString(e)); return LF_ERR_CUDA; }
    if (p.tail.trace && ++trace_calls == 12) {
      cudaStreamSynchronize(s);
      static unsigned long long h[148 * 8];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull;
      for (int c = 0; c < grid; ++c) if (h[c * 8] < t0) t0 = h[c * 8];
      const char* nm[8] = {"entry", "gemm_done", "barrier_done", "dw_reduced", "end", "entries_pushed", "chunks_summed", "entries_summed"};
      for (int k = 0; k < 8; ++k) {
        if (k >= 5 && !p.tail.peer_on) break;
        double mn = 1e30, mx = 0, sum = 0;
        for (int c = 0; c < grid; ++c) { const double v = (double)(h[c * 8 + k] - t0) / 1000.0; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
        fprintf(stderr, "[dw trace] %-13s min %7.2f  avg %7.2f  max %7.2f us\n", nm[k], mn, sum / grid, mx);
      }
    }
    return check_launch(d.name);
  }
  LF_LAUNCH(d.name, s, launch_pdl(kernel, dim3(grid), dim3(64 + 128 * p.epi_halves + kConvThreads * p.x3), smem, s, mA[0], mB[0], mA[1], mB[1], mO[0], mO[1], p));
  return check_launch(d.name);
}

}  // namespace lf

// Test hook (no reference counterpart): one tensor-pipe GEMM through the C ABI so the kernel can be
// checked in isolation.  out(M,N) = A*B (+bias) with the operand orientations described in lf_tc.cuh.
// Same hook for bf16 operands (A, B bf16; out fp32, or bf16 when out_bf16 != 0).
extern "C" int lf_debug_tc_gemm16(const void* A, const void* B, void* out, int32_t M, int32_t N, int32_t K, int32_t lda,
                                  int32_t ldb, int32_t ld_out, int32_t a_mn_major, int32_t b_mn_major, int32_t block_n,
                                  int32_t splits, int64_t split_stride, int32_t out_bf16, void* stream) {
  lf::TcGemmDesc d;
  d.nbatch = 1;
  d.A[0] = d.A[1] = A; d.B[0] = d.B[1] = B; d.bias[0] = d.bias[1] = nullptr; d.out[0] = d.out[1] = out;
  d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ld_out = ld_out;
  d.a_mn_major = a_mn_major; d.b_mn_major = b_mn_major; d.block_n = block_n;
  d.splits = splits; d.split_stride = split_stride; d.elem = 2; d.out_elem = out_bf16 ? 2 : 4; d.name = "tc_gemm16_debug";
  return lf::tc_gemm(d, (cudaStream_t)stream);
}

static int debug_tc_gemm32(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                           int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                           int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride,
                           void* stream, int x3) {
  lf::TcGemmDesc d;
  d.x3 = x3;
  d.nbatch = 1;
  d.A[0] = A; d.B[0] = B; d.bias[0] = bias; d.out[0] = out;
  d.A[1] = A; d.B[1] = B; d.bias[1] = bias; d.out[1] = out;
  d.elem = 4; d.out_elem = 4;
  d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ld_out = ld_out;
  d.a_mn_major = a_mn_major; d.b_mn_major = b_mn_major; d.block_n = block_n;
  d.splits = splits; d.split_stride = split_stride; d.balance_m = 0; d.name = x3 ? "tc_gemm_x3_debug" : "tc_gemm_debug";
  return lf::tc_gemm(d, (cudaStream_t)stream);
}

extern "C" int lf_debug_tc_gemm(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                                int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                                int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride,
                                void* stream) {
  return debug_tc_gemm32(A, B, bias, out, M, N, K, lda, ldb, ld_out, a_mn_major, b_mn_major, block_n, splits, split_stride, stream, 0);
}

// Same hook with the 3xTF32 operand split (the exact-fp32 tier of the wide heads).
extern "C" int lf_debug_tc_gemm_x3(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                                   int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                                   int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride,
                                   void* stream) {
  return debug_tc_gemm32(A, B, bias, out, M, N, K, lda, ldb, ld_out, a_mn_major, b_mn_major, block_n, splits, split_stride, stream, 1);
}
