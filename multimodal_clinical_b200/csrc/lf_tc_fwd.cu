// Fused QMF forward for wide heads (32 <= C <= 128): logits of BOTH modalities and the whole per-sample row math
// in ONE persistent tcgen05 kernel, so `rows_forward` and its re-read of z1 / z2 disappear and the row math runs in
// the issue slots the bandwidth-bound GEMM leaves idle (the GEMM's six warps use < 20 % of them).
//
//   work item  = one M tile of 128 samples, both modalities: z_m = F_m W_m^T accumulate into TMEM columns
//                [buf][m][128] (2 buffers x 2 modalities x 128 columns = all 512 columns), so the epilogue of tile i
//                overlaps the TMA loads and MMAs of tile i+1
//   warp 0     TMA producer (F_m tile [128 x 128 B] + W_m tile [block_n x 128 B] per stage, multi-stage mbarrier ring)
//   warp 1     TMEM allocator + single-thread MMA issuer (kind::f16 on bf16 operands, kind::tf32 on fp32 operands)
//   warps 2..  epilogue, 4 x H warps: warp (q, h) owns the samples of TMEM lane quarter q (one thread per sample, so
//              reductions over classes are loops in registers, no shuffles) and the h-th share of the 16-column
//              chunks; the H warps of a quarter combine their per-sample partial max / argmax / sums through shared
//              memory at the end of each pass (64 B per sample, one named barrier).  H = 4 keeps four epilogue warps
//              per scheduler: with one (H = 1, the first version) every dependent instruction exposed its full latency
//              and the fully unrolled 9 K-instruction epilogue thrashed the instruction cache (75 us for K4).
//              Two passes over the chunks of a row, loops NOT unrolled across chunks, softmax denominators online
//              (running max, sum rescaled when the max moves):
//                1  z_m = acc + bias -> HBM (128-bit stores), argmax and sum exp of z1 / z2, avg = (z1 + z2) / 2 -> HBM,
//                   argmax(avg), column sums for the EMA (the only cross-sample reduction: a 16-shuffle transposing
//                   butterfly per chunk)
//                2  z_df = z1 c1 + z2 c2 -> HBM, argmax, sum exp
//              TMEM is re-read in pass 2 instead of holding 2 x 112 values in registers.
//
// Outputs and partial-statistics layout are exactly those of tc_logits + rows_forward (lf_rows_vec.cu), so
// finalize_stats, step_mid and the backward pass are unchanged.  Reference arithmetic: lf_rows.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "lf_common.cuh"
#include "lf_rowmath.cuh"
#include "lf_rows.cuh"
#include "lf_tc.cuh"
#include "lf_tc_ptx.cuh"

namespace lf {

constexpr float kFwLog2e = 1.4426950408889634f;

struct TcFwdParams {
  int B, C, D;
  int ld_z, ld_f;          // row pitches of z1/z2 and of avg/z_df (multiples of 4)
  int block_n;             // 16 * NCH
  int stages, elem, kb_elems, num_kb, m_tiles;
  int tile_m;              // samples per M tile (<= 128, multiple of 8): chosen so that the tiles fill whole waves of CTAs
  int nb_total;            // partial rows the finalize kernel sums; rows beyond the grid are zeroed here
  int l2_hints;            // bit 0: features evict_last, bit 1: avg / z_df stores streaming (lf_tc_ptx.cuh)
  unsigned long long* trace;   // LF_FWD_TRACE=1: [grid][8] %globaltimer stamps of the roles (printed once, see tc_heads_forward_qmf)
  const float* bias[2];
  float* z[2];
  float* avg;
  float* zdf;
  float* conf;             // (2,B)
  float* rowstat;          // (B,4): lse1, lse2, lse(z_df), -
  const int64_t* label;
  float* partials;         // [nb_total][stat_len]
  double* stats;           // when sync != null: the LAST CTA to finish sums the per-CTA partial rows into stats (fixed order)
  unsigned* sync;          // zero-initialised arrival counter, left zero
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fw_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// two 16-column TMEM loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float* a, float* b) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
}
__device__ __forceinline__ void add_bias16(float (&v)[16], const float* __restrict__ sb) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(sb + 4 * j);
    v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
  }
}
// columns >= C := -inf (only the chunk that straddles C and the ones after it take the selects)
__device__ __forceinline__ void mask16(float (&v)[16], int c0, int C) {
  if (c0 + 15 >= C) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i >= C) v[i] = -INFINITY;
  }
}
// Coalesced store of one 16-column chunk of the warp's 32 samples.  A thread holds one sample's 64 bytes; stored
// directly, a warp instruction would scatter 32 half-sectors over 32 cache lines (measured: the store path, not the
// arithmetic, bounded the first versions of this kernel).  The chunk is transposed through a 2 KB per-warp staging
// tile (16-byte pieces XOR-swizzled by the row so both directions are conflict-free) and written as 8 rows x 64
// contiguous bytes per instruction.
template <bool STREAM = false>
__device__ __forceinline__ void store16_coalesced(float* stg, float* __restrict__ base, int ld, int row0, int B, int c0,
                                                  const float (&v)[16], int lane) {
  float4* mine = reinterpret_cast<float4*>(stg + lane * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) mine[j ^ ((lane >> 1) & 3)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int pc = lane & 3, col = c0 + 4 * pc;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (lane >> 2) + 8 * k;
    const float4 t = reinterpret_cast<const float4*>(stg + r * 16)[pc ^ ((r >> 1) & 3)];
    const int row = row0 + r;
    if (row < B && col < ld) {
      if (STREAM) __stcs(reinterpret_cast<float4*>(base + (size_t)row * ld + col), t);       // never read again in the step
      else *reinterpret_cast<float4*>(base + (size_t)row * ld + col) = t;
    }
  }
  __syncwarp();
}
// running max / first argmax over one chunk: trees instead of a 16-deep compare-select chain (the epilogue is
// latency-bound: four warps per scheduler, each a chain of dependent steps)
__device__ __forceinline__ void max_arg16(const float (&v)[16], int c0, float& mx, int& arg) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = fmaxf(v[2 * i], v[2 * i + 1]);
  const float cm = fmaxf(fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])), fmaxf(fmaxf(t[4], t[5]), fmaxf(t[6], t[7])));
  int e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = (v[i] == cm) ? i : 16;
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) e[i] = min(e[i], e[i + w]);
  const bool up = cm > mx;                                  // strict: an earlier chunk keeps a tie (torch.argmax: first index)
  mx = up ? cm : mx;
  arg = up ? c0 + e[0] : arg;
}

template <int NCH, int H>
__global__ void __launch_bounds__(64 + 128 * H, 1)
tc_fwd_qmf_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1, TcFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  const uint32_t a_bytes = TC_BLOCK_M * 128;
  const uint32_t b_bytes = (uint32_t)p.block_n * 128;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  uint8_t* tail = smem + (size_t)stages * stage_bytes;
  constexpr int EW = 4 * H;                                      // epilogue warps
  constexpr int FW_THREADS = 64 + 128 * H;
  float* sbias = reinterpret_cast<float*>(tail);                 // [2][128], zero beyond C
  float4* xch = reinterpret_cast<float4*>(sbias + 256);          // [2 parities][4 quarters][H][32 lanes] per-sample partials
  float* sred = reinterpret_cast<float*>(xch + 2 * 4 * H * 32);  // [EW warps][2][128] column sums, then [EW][16] statistics
  float* sstg = sred + EW * 256 + EW * 16;                       // [EW warps][32 samples][16 columns] store staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sstg + EW * 512);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;                  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int C = p.C, B = p.B;
  unsigned long long* tr = p.trace ? p.trace + (size_t)blockIdx.x * 8 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = gtimer();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapW0); tma_prefetch_desc(&mapW1);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  for (int i = threadIdx.x; i < 256; i += FW_THREADS) {
    const int m = i >> 7, c = i & 127;
    sbias[i] = c < C ? p.bias[m][c] : 0.f;
  }
  for (int i = threadIdx.x; i < EW * 256; i += FW_THREADS) sred[i] = 0.f;
  __syncthreads();

  if (tr && threadIdx.x == 0) tr[1] = gtimer();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const uint64_t pol_last = l2_policy_evict_last();
      bool first = true;
      // (tried: cp.async.bulk.prefetch.tensor of the feature boxes one / two tiles ahead into L2 -- the main loop got
      // 10 % slower, 29 vs 26 us to the last load)
      for (int item = blockIdx.x; item < p.m_tiles; item += gridDim.x) {
        const int m0 = item * p.tile_m;
        for (int m = 0; m < 2; ++m) {
          const CUtensorMap* mapA = m == 0 ? &mapA0 : &mapA1;
          const CUtensorMap* mapW = m == 0 ? &mapW0 : &mapW1;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + (size_t)s * stage_bytes;
            mbar_expect_tx(&full_bar[s], (uint32_t)p.tile_m * 128u + b_bytes);
            if (p.l2_hints & 1) tma_load_2d_hint(mapA, &full_bar[s], sa, kb * p.kb_elems, m0, pol_last);
            else tma_load_2d(mapA, &full_bar[s], sa, kb * p.kb_elems, m0);       // [128 B of K x tile_m samples]
            tma_load_2d(mapW, &full_bar[s], sa + a_bytes, kb * p.kb_elems, 0);   // [128 B of K x block_n classes]
            if (tr && first) { tr[2] = gtimer(); first = false; }
            if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
          }
        }
      }
      if (tr) tr[3] = gtimer();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = p.elem == 4 ? make_idesc_tf32(TC_BLOCK_M, p.block_n, 0, 0) : make_idesc_bf16(TC_BLOCK_M, p.block_n, 0, 0);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, 2);          // K-major SWIZZLE_128B, 8-row groups 1024 B apart
      const uint32_t smem0 = smem_u32(smem);
      uint32_t li = 0, s = 0, ph = 0;
      for (int item = blockIdx.x; item < p.m_tiles; item += gridDim.x, ++li) {
        const uint32_t buf = li & 1;
        mbar_wait(&tmem_empty_bar[buf], ((li >> 1) & 1) ^ 1);       // the epilogue has drained this buffer
        tc_fence_after();
        for (int m = 0; m < 2; ++m) {
          const uint32_t acc = tmem_base + (buf * 2 + m) * 128;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[s], ph);
            if (tr && li == 0 && m == 0 && kb == 0) tr[4] = gtimer();
            tc_fence_after();
            const uint32_t sa = smem0 + s * stage_bytes;
            uint64_t da = desc0 + (uint64_t)(sa >> 4), db = desc0 + (uint64_t)((sa + a_bytes) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k, da += 2, db += 2) {          // 32 B of K per instruction (zero-filled past D)
              if (p.elem == 4) umma_tf32(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_f16(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
            if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[buf]);                           // both accumulators of the tile complete
      }
      if (tr) tr[5] = gtimer();
    }
  } else {
    // ===================== epilogue: 4 quarters x H column shares =====================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int ew = warp - 2;                         // 0 .. 4H-1
    const int h = ew >> 2;                           // share of the chunks
    constexpr int PER = (NCH + H - 1) / H;
    const int ch_lo = h * PER, ch_hi = min(NCH, ch_lo + PER);
    float* mycol = sred + ew * 256;                  // this warp's column sums [2][128]
    float* stg = sstg + ew * 512;
    float st[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) st[i] = 0.f;
    uint32_t xk = 0;                                 // exchange counter (slot parity)
    // all-gather of one float4 per sample among the H warps of this quarter
    auto exchange = [&](const float4& mine) -> const float4* {
      float4* slot = xch + ((xk & 1) * 4 + q) * (H * 32);
      ++xk;
      if (H > 1) {
        slot[h * 32 + lane] = mine;
        named_bar_sync(2 + q, 32 * H);
      }
      return slot + lane;                            // [hh * 32] = share hh of this sample
    };

    uint32_t li = 0;
    for (int item = blockIdx.x; item < p.m_tiles; item += gridDim.x, ++li) {
      const uint32_t buf = li & 1;
      const int row0 = item * p.tile_m + q * 32;
      const int row = row0 + lane;
      const int row_end = min(B, (item + 1) * p.tile_m);      // TMEM lanes past the tile hold stale rows: never stored
      const bool valid = row < row_end;
      const int y = (int)p.label[valid ? row : B - 1];
      const bool yok = (unsigned)y < (unsigned)C;
      mbar_wait_warp(&tmem_full_bar[buf], (li >> 1) & 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + (buf * 2) * 128 + ((uint32_t)(q * 32) << 16), acc1 = acc0 + 128;

      // ---- pass 1: logits and mean-fused logits out, EMA column sums, argmax of z1 / z2 / avg, and the softmax
      // denominators of z1 / z2 ONLINE (running max, sum rescaled when the max moves): one read of the accumulators
      float m1 = -INFINITY, m2 = -INFINITY, ma = -INFINITY, s1 = 0.f, s2 = 0.f;
      int i1 = 0x7fffffff, i2 = 0x7fffffff, ia = 0x7fffffff;
#pragma unroll 1
      for (int ch = ch_lo; ch < ch_hi; ++ch) {
        const int c0 = ch * 16;
        float v1[16], v2[16];
        tmem_ld16x2(acc0 + c0, acc1 + c0, v1, v2);
        add_bias16(v1, sbias + c0); add_bias16(v2, sbias + 128 + c0);
        store16_coalesced(stg, p.z[0], p.ld_z, row0, row_end, c0, v1, lane);
        store16_coalesced(stg, p.z[1], p.ld_z, row0, row_end, c0, v2, lane);
        {
          float t1[16], t2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { t1[i] = valid ? v1[i] : 0.f; t2[i] = valid ? v2[i] : 0.f; }
          const float r1 = transpose_reduce<16>(t1, lane);      // lanes 2c, 2c+1: column c0 + c over this warp's 32 samples
          const float r2 = transpose_reduce<16>(t2, lane);
          if ((lane & 1) == 0) { mycol[c0 + (lane >> 1)] += r1; mycol[128 + c0 + (lane >> 1)] += r2; }
        }
        {
          float av[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) av[i] = (v1[i] + v2[i]) / 2.f;
          if (p.l2_hints & 2) store16_coalesced<true>(stg, p.avg, p.ld_f, row0, row_end, c0, av, lane);
          else store16_coalesced(stg, p.avg, p.ld_f, row0, row_end, c0, av, lane);
          mask16(av, c0, C);
          max_arg16(av, c0, ma, ia);
        }
        mask16(v1, c0, C); mask16(v2, c0, C);
        const float o1 = m1, o2 = m2;
        max_arg16(v1, c0, m1, i1); max_arg16(v2, c0, m2, i2);
        const float k1 = m1 * kFwLog2e, k2 = m2 * kFwLog2e;
        s1 *= exp_sub(o1, k1); s2 *= exp_sub(o2, k2);          // exp(old max - new max): 1 when the max did not move, 0 at the start
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1 += exp_sub(v1[i], k1); s2 += exp_sub(v2[i], k2); }
      }
      if (H > 1) {
        // a slot is reused two exchanges later: its reads must precede the NEXT exchange's barrier, so the first
        // gather is copied to registers before the second one starts
        const float4* oa = exchange(make_float4(m1, s1, __int_as_float(i1), ma));
        float4 ta[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) ta[hh] = oa[hh * 32];
        const float4* ob = exchange(make_float4(m2, s2, __int_as_float(i2), __int_as_float(ia)));
        float4 tb[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) tb[hh] = ob[hh * 32];
        float M1 = -INFINITY, M2 = -INFINITY;
        ma = -INFINITY; i1 = 0x7fffffff; i2 = 0x7fffffff; ia = 0x7fffffff;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {             // shares are ordered by column: strict > keeps the first index
          if (ta[hh].x > M1) { M1 = ta[hh].x; i1 = __float_as_int(ta[hh].z); }
          if (tb[hh].x > M2) { M2 = tb[hh].x; i2 = __float_as_int(tb[hh].z); }
          if (ta[hh].w > ma) { ma = ta[hh].w; ia = __float_as_int(tb[hh].w); }
        }
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {             // fixed order: every warp of the quarter gets the same sums
          s1 += ta[hh].y * exp_sub(ta[hh].x, M1 * kFwLog2e); s2 += tb[hh].y * exp_sub(tb[hh].x, M2 * kFwLog2e);   // empty share: 0 * 0
        }
        m1 = M1; m2 = M2;
      }
      const float lse1 = m1 + __logf(s1), lse2 = m2 + __logf(s2);
      // energy = log(sum(exp z)) is NOT stabilised in the reference (QMF.py:113): same value unless the plain fp32
      // sum overflows, where the reference yields +inf
      const float c1 = (lse1 > 88.72283f ? INFINITY : lse1) / 10.f;
      const float c2 = (lse2 > 88.72283f ? INFINITY : lse2) / 10.f;
      // ---- pass 2: energy-fused logits out, argmax, online softmax denominator
      float md = -INFINITY, sd = 0.f;
      int idf = 0x7fffffff;
#pragma unroll 1
      for (int ch = ch_lo; ch < ch_hi; ++ch) {
        const int c0 = ch * 16;
        float v1[16], v2[16], vd[16];
        tmem_ld16x2(acc0 + c0, acc1 + c0, v1, v2);
        add_bias16(v1, sbias + c0); add_bias16(v2, sbias + 128 + c0);
#pragma unroll
        for (int i = 0; i < 16; ++i) vd[i] = v1[i] * c1 + v2[i] * c2;
        if (p.l2_hints & 2) store16_coalesced<true>(stg, p.zdf, p.ld_f, row0, row_end, c0, vd, lane);
        else store16_coalesced(stg, p.zdf, p.ld_f, row0, row_end, c0, vd, lane);
        mask16(vd, c0, C);
        const float od = md;
        max_arg16(vd, c0, md, idf);
        const float kd = md * kFwLog2e;
        sd *= exp_sub(od, kd);
#pragma unroll
        for (int i = 0; i < 16; ++i) sd += exp_sub(vd[i], kd);
      }
      // the accumulators are drained: hand the TMEM buffer back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) fw_mbar_arrive(&tmem_empty_bar[buf]);
      if (H > 1) {
        // the barrier of this exchange also orders the other shares' stores of z (read back below) before the reads
        const float4* o = exchange(make_float4(md, sd, __int_as_float(idf), 0.f));
        float4 t[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) t[hh] = o[hh * 32];
        float Md = -INFINITY;
        idf = 0x7fffffff;
#pragma unroll
        for (int hh = 0; hh < H; ++hh)
          if (t[hh].x > Md) { Md = t[hh].x; idf = __float_as_int(t[hh].z); }
        sd = 0.f;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) sd += t[hh].y * exp_sub(t[hh].x, Md * kFwLog2e);
        md = Md;
      }

      if (valid && h == 0) {
        const float lsed = md + __logf(sd);
        // z[y] was stored a moment ago by a warp of this quarter (ordered by the barrier above / program order)
        const float zy1 = yok ? *(volatile const float*)(p.z[0] + (size_t)row * p.ld_z + y) : 0.f;
        const float zy2 = yok ? *(volatile const float*)(p.z[1] + (size_t)row * p.ld_z + y) : 0.f;
        const float zyd = yok ? zy1 * c1 + zy2 * c2 : 0.f;
        p.conf[row] = c1;
        p.conf[B + row] = c2;
        *reinterpret_cast<float4*>(p.rowstat + (size_t)row * 4) = make_float4(lse1, lse2, lsed, 0.f);
        st[LF_STAT_CE_JOINT] += lsed - zyd;
        st[LF_STAT_CE_X1] += lse1 - zy1;
        st[LF_STAT_CE_X2] += lse2 - zy2;
        st[LF_STAT_SCORE_X1] += __expf(zy1 - lse1);
        st[LF_STAT_SCORE_X2] += __expf(zy2 - lse2);
        st[LF_STAT_CNT_X1] += (i1 == y);
        st[LF_STAT_CNT_X2] += (i2 == y);
        st[LF_STAT_CNT_JOINT] += (ia == y);
        st[LF_STAT_CNT_DF] += (idf == y);
      }
    }

    if (tr && ew == 0 && lane == 0) tr[6] = gtimer();
    // ---- CTA reduction in fixed order -> one partial row per CTA (layout of rows_forward)
#pragma unroll
    for (int i = 0; i < 9; ++i) st[i] = warp_sum(st[i]);
    float* sst = sred + EW * 256;
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < 9; ++i) sst[ew * 16 + i] = st[i];
    named_bar_sync(1, 128 * H);
    const int et = threadIdx.x - 64;                 // 0 .. 128H-1
    float* out = p.partials + (size_t)blockIdx.x * stat_len_dev(C);
    if (et < LF_STATS_HEADER) {
      float s = 0.f;
      if (et < 9)
        for (int w = 0; w < EW; ++w) s += sst[w * 16 + et];
      out[et] = s;
    }
    for (int c = et; c < 2 * C; c += 128 * H) {
      const int m = c >= C ? 1 : 0, cc = c - m * C;
      float s = 0.f;
      for (int w = 0; w < EW; ++w) s += sred[w * 256 + m * 128 + cc];
      out[LF_STATS_HEADER + c] = s;
    }
    for (int r = blockIdx.x + gridDim.x; r < p.nb_total; r += gridDim.x)       // rows no CTA owns
      for (int c = et; c < stat_len_dev(C); c += 128 * H) p.partials[(size_t)r * stat_len_dev(C) + c] = 0.f;
    if (p.sync) {
      // ---- the separate finalize_stats launch, folded in: the last CTA to arrive sums every CTA's partial row, rows in
      // CTA order and in fp64, so the result does not depend on which CTA happens to be last
      double* s_half = reinterpret_cast<double*>(xch);     // [256]: the exchange slots are free by now (H = 4: 4 KB)
      volatile int* s_last = reinterpret_cast<volatile int*>(sstg);
      named_bar_sync(1, 128 * H);                    // this CTA's row is written
      if (et == 0) {
        __threadfence();
        *s_last = atomicAdd(p.sync, 1u) == gridDim.x - 1;
      }
      named_bar_sync(1, 128 * H);
      if (*s_last) {
        __threadfence();
        const int len = stat_len_dev(C), nrow = (int)gridDim.x;
        constexpr int NG = (128 * H >= 512) ? 2 : 1;  // row groups of 256 threads (one column each)
        const int half = NG == 2 ? (nrow + 1) / 2 : nrow;
        const int g = et / 256, c0 = et % 256;
        for (int cb = 0; cb < len; cb += 256) {       // uniform trip count: every thread takes every barrier
          const int c = cb + c0;
          double s = 0.0;
          if (c < len && g < NG) {
            const int r0 = g * half, r1 = min(nrow, r0 + half);
            int r = r0;
            for (; r + 8 <= r1; r += 8) {
              float v[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = __ldcg(p.partials + (size_t)(r + k) * len + c);
#pragma unroll
              for (int k = 0; k < 8; ++k) s += (double)v[k];
            }
            for (; r < r1; ++r) s += (double)__ldcg(p.partials + (size_t)r * len + c);
            if (g == 1) s_half[c0] = s;
          }
          named_bar_sync(1, 128 * H);
          if (c < len && g == 0) p.stats[c] = NG == 2 ? s + s_half[c0] : s;
          named_bar_sync(1, 128 * H);
        }
        if (et == 0) { *p.sync = 0u; __threadfence(); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (tr && threadIdx.x == 0) tr[7] = gtimer();
}

// ---------------------------------------------------------------------------------- host side
bool tc_fwd_supported(int mode, int B, int D, int C, int ld_z, int ld_f, int elem) {
  if (getenv("LF_NO_FUSED_FWD")) return false;
  if (mode != LF_MODE_QMF || C < 32 || C > 128 || B < 1) return false;
  if (ld_z % 4 || ld_f % 4 || ld_z < C || ld_f < C) return false;
  if ((D * elem) % 16) return false;
  return true;
}

constexpr int fw_h(int nch) { return nch >= 4 ? 4 : 2; }          // epilogue warps per TMEM lane quarter
static size_t fw_fixed_smem(int h) {
  return (256 + 4 * h * 256 + 4 * h * 16 + 4 * h * 512) * sizeof(float) + (size_t)2 * 4 * h * 32 * sizeof(float4) + 256;
}

template <int NCH>
static int launch_fwd(const CUtensorMap* mA, const CUtensorMap* mW, const TcFwdParams& p, int grid, size_t smem, cudaStream_t s) {
  constexpr int H = fw_h(NCH);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_fwd_qmf_kernel<NCH, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    attr_set = true;
  }
  LF_LAUNCH("tc_forward_qmf", s, launch_pdl(tc_fwd_qmf_kernel<NCH, H>, dim3(grid), dim3(64 + 128 * H), smem, s, mA[0], mA[1], mW[0], mW[1], p));
  return check_launch("tc_fwd_qmf_kernel");
}

// feat / weight: bf16 (elem 2) or fp32 consumed as TF32 (elem 4); everything else fp32.
int tc_heads_forward_qmf(const void* const feat[2], const void* const weight[2], const float* const bias[2], int elem, int B, int D,
                         int C, float* const z[2], int ld_z, float* avg, float* zdf, int ld_f, float* conf, float* rowstat,
                         const int64_t* label, float* partials, int nb_total, int* grid_out, double* stats, unsigned* sync,
                         cudaStream_t s) {
  TcFwdParams p;
  p.stats = stats; p.sync = stats ? sync : nullptr;
  p.l2_hints = l2_hints_mask() & 3;
  p.B = B; p.C = C; p.D = D; p.ld_z = ld_z; p.ld_f = ld_f;
  p.block_n = div_up(C, 16) * 16;
  p.elem = elem == 2 ? 2 : 4;
  p.kb_elems = 128 / p.elem;
  p.num_kb = div_up(D, p.kb_elems);
  // M tiles of <= 128 samples sized so that they fill whole waves of the 148 persistent CTAs (B = 32768: 293 tiles of
  // 112 samples = 1.98 per CTA instead of 256 tiles = 2 for 108 CTAs and 1 for 40: the kernel is a bandwidth-bound
  // stream and the CTAs that finish early cannot lend their share of the in-flight bytes to the others)
  {
    const int base = div_up(B, TC_BLOCK_M), waves = div_up(base, 148);
    int tm = div_up(div_up(B, waves * 148), 8) * 8;
    if (tm < 64) tm = 64;
    if (tm > TC_BLOCK_M || base <= 148 / 2) tm = TC_BLOCK_M;
    p.tile_m = tm;
  }
  p.m_tiles = div_up(B, p.tile_m);
  p.nb_total = nb_total;
  static unsigned long long* trace_buf = nullptr;
  static int trace_calls = 0;
  p.trace = nullptr;
  if (getenv("LF_FWD_TRACE")) {
    if (!trace_buf) cudaMalloc(&trace_buf, 148 * 8 * sizeof(unsigned long long));
    p.trace = trace_buf;
  }
  for (int m = 0; m < 2; ++m) { p.bias[m] = bias[m]; p.z[m] = z[m]; }
  p.avg = avg; p.zdf = zdf; p.conf = conf; p.rowstat = rowstat; p.label = label; p.partials = partials;
  for (int m = 0; m < 2; ++m)
    if (((uintptr_t)z[m] & 15) || ((uintptr_t)avg & 15) || ((uintptr_t)zdf & 15) || ((uintptr_t)rowstat & 15)) {
      set_error("tc_heads_forward_qmf: outputs must be 16-byte aligned");
      return LF_ERR_BAD_ARG;
    }
  CUtensorMap mA[2], mW[2];
  for (int m = 0; m < 2; ++m) {
    int rc = make_map(&mA[m], feat[m], D, B, D, p.kb_elems, p.tile_m, false, p.elem);
    if (rc) return rc;
    rc = make_map(&mW[m], weight[m], D, C, D, p.kb_elems, p.block_n, false, p.elem);
    if (rc) return rc;
  }
  const uint32_t a_bytes = TC_BLOCK_M * 128, b_bytes = (uint32_t)p.block_n * 128;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  const size_t fixed = fw_fixed_smem(fw_h(p.block_n / 16));
  int stages = 8;
  while (stages > 2 && (size_t)stages * stage_bytes + fixed > 226 * 1024) --stages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  const int grid = p.m_tiles < 148 ? p.m_tiles : 148;
  if (grid_out) *grid_out = grid;
  int rc = LF_ERR_UNSUPPORTED;
  switch (p.block_n / 16) {
    case 2: rc = launch_fwd<2>(mA, mW, p, grid, smem, s); break;
    case 3: rc = launch_fwd<3>(mA, mW, p, grid, smem, s); break;
    case 4: rc = launch_fwd<4>(mA, mW, p, grid, smem, s); break;
    case 5: rc = launch_fwd<5>(mA, mW, p, grid, smem, s); break;
    case 6: rc = launch_fwd<6>(mA, mW, p, grid, smem, s); break;
    case 7: rc = launch_fwd<7>(mA, mW, p, grid, smem, s); break;
    case 8: rc = launch_fwd<8>(mA, mW, p, grid, smem, s); break;
    default:
      set_error("tc_heads_forward_qmf: unsupported class count %d", C);
      return LF_ERR_UNSUPPORTED;
  }
  if (rc == LF_OK && p.trace && ++trace_calls == 12) {
    // LF_FWD_TRACE=1 (eager launches only): per-role %globaltimer stamps of the 12th call, relative to the first CTA's entry
    cudaStreamSynchronize(s);
    static unsigned long long h[148 * 8];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid; ++c) if (h[c * 8] < t0) t0 = h[c * 8];
    const char* nm[8] = {"entry", "setup", "first_load", "last_load", "first_full", "last_commit", "epi_done", "exit"};
    for (int k = 0; k < 8; ++k) {
      double mn = 1e30, mx = 0, sum = 0;
      for (int c = 0; c < grid; ++c) { const double v = (double)(h[c * 8 + k] - t0) / 1000.0; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
      fprintf(stderr, "[fwd trace] %-12s min %7.2f  avg %7.2f  max %7.2f us\n", nm[k], mn, sum / grid, mx);
    }
  }
  return rc;
}

}  // namespace lf
