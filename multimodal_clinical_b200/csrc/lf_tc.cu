// Tensor-pipe path for the wide heads (Food101 101-way, VGGSound 309-way): one warp-specialised
// tcgen05 + TMA GEMM kernel (kind::tf32, fp32 accumulators in TMEM) used for all three head products.
//
//   logits   Z  = F  W^T (+b)   M=B  N=Cpad K=D    A = F  K-major   B = W  K-major
//   dfeat    dF = dZ W          M=B  N=D    K=C    A = dZ K-major   B = W  MN-major (no transpose pass)
//   dweight  dW = dZ^T F        M=C  N=D    K=B    A = dZ MN-major  B = F  MN-major, split-K over CTAs
//
// Operands stream HBM -> shared memory with TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight from the
// row-major fp32 tensors — K-major operands as one [32 k x rows] box per stage, MN-major operands as
// [32 mn x 32 k] boxes in the 128B-swizzle/32B-atom mode (the only MN-major layout tf32 accepts).  fp32 data is
// consumed as TF32 by the tensor core (mantissa truncated in the datapath), so there is no conversion
// pass and no extra HBM traffic.  Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// MMA issuer, warps 2..5 = epilogue (tcgen05.ld -> registers -> +bias -> global).  A multi-stage mbarrier
// ring (full/empty) connects producer and issuer; tcgen05.commit releases stages and signals the epilogue.
//
// Out-of-bounds rows/columns are zero-filled by TMA on load and masked on store, so ragged M, N, K
// (C = 101, 309; K = 104, 312) need no padding copies.  Parity class: 2e-2 (tests/test_parity_gpu.py).
#include <cuda.h>
#include "lf_common.cuh"
#include "lf_tc.cuh"

namespace lf {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 32;                 // fp32 elements per stage along K: 128 B = one swizzle span
constexpr int TC_UMMA_K = 8;                   // kind::tf32: 32 B of K per instruction
constexpr int TC_THREADS = 192;

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version field = 1)
// layout_type: 2 = SWIZZLE_128B (16-byte base; K-major operands), 1 = SWIZZLE_128B_BASE32B (the only
// swizzle the tensor core accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                         // c_format = F32
  d |= 2u << 7;                         // a_format = TF32
  d |= 2u << 10;                        // b_format = TF32
  d |= (uint32_t)(a_mn & 1) << 15;      // a_major: 0 = K, 1 = MN
  d |= (uint32_t)(b_mn & 1) << 16;      // b_major
  d |= (uint32_t)(n >> 3) << 17;        // n_dim
  d |= (uint32_t)(m >> 4) << 24;        // m_dim
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------- the kernel
// grid: (m tiles, n tiles, batch * splits).  Dynamic smem: 1024-aligned stages of [A tile | B tile].
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1, TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const CUtensorMap* mapA = batch == 0 ? &mapA0 : &mapA1;
  const CUtensorMap* mapB = batch == 0 ? &mapB0 : &mapB1;
  const int m0 = blockIdx.x * TC_BLOCK_M, n0 = blockIdx.y * p.block_n;
  const int k_begin = split * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int num_kb = (k_end - k_begin + TC_BLOCK_K - 1) / TC_BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(mapA); tma_prefetch_desc(mapB);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
        const int k0 = k_begin + kb * TC_BLOCK_K;
        if (!p.a_mn_major) {
          tma_load_2d(mapA, &full_bar[s], sa, k0, m0);                       // [32 k x 128 rows]
        } else {
#pragma unroll
          for (int i = 0; i < TC_BLOCK_M / 32; ++i)                          // [32 m x 32 k] boxes
            tma_load_2d(mapA, &full_bar[s], sa + i * 4096, m0 + 32 * i, k0);
        }
        if (!p.b_mn_major) {
          tma_load_2d(mapB, &full_bar[s], sb, k0, n0);                       // [32 k x block_n rows]
        } else {
          for (int i = 0; i < p.block_n / 32; ++i)
            tma_load_2d(mapB, &full_bar[s], sb + i * 4096, n0 + 32 * i, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, p.block_n, p.a_mn_major, p.b_mn_major);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        const int krem = k_end - (k_begin + kb * TC_BLOCK_K);
        const int ksteps = krem >= TC_BLOCK_K ? TC_BLOCK_K / TC_UMMA_K : (krem + TC_UMMA_K - 1) / TC_UMMA_K;
        for (int k = 0; k < ksteps; ++k) {
          // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); step 32 B inside the row.
          // MN-major SW128/32B-base: rows are k, 128 B = 32 mn each; the atom is 4 k-rows (512 B, SBO),
          //   32-element MN chunks are 4096 B apart (LBO); one instruction eats 8 k-rows -> step 1024 B.
          const uint64_t da = p.a_mn_major ? make_smem_desc(sa + k * 1024, 4096, 512, 1)
                                           : make_smem_desc(sa + k * 32, 16, 1024, 2);
          const uint64_t db = p.b_mn_major ? make_smem_desc(sb + k * 1024, 4096, 512, 1)
                                           : make_smem_desc(sb + k * 32, 16, 1024, 2);
          umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);               // frees the stage once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);                 // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    float* out = p.out[batch] + (size_t)split * p.split_stride + (size_t)row * p.ld_out;
    const float* bias = p.bias[batch];
    const bool row_ok = row < p.M;
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (num_kb == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      const int col = n0 + c0;
      if (row_ok) {
        if (p.vec_store && col + 16 <= p.N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (bias) { o.x += bias[col + i]; o.y += bias[col + i + 1]; o.z += bias[col + i + 2]; o.w += bias[col + i + 3]; }
            *reinterpret_cast<float4*>(out + col + i) = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < p.N) out[col + i] = v[i] + (bias ? bias[col + i] : 0.f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
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

// 2-D fp32 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch ld elements.
static int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld,
                    int box_inner, int box_outer, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LF_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ld * 4) % 16) { set_error("TMA operand must be 16-byte aligned with a 16-byte row pitch"); return LF_ERR_BAD_ARG; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return LF_ERR_CUDA; }
  return LF_OK;
}

static int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

int tc_gemm(const TcGemmDesc& d, cudaStream_t s) {
  if (d.block_n % 16 || d.block_n < 16 || d.block_n > 256 || (d.b_mn_major && d.block_n % 32)) {
    set_error("tc_gemm: bad block_n %d", d.block_n);
    return LF_ERR_BAD_ARG;
  }
  CUtensorMap mA[2], mB[2];
  for (int b = 0; b < d.nbatch; ++b) {
    int rc;
    // K-major: inner = K, outer = MN rows, box [32 x rows].  MN-major: inner = MN, outer = K rows, box [32 x 32].
    rc = d.a_mn_major ? make_map(&mA[b], d.A[b], d.M, d.K, d.lda, 32, TC_BLOCK_K, true)
                      : make_map(&mA[b], d.A[b], d.K, d.M, d.lda, TC_BLOCK_K, TC_BLOCK_M, false);
    if (rc) return rc;
    rc = d.b_mn_major ? make_map(&mB[b], d.B[b], d.N, d.K, d.ldb, 32, TC_BLOCK_K, true)
                      : make_map(&mB[b], d.B[b], d.K, d.N, d.ldb, TC_BLOCK_K, d.block_n, false);
    if (rc) return rc;
  }
  if (d.nbatch == 1) { mA[1] = mA[0]; mB[1] = mB[0]; }
  TcGemmParams p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.block_n = d.block_n;
  p.splits = d.splits < 1 ? 1 : d.splits;
  int kps = div_up(d.K, p.splits);
  kps = div_up(kps, TC_BLOCK_K) * TC_BLOCK_K;
  p.k_per_split = kps;
  p.a_mn_major = d.a_mn_major; p.b_mn_major = d.b_mn_major;
  for (int b = 0; b < 2; ++b) { p.out[b] = d.out[b < d.nbatch ? b : 0]; p.bias[b] = d.bias[b < d.nbatch ? b : 0]; }
  p.ld_out = d.ld_out; p.split_stride = d.split_stride;
  p.vec_store = (d.ld_out % 4 == 0) && (((uintptr_t)d.out[0] & 15) == 0) && (d.nbatch < 2 || ((uintptr_t)d.out[1] & 15) == 0) &&
                (d.split_stride % 4 == 0);
  p.tmem_cols = pow2_cols(d.block_n);
  const int num_kb = div_up(kps, TC_BLOCK_K);
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t b_bytes = (uint32_t)d.block_n * TC_BLOCK_K * 4;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  int stages = num_kb < 4 ? (num_kb < 1 ? 1 : num_kb) : 4;
  while (stages > 1 && (size_t)stages * stage_bytes + 1024 + 256 > 220 * 1024) --stages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    attr_set = true;
  }
  dim3 grid(div_up(d.M, TC_BLOCK_M), div_up(d.N, d.block_n), d.nbatch * p.splits);
  LF_LAUNCH(d.name, s, (tc_gemm_kernel<<<grid, TC_THREADS, smem, s>>>(mA[0], mB[0], mA[1], mB[1], p)));
  return check_launch(d.name);
}

}  // namespace lf

// Test hook (no reference counterpart): one tensor-pipe GEMM through the C ABI so the kernel can be
// checked in isolation.  out(M,N) = A*B (+bias) with the operand orientations described in lf_tc.cuh.
extern "C" int lf_debug_tc_gemm(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                                int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                                int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride,
                                void* stream) {
  lf::TcGemmDesc d;
  d.nbatch = 1;
  d.A[0] = A; d.B[0] = B; d.bias[0] = bias; d.out[0] = out;
  d.A[1] = A; d.B[1] = B; d.bias[1] = bias; d.out[1] = out;
  d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ld_out = ld_out;
  d.a_mn_major = a_mn_major; d.b_mn_major = b_mn_major; d.block_n = block_n;
  d.splits = splits; d.split_stride = split_stride; d.name = "tc_gemm_debug";
  return lf::tc_gemm(d, (cudaStream_t)stream);
}
