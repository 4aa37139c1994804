// Streaming-read microbenchmarks on B200 (tools only; not part of the library):
//   ldg      : every thread float4 loads, grid-stride, unrolled
//   tma2d    : one producer thread per CTA, [32 x ROWS] SWIZZLE_128B boxes into an S-stage smem ring, consumer
//              warp just waits full and arrives empty (the load path of the tcgen05 kernels without the MMAs)
//   bulk1d   : cp.async.bulk of contiguous CHUNK-byte pieces into the same ring
// Usage: membench <MB> ; prints GB/s per variant.   nvcc -arch=sm_100a -O3 -o membench membench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(ph) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(512) ldg_kernel(const float4* __restrict__ p, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n4; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n4; i += stride) { float4 v = __ldg(p + i); acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456f) out[0] = acc;
}

// rows x cols fp32 matrix; CTA owns rows [r0, r0+rows_per_cta); walks 128-row (ROWS) tiles, k-blocks of 32 cols
__global__ void __launch_bounds__(64) tma2d_kernel(const __grid_constant__ CUtensorMap map, int rows, int cols, int rows_per_cta,
                                                   int box_rows, int stages, int nprod) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t stage_bytes = (uint32_t)box_rows * 128;
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta;
  const int nt = (rows_per_cta + box_rows - 1) / box_rows, nkb = cols / 32;
  const uint32_t n_it = (uint32_t)nt * nkb;
  if (threadIdx.x == 0) {
    for (uint32_t it = 0; it < n_it; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], stage_bytes);
      tma_load_2d(&map, &full[s], smem + (size_t)s * stage_bytes, (it % nkb) * 32, r0 + (it / nkb) * box_rows);
    }
  } else if (threadIdx.x == 32) {
    for (uint32_t it = 0; it < n_it; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
}

__global__ void __launch_bounds__(64) bulk1d_kernel(const uint8_t* __restrict__ base, size_t bytes_per_cta, uint32_t chunk, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * chunk);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* src = base + (size_t)blockIdx.x * bytes_per_cta;
  const uint32_t n_it = (uint32_t)(bytes_per_cta / chunk);
  if (threadIdx.x == 0) {
    for (uint32_t it = 0; it < n_it; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], chunk);
      bulk_load_1d(smem + (size_t)s * chunk, src + (size_t)it * chunk, chunk, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    for (uint32_t it = 0; it < n_it; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F> float time_us(F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(0); f(1); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f(i);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms * 1000.f / iters;
}

int main(int argc, char** argv) {
  const size_t MB = argc > 1 ? atoi(argv[1]) : 96;
  const int nbuf = 4;                               // rotate over > L2 worth of buffers
  const size_t bytes = MB << 20;
  std::vector<float*> bufs(nbuf);
  for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 1, bytes)); }
  float* out; CK(cudaMalloc(&out, 4));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int iters = 20;

  for (int blk : {256, 512}) for (int gm : {2, 4, 8}) {
    float us = time_us([&](int i) { ldg_kernel<<<148 * gm, blk>>>((const float4*)bufs[i % nbuf], bytes / 16, out); }, iters);
    printf("ldg        grid=%4d block=%3d : %7.1f us  %7.1f GB/s\n", 148 * gm, blk, us, bytes / us / 1e3);
  }
  for (int cols : {128, 768}) {
    const int rows = (int)(bytes / (cols * 4));
    for (int box_rows : {64, 128, 256}) for (int stages : {4, 8, 12}) {
      if ((size_t)stages * box_rows * 128 > 200 * 1024) continue;
      std::vector<CUtensorMap> maps(nbuf);
      for (int i = 0; i < nbuf; ++i) {
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, bufs[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      }
      for (int gm : {1, 2}) {
        const int grid = 148 * gm;
        int rpc = (rows + grid - 1) / grid; rpc = (rpc + box_rows - 1) / box_rows * box_rows;
        const size_t smem = (size_t)stages * box_rows * 128 + 256;
        if (gm == 2 && smem > 110 * 1024) continue;
        CK(cudaFuncSetAttribute(tma2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        float us = time_us([&](int i) { tma2d_kernel<<<grid, 64, smem>>>(maps[i % nbuf], rows, cols, rpc, box_rows, stages, 1); }, iters);
        printf("tma2d cols=%4d box_rows=%3d stages=%2d grid=%3d (%3zu KB in flight/SM): %7.1f us  %7.1f GB/s\n", cols, box_rows, stages,
               grid, (size_t)stages * box_rows * 128 * gm / 1024, us, (double)rows * cols * 4 / us / 1e3);
      }
    }
  }
  for (uint32_t chunk : {8192u, 16384u, 32768u}) for (int stages : {4, 8, 12}) {
    if ((size_t)stages * chunk > 200 * 1024) continue;
    for (int gm : {1, 2}) {
      const int grid = 148 * gm;
      const size_t smem = (size_t)stages * chunk + 256;
      if (gm == 2 && smem > 110 * 1024) continue;
      size_t bpc = bytes / grid / chunk * chunk;
      CK(cudaFuncSetAttribute(bulk1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      float us = time_us([&](int i) { bulk1d_kernel<<<grid, 64, smem>>>((const uint8_t*)bufs[i % nbuf], bpc, chunk, stages); }, iters);
      printf("bulk1d chunk=%5u stages=%2d grid=%3d (%3zu KB in flight/SM): %7.1f us  %7.1f GB/s\n", chunk, stages, grid,
             (size_t)stages * chunk * gm / 1024, us, (double)bpc * grid / us / 1e3);
    }
  }
  return 0;
}
