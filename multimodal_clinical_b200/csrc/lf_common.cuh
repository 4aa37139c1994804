// Shared device/host helpers for the late-fusion step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lf_fusion.h"

namespace lf {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);
int check_launch(const char* what);
// launch accounting (lf_launch_count) and optional per-kernel CUDA-event timing (lf_profile_*)
void prof_begin(const char* name, cudaStream_t s);
void prof_end(const char* name, cudaStream_t s);
#define LF_LAUNCH(name, stream, ...)      \
  do {                                    \
    ::lf::prof_begin(name, stream);       \
    __VA_ARGS__;                          \
    ::lf::prof_end(name, stream);         \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// argmax with torch semantics: first index among equal maxima.
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(kFull, v, o);
    int oi = __shfl_xor_sync(kFull, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl may become resident while its
// predecessor on the stream is still draining, so its prologue (barrier init, TMEM allocation, descriptor prefetch,
// shared-memory setup) overlaps the predecessor's tail.  pdl_wait() blocks until the predecessor has completed and
// its writes are visible: it must precede the first access to global memory.  pdl_trigger() lets the successor start
// its own prologue.  Both are no-ops for a kernel launched the ordinary way.  The attribute is only set when LF_PDL=1 (measured: no
// gain under CUDA-graph replay, where kernel-to-kernel gaps are already ~1 us).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
int l2_hints_mask();
bool l2_hints_enabled();     // L2 eviction-priority hints of the tensor-pipe step (lf_tc_ptx.cuh); LF_NO_L2_HINTS=1 turns them off
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// utils/EMA.py:33  x <- smoothing * x_new + (1 - smoothing) * x, evaluated like torch does (two products, one sum: no
// FMA contraction), so every kernel that applies it -- and the reference -- produce the same bits
__device__ __forceinline__ float ema_mix(float mean, float x, float beta) {
  return __fadd_rn(__fmul_rn(beta, mean), __fmul_rn(__fsub_rn(1.0f, beta), x));
}

// relu that lets NaN through (torch.clamp_min semantics; fmaxf would swallow it)
__device__ __forceinline__ float relu_nan(float x) { return (x > 0.f || x != x) ? x : 0.f; }

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- workspace carving (shared by lf_api.cu and tests of lf_workspace_bytes) -------------------
struct HeadsWorkspace {
  unsigned* sync;        // [LF_WS_SYNC_BYTES / 4] inter-CTA counters, zero between calls: [0] forward tail, [4..5] dW tail
  float* row_partials;   // [kMaxRowBlocks][stat_len]   per-block partial statistics
  float* dw_partials;    // [2][splits][C][D]
  float* db_partials;    // [part_rows][2][C]  column sums of dz per producing CTA
  float* cal_partials;   // [kMaxRowBlocks][2]
  void* w16;             // [2][C][D] bf16 copy of the head weights (LF_PREC_BF16: autocast's per-step cast)
  float* narrow_dw;      // [2][148][C*D] per-CTA dW partials of the fused narrow-head kernel (C <= 32), else null
  size_t total;
};
constexpr int kMaxRowBlocks = 592;   // 4 CTAs per SM on 148 SMs
constexpr int kMaxSplits = 64;

inline int stat_len(int C) { return LF_STATS_HEADER + 2 * C; }
int dw_splits(int B, int D, int C);
HeadsWorkspace carve_heads_workspace(void* base, int B, int D, int C);

}  // namespace lf
