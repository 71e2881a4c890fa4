// Common device helpers shared by every kernel of the engine.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace fu {

typedef __nv_bfloat16 bf16;

// ---- 4-wide vector access for the two storage types (fp32 parity / bf16 throughput) ----
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
// value as it will read back after being stored in T
__device__ __forceinline__ float rnd(float v, const float*) { return v; }
__device__ __forceinline__ float rnd(float v, const bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- programmatic dependent launch (PDL) ----
// Every kernel of the step calls pdl_wait() before its first global-memory access: when the kernel was launched with
// cudaLaunchAttributeProgrammaticStreamSerialization it may have been scheduled while its predecessor in the stream was
// still running, and this is the point where it waits for that grid (and, transitively, everything before it) to have
// completed and flushed.  Launched normally the instruction returns at once.  pdl_trigger() lets the NEXT kernel of the
// stream be scheduled as soon as every block of this one has started: its prologue (barrier initialisation, TMEM
// allocation, tensor-map prefetch, index arithmetic) then overlaps this kernel's tail instead of following it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- division by a launch-time constant ----
// q = n / d for 0 <= n < 2^31 as multiply-high + add + shift (3 instructions; a 32-bit integer division compiles to
// ~25).  The persistent convolution kernels decode a tile index with up to six divisions per tile in every epilogue
// warp: ~150 of the ~800 warp instructions per tile the thin layers executed (ncu source page, round 2).
// d >= 1; s = ceil(log2 d); m = floor(2^32 (2^s - d) / d) + 1  (Granlund-Montgomery round-up variant).
struct FastDiv {
  uint32_t d, m, s;
  __host__ __device__ FastDiv() : d(1), m(1), s(0) {}
  __host__ explicit FastDiv(int dd) {
    d = dd < 1 ? 1u : (uint32_t)dd;
    s = 0;
    while ((1ull << s) < d) ++s;
    m = (uint32_t)((((1ull << s) - d) << 32) / d) + 1u;
  }
  __device__ __forceinline__ int div(int n) const { return (int)((__umulhi(m, (uint32_t)n) + (uint32_t)n) >> s); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// host-side bump allocator over one device allocation (dry run when base == nullptr)
struct Bump {
  size_t off = 0;
  char* base = nullptr;
  template <typename U>
  U* take(size_t n) {
    off = (off + 255) / 256 * 256;
    U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
    off += n * sizeof(U);
    return p;
  }
};

// kernel launch with or without the PDL attribute (see pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t fu_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                             Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t fu_launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                     unsigned cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 2u : 1u;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// process-wide switch (FU_PDL=0 turns programmatic dependent launch off)
inline bool fu_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* s = getenv("FU_PDL"); v = (s && atoi(s) == 0) ? 0 : 1; }
  return v != 0;
}

#define FU_STR2(x) #x
#define FU_STR(x) FU_STR2(x)

}  // namespace fu
