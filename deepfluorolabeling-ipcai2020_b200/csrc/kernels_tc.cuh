// tcgen05 / TMA tensor-core kernels of the engine (throughput mode, bf16 NHWC, sm_100a).
//
// Convolution = implicit GEMM without im2col:
//   D[128 pixels, BN out-channels] += A_tap[128 pixels, KC in-channels] * W_tap[BN, KC]^T
// for every filter tap and KC-channel chunk.  The A operand of a tap is ONE 4-D TMA box
// (KC, tw, th, tn) of the NHWC activation tensor at coordinates shifted by the tap offset;
// out-of-bounds rows/columns are zero-filled by TMA, which is exactly the zero padding of
// nn.Conv2d(padding=1) (unet.py:211-212).  Both operands land in shared memory in the
// canonical K-major swizzled layout tcgen05.mma reads; accumulators live in TMEM (double
// buffered), the epilogue (bias, ReLU, BN-apply + residual, BN statistics) runs from TMEM
// through registers into a swizzled staging tile that TMA stores back to NHWC global memory.
// A channel slice of a wider buffer (skip concat, unet.py:257) is just a tensor map with a
// larger pixel stride: no copy.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fluoro_unet.h"
#include "common.cuh"

#define FU_TC_BUILD "1"
// packed fp32x2 epilogue arithmetic (FADD2 / FFMA2): measured on the B200 and left off -- the thin 3x3 layers do not
// change (32->32 @192x192 forward 63.6 vs 64.0 us) and the 1x1 layers with a second epilogue operand get slower
// (64->32 @192x192: 98.7 -> 113 us), see gpurun A/B in DESIGN.md section 5
#ifndef FU_TC_TIMELINE
#define FU_TC_TIMELINE 0
#endif
#ifndef FU_EPI_PACKED
#define FU_EPI_PACKED 0
#endif

namespace fu {

// packed fp32 pairs (FADD2 / FFMA2, sm_100): one issue slot for two lanes of an elementwise epilogue op
__device__ __forceinline__ float2 fu_add2(float2 a, float2 b) {
#if FU_EPI_PACKED
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 fu_fma2(float2 a, float2 b, float2 c) {
#if FU_EPI_PACKED
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// ===========================================================================
// PTX wrappers
// ===========================================================================
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.  The report lives out of
// line: inlined at every wait site it put ~12 instructions of printf set-up into each hot loop (instruction-cache
// footprint is what limits the single MMA-issuing warp, see tc_mma_taps_resident).
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("fluoro_unet: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) mbar_timeout(bar, parity);
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// multicast variants: the box lands at the same shared-memory offset (and signals the mbarrier at the same offset) in
// every CTA of the cluster whose bit is set in `mask`
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// shared -> global bulk reduction (fp32 add) of `bytes` contiguous bytes (16-byte aligned, multiple of 16)
__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with both descriptors given as (shared high word, 32-bit low word).  The start-address field of a
// shared-memory descriptor is bits 0-13 of the low word (16-byte units, < 256 KB: it never carries), so all per-tap /
// per-K-step descriptor arithmetic is ONE 32-bit add instead of a 64-bit add + uniform-register shuffling: the thin
// layers (N = 32/64: 16-32 tensor cycles per MMA) were bound by the ~20 SASS instructions the issuing thread spent
// per MMA (ncu source page, profiles/r02_ncu_thin_conv_before.txt), not by the tensor pipe.
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same arrival on the mbarrier at this offset in every CTA of the cluster selected by `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// true in exactly one lane of a fully converged warp.  Code guarded by it is known by the compiler to run
// in a single thread, so tcgen05.mma / cp.async.bulk.tensor (which take uniform registers) are emitted as
// straight-line code instead of a per-active-lane ELECT loop (measured: ~30 extra instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ptx

// shared-memory matrix descriptor of a K-major operand tile whose rows are `row_bytes` wide
// (row_bytes = swizzle span: 32/64/128 B), 8-row groups `8*row_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);  // SWIZZLE_128B / 64B / 32B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, 16-byte units
  d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3FFF) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= layout << 61;
  return d;
}
// descriptor of the same tile family at another shared-memory address (< 256 KB, so no carry out of the field)
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t base_desc_no_addr, uint32_t smem_addr) {
  return base_desc_no_addr + (uint64_t)(smem_addr >> 4);
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// shared-memory matrix descriptor of an MN-major operand: 64-element (128 B, SWIZZLE_128B) blocks along
// M/N that are `lbo` bytes apart, K rows 128 B apart, 8-row K groups `sbo` = 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// bf16 x bf16 -> fp32, both operands MN-major (K = pixels is the strided dimension), M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(uint32_t n, uint32_t m = 128u) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// KS K-steps of one filter tap into the accumulators of a pixel-tile pair (32 bytes = +2 in the address field per step)
template <int KS>
__device__ __forceinline__ void tc_mma_tap(uint32_t d0, uint32_t d1, bool two, uint32_t a_lo, uint32_t a2_lo, uint32_t b_lo,
                                           uint32_t hi, uint32_t idesc, uint32_t accum) {
#pragma unroll
  for (int j = 0; j < KS; ++j) {
    ptx::umma_bf16_lo(d0, a_lo + 2u * j, b_lo + 2u * j, hi, idesc, j == 0 ? accum : 1u);
    if (two) ptx::umma_bf16_lo(d1, a2_lo + 2u * j, b_lo + 2u * j, hi, idesc, j == 0 ? accum : 1u);
  }
}
// All taps of one halo tile against the resident weights of its channel chunk.  Deliberately ROLLED loops: the
// issuing warp shares its instruction caches with the epilogue warps, and the fully unrolled version of this
// sequence (9 taps x KS steps x 2 tiles, ~5-10 KB of SASS per super tile) was evicted between super tiles -- 48 % of
// the issuing warp's stall samples were "no instruction" (ncu source page, profiles/r02_ncu_thin_conv_before.txt).
// g < 0: all nine taps; otherwise the three taps with kw == g (one-load-per-kw mode).  Offsets in 16-byte units.
// baton != 0: mbarrier to arrive on when the last filter row starts (hands the tensor pipe to the other issuer, below).
__device__ __forceinline__ void tc_mma_taps_resident(int ksteps, uint32_t d0, uint32_t d1, bool two, uint32_t a_lo0,
                                                     uint32_t a_tile16, uint32_t b_lo0, uint32_t b16, uint32_t kh16,
                                                     uint32_t kw16, int g, uint32_t hi, uint32_t idesc, uint32_t accum,
                                                     uint32_t baton = 0u, int baton_kh = 2) {
  uint32_t a_row = a_lo0;
  uint32_t b_lo = g < 0 ? b_lo0 : b_lo0 + (uint32_t)g * b16;
  const int nkw = g < 0 ? 3 : 1;
  const uint32_t b_step = g < 0 ? b16 : 3u * b16;
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh, a_row += kh16) {
    uint32_t a_lo = a_row;
    if (kh == baton_kh && baton) ptx::mbar_arrive(baton);
#pragma unroll 1
    for (int kw = 0; kw < nkw; ++kw, a_lo += kw16, b_lo += b_step) {
      if (ksteps == 4) tc_mma_tap<4>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, hi, idesc, accum);
      else if (ksteps == 2) tc_mma_tap<2>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, hi, idesc, accum);
      else tc_mma_tap<1>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, hi, idesc, accum);
      accum = 1u;
    }
  }
  if (baton_kh >= 3 && baton) ptx::mbar_arrive(baton);
}

// ===========================================================================
// the convolution kernel (forward and data gradient; 3x3/pad 1 and 1x1)
// ===========================================================================
// BatchNorm finalisation folded into the prologue of the kernel that applies it (the residual 1x1 conv, whose epilogue
// computes BN2(r2) + conv1x1(x)): every CTA turns the fp64 batch sums into the folded scale / shift it needs, CTA 0
// also stores mean / invstd for the backward pass and updates the running statistics (nn.BatchNorm2d, unet.py:222).
struct TcBnFin {
  const double* stat;           // [2C] sum, sum of squares (null: not used, bn_a / bn_b are given)
  long long P;
  int training;
  const float* gamma; const float* beta;
  float* rmean; float* rvar; long long* nbt;
  float momentum, eps;
  float* mean_o; float* invstd_o;
};

struct TcConvParams {
  int B, H, W;
  int Cin, N;
  int ksz, pad;
  int KC, BN, CS;               // channels per k-chunk, N tile, channels per store box
  int tw, th, tn;               // pixel tile (tw*th*tn <= 128)
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int stages;
  int relu;
  const float* bias;            // [N] or null
  const bf16* t; int t_ld;      // optional second operand of the epilogue, NHWC with pixel stride t_ld
  const float* bn_a; const float* bn_b;   // out += bn_a[c]*t + bn_b[c]   (null: out += t)
  double* stat;                 // optional [2N]: per-channel sum / sum of squares of the stored values
  // 2x2/stride-2 variants (Conv2d(C,C,2,2) and ConvTranspose2d(.,.,2,2), unet.py:93,240): the pixel grid
  // is (W, H = B*rows) with B = 1 (image rows merged; no padding so tiles may span images)
  int nstaging;                 // staging tiles for the TMA store (2 = double buffered)
  int a5;                       // A operand gathered with stride 2: 5-D map (C,2,W,2,H), tap = (kh,kw)
  int c5;                       // output scattered (pixel shuffle): 5-D map (Cst,2,W,2,H), GEMM column = (a,b,co)
  TcBnFin fin;                  // fin.stat != null: bn_a / bn_b are computed here instead of being read
  int Cst;                      // channels of the scattered output tensor (N = 4*Cst)
  int cst_shift;                // log2(Cst): channel counts are powers of two (unet.py:86: 2**(wf+i)), so the per-chunk
                                // (a,b) block / channel split is a shift and a mask
  // split-bf16 x3 parity mode (kernel template F32 = true): fp32 values travel as bf16 pairs hi = bf16(v),
  // lo = bf16(v - hi); the K loop runs three passes per channel chunk, (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi),
  // into the same fp32 accumulator (the dropped lo*lo term is ~2^-16 relative).  The A map spans the operand's
  // split twin, whose lo half sits `a_lo` channels after the hi half; the weight map's K dimension is [hi | lo]
  // with lo at `b_lo`.  Output and second epilogue operand are fp32.
  int a_lo, b_lo;
  const float* tf;              // (F32) second epilogue operand, fp32 NHWC with pixel stride t_ld
  int post;                     // 1: bn_a / bn_b are applied AFTER the ReLU (out = bn_a * relu(acc + bias) + bn_b, no `t`):
                                // eval-mode BatchNorm folded into the producing convolution
  int dual;                     // two MMA-issuing warps on alternate tiles (stages >= 2 x the K iterations of a tile, so
                                // that an issuer one tile ahead is never a whole ring round ahead; see TcConv3Params)
  FastDiv fd_ntiles, fd_tw, fd_th;     // tile decode: divisions by n_tiles, tiles_w, tiles_h
  int cs_shift;                 // log2(CS)
  int alias_staging;            // 1 (every CTA has exactly one tile): the epilogue's staging tiles overlay the operand stages,
                                // which are dead once the tile's last MMA has completed, so the whole shared memory holds
                                // operands in flight.  The deep layers (12x12, 24x24: 144 tiles) are bound by the bytes in
                                // flight per SM: 32-48 KB per K iteration against 256-512 tensor cycles needs more than the
                                // 3-4 stages left beside 64 KB of staging (ncu: the MMA warp polls its full barrier ~2x per
                                // iteration, tensor pipe 36 % busy, profiles/r02_ncu_deep_conv_512_before.txt)
  int kpair;                    // 1: a pipeline stage holds TWO 64-channel K chunks of both operands, each pair fetched by one
                                // TMA instruction (A through a 5-D view whose slowest box dimension is the chunk index, so
                                // the two [pixels][64] tiles land back to back; B likewise in 4-D): half as many barrier
                                // round trips, TMA issues and elected sections per FLOP.  The deep layers run at the
                                // producer / issuer warps' loop rate (~600-700 cycles per 4-MMA iteration against 256
                                // tensor cycles; ncu: the producer never waits for a free stage), not at a memory limit
  int mc;                       // 0: no cluster.  1 / 2: launched as clusters of two CTAs (one tile each) that share their A
                                // tile (1: same pixels, neighbouring N tiles) or their B tile (2: same N tile, neighbouring
                                // pixel tiles).  The shared operand of K iteration k is loaded by CTA k % 2 and MULTICAST into
                                // both, so each CTA pulls 25-33 % fewer bytes through the L2: the deep layers move
                                // 144 CTAs x 32-48 KB per iteration = 10-12 TB/s, which is the L2's limit, not the tensor
                                // pipe's (36 % busy).  Stage release is cluster-wide: every MMA warp commits to the empty
                                // barrier of BOTH CTAs (count 2).
  int t_smem;                   // 1: the second epilogue operand `t` is brought into shared memory by the TMA producer
                                // (map tmT, two tiles per epilogue group, laid out like the staging tile) instead of being
                                // loaded from global memory by the epilogue threads: with 1-8 MMAs per tile the accumulator
                                // is ready at once and the threads' own loads put a full memory latency into every tile
                                // (ncu, residual 1x1 64->32 @192x192: 59 % of the stall samples long-scoreboard, 3.8 TB/s)
};
constexpr int kTcTBufs = 2;           // `t` tiles per epilogue group

constexpr int kTcThreads = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue (wgrad kernels; the conv kernel has 4*G epilogue warps)
constexpr int kTcMaxStages = 8;
constexpr int kTc3Threads = 320;      // halo kernel: producer, MMA, 2 x 4 epilogue warps (one group per paired pixel tile)

// G = epilogue groups of 4 warps.  Tile i of a CTA (in its own order) is drained by group i % G from TMEM
// stage i % (2G), so up to 2G accumulators are in flight.  The 1x1 / 2x2 layers have 4-16 MMAs per tile and
// are bound by the epilogue's per-warp latency chain (TMEM load, second-operand fetch, convert, staging
// store, TMA store, statistics): four groups cut their time ~3x; the deep 3x3 layers keep G = 1 or 2.
template <int G, bool F32 = false>
__global__ void __launch_bounds__(96 + 128 * G, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmT, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw);

  // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so role-local scalars can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // Warp roles.  The epilogue warps come FIRST and the TMA producer / MMA issuer LAST: the SM sub-partition schedulers
  // prefer the highest warp id among eligible warps, and the single MMA-issuing thread is the critical path of the
  // thin layers -- as warp 1 it was starved by the (instruction-heavy) epilogue warps sharing its scheduler (ncu:
  // ~1000 idle tensor-pipe cycles between tiles, profiles/r02_ncu_thin_conv_before.txt).
  constexpr int prod_warp = 4 * G, mma_warp = 4 * G + 1;
  const uint32_t row_bytes = (uint32_t)p.KC * 2u;
  const uint32_t a_bytes = 128u * row_bytes, b_bytes = (uint32_t)p.BN * row_bytes;
  const uint32_t kp = p.kpair ? 2u : 1u;                   // K chunks per stage: [A0 A1 | B0 B1]
  const uint32_t stage_bytes = kp * (a_bytes + b_bytes);
  const uint32_t staging_off = p.alias_staging ? 0u : (uint32_t)p.stages * stage_bytes;
  // F32: one 128-row x 32-column fp32 chunk (16 KB) per staging tile, always double buffered
  const uint32_t staging_bytes = F32 ? 16384u : 128u * (uint32_t)p.BN * 2u;
  const int vlen = p.c5 ? p.Cst : p.N;                              // channels of the bias / bn vectors
  const int nstg = F32 ? 2 : p.nstaging;                             // 1 or 2 staging tiles (double buffered stores)
  const uint32_t tbuf_off = (uint32_t)p.stages * stage_bytes +
                            (p.alias_staging ? 0u : (uint32_t)(G * nstg) * staging_bytes);  // `t` tiles [G][kTcTBufs] (t_smem)
  const uint32_t vec_off = tbuf_off + (p.t_smem ? (uint32_t)(G * kTcTBufs) * staging_bytes : 0u);   // bias | bn_a | bn_b, [3][vlen] floats
  const uint32_t park_off = vec_off + 3u * (uint32_t)vlen * 4u;     // (F32) statistics partials [G][4][2][BN] floats
  const uint32_t bar_off = (park_off + (F32 ? (uint32_t)(G * 8 * p.BN) * 4u : 0u) + 7u) & ~7u;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kTcMaxStages + s); };
  constexpr int NACC = 2 * G;             // TMEM accumulator stages
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kTcMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kTcMaxStages + NACC + a); };
  const uint32_t slot_addr = bar_base + 8u * (uint32_t)(2 * kTcMaxStages + 2 * NACC);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * kTcMaxStages + 2 * NACC));
  // `t` tile barriers: index g * kTcTBufs + b
  auto tt_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * kTcMaxStages + 2 * NACC + 1 + i); };
  auto tt_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * kTcMaxStages + 2 * NACC + 1 + G * kTcTBufs + i); };

  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(NACC * p.BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), p.mc ? 2u : 1u); }
    for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 128); }
    for (int i = 0; i < G * kTcTBufs; ++i) { ptx::mbar_init(tt_full(i), 1); ptx::mbar_init(tt_empty(i), 128); }
    if (p.t_smem) ptx::prefetch_tmap(&tmT);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmB); ptx::prefetch_tmap(&tmC);
  }
  if (warp == mma_warp) {
    ptx::tmem_alloc(slot_addr, tmem_cols);
    ptx::tmem_relinquish();
  }
  // everything above touched only this CTA's shared memory / TMEM and the kernel parameters: under programmatic
  // dependent launch it ran while the previous kernel of the stream was finishing.  From here on its results are read.
  pdl_wait(); pdl_trigger();
  {
    // per-channel epilogue vectors in shared memory (no global load in the epilogue's dependency chain)
    float* vec = reinterpret_cast<float*>(smem + vec_off);
    if (p.fin.stat) {
      const TcBnFin& f = p.fin;
      if (blockIdx.x == 0 && threadIdx.x == 0 && f.training && f.nbt) *f.nbt += 1;
      for (int i = threadIdx.x; i < vlen; i += blockDim.x) {
        float mean, var;
        double unb = 0.0;
        if (f.training) {
          const double m = f.stat[i] / (double)f.P;
          double v = f.stat[vlen + i] / (double)f.P - m * m;
          if (v < 0.0) v = 0.0;
          mean = (float)m; var = (float)v;
          unb = f.P > 1 ? v * ((double)f.P / (double)(f.P - 1)) : v;
        } else {
          mean = f.rmean[i]; var = f.rvar[i];
        }
        const float invstd = rsqrtf(var + f.eps);
        const float a = f.gamma[i] * invstd;
        vec[i] = p.bias ? p.bias[i] : 0.f;
        vec[vlen + i] = a;
        vec[2 * vlen + i] = f.beta[i] - mean * a;
        if (blockIdx.x == 0) {
          f.mean_o[i] = mean; f.invstd_o[i] = invstd;
          if (f.training) {      // (train mode never reads the running statistics, eval mode never writes them)
            f.rmean[i] = (1.f - f.momentum) * f.rmean[i] + f.momentum * mean;
            f.rvar[i] = (1.f - f.momentum) * f.rvar[i] + f.momentum * (float)unb;
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < vlen; i += blockDim.x) {
        vec[i] = p.bias ? p.bias[i] : 0.f;
        vec[vlen + i] = p.bn_a ? p.bn_a[i] : 1.f;
        vec[2 * vlen + i] = p.bn_a ? p.bn_b[i] : 0.f;
      }
    }
    if (F32) {
      float* park = reinterpret_cast<float*>(smem + park_off);
      for (int i = threadIdx.x; i < G * 8 * p.BN; i += blockDim.x) park[i] = 0.f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (p.mc) ptx::cluster_sync();          // the peer's barriers are initialised before anything is multicast into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  const uint32_t crank = p.mc ? ptx::cluster_ctarank() : 0u;

  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
  const int total_tiles = tiles_m * p.n_tiles;
  const int cchunks = p.Cin / p.KC;
  const int vchunks = F32 ? 3 * cchunks : cchunks;          // split mode: three passes per real channel chunk
  const int k_iters = p.ksz * p.ksz * vchunks / (int)kp;    // pipeline stages per tile
  const uint32_t a_tx = (uint32_t)(p.tw * p.th * p.tn) * row_bytes;

  auto decode = [&](int tile, int& w0, int& h0, int& n0, int& nb) {
    int nt, mt, r;
    p.fd_ntiles.divmod(tile, mt, nt);
    nb = nt * p.BN;
    p.fd_tw.divmod(mt, mt, r); w0 = r * p.tw;
    p.fd_th.divmod(mt, mt, r); h0 = r * p.th;
    n0 = mt * p.tn;
  };

  if (warp == prod_warp) {
    // ------------------------------ TMA producer (whole warp loops, one elected lane issues) -----
    {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      const uint32_t t_sub_bytes = 128u * (uint32_t)p.CS * 2u;
      const uint32_t t_tx = (uint32_t)(p.tw * p.th * p.tn) * (uint32_t)p.BN * 2u;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int w0, h0, n0, nb;
        decode(tile, w0, h0, n0, nb);
        if (p.t_smem) {
          // the tile's second epilogue operand, for the group that will drain it (tile i of the CTA -> group i % G,
          // buffer (i / G) % kTcTBufs of that group)
          const int g = it % G, u = it / G, b = u % kTcTBufs, bi = g * kTcTBufs + b;
          ptx::mbar_wait(tt_empty(bi), (((uint32_t)(u / kTcTBufs)) & 1u) ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t dst = smem_base + tbuf_off + (uint32_t)bi * staging_bytes;
            ptx::mbar_expect_tx(tt_full(bi), t_tx);
            for (int s2 = 0; s2 < p.BN / p.CS; ++s2) {
              if (p.c5) {
                const int col = nb + s2 * p.CS, ab = col >> p.cst_shift;
                ptx::tma_load_5d(dst + (uint32_t)s2 * t_sub_bytes, &tmT, tt_full(bi), col & (p.Cst - 1), ab & 1, w0, ab >> 1, h0);
              } else {
                ptx::tma_load_4d(dst + (uint32_t)s2 * t_sub_bytes, &tmT, tt_full(bi), nb + s2 * p.CS, w0, h0, n0);
              }
            }
          }
          __syncwarp();
        }
        // (tap, channel chunk) advanced incrementally: this loop bounds the deep layers -- the MMA warp polls its full
        //  barrier ~2x per iteration (ncu, 512->512 @12x12) -- and two integer divisions were half of its instructions
        int tap = 0, vc = 0, kh = 0, kw = 0;
        for (int kit = 0; kit < k_iters; ++kit) {
          int c0 = vc * p.KC, cb0 = c0;
          if (F32) {                       // pass 0: (hi, W_hi), pass 1: (hi, W_lo), pass 2: (lo, W_hi)
            const int pass = vc / cchunks;
            c0 = (vc - pass * cchunks) * p.KC;
            cb0 = c0 + (pass == 1 ? p.b_lo : 0);
            c0 += pass == 2 ? p.a_lo : 0;
          }
          const int tap_c = tap, kh_c = kh, kw_c = kw, vc_c = vc;
          vc += (int)kp;
          if (vc == vchunks) { vc = 0; ++tap; if (++kw == p.ksz) { kw = 0; ++kh; } }
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t a_dst = smem_base + (uint32_t)stage * stage_bytes;
            ptx::mbar_expect_tx(full_bar(stage), kp * (a_tx + b_bytes));
            const bool mine = ((uint32_t)kit & 1u) == crank;       // (cluster) this CTA loads the shared operand of this iteration
            if (p.kpair) {
              // chunks vc_c, vc_c + 1 of both operands: A tiles back to back (dense: a_tx bytes each), B tiles behind them
              ptx::tma_load_5d(a_dst, &tmA, full_bar(stage), 0, w0 + kw_c - p.pad, h0 + kh_c - p.pad, n0, vc_c);
              ptx::tma_load_4d(a_dst + 2u * a_bytes, &tmB, full_bar(stage), 0, tap_c, nb, vc_c);
            } else
            if (p.a5) ptx::tma_load_5d(a_dst, &tmA, full_bar(stage), c0, kw_c, w0, kh_c, h0);
            else if (p.mc == 1) { if (mine) ptx::tma_load_4d_mc(a_dst, &tmA, full_bar(stage), c0, w0 + kw_c - p.pad, h0 + kh_c - p.pad, n0, (uint16_t)3); }
            else ptx::tma_load_4d(a_dst, &tmA, full_bar(stage), c0, w0 + kw_c - p.pad, h0 + kh_c - p.pad, n0);
            if (p.kpair) { }
            else if (p.mc == 2) { if (mine) ptx::tma_load_3d_mc(a_dst + a_bytes, &tmB, full_bar(stage), cb0, tap_c, nb, (uint16_t)3); }
            else ptx::tma_load_3d(a_dst + a_bytes, &tmB, full_bar(stage), cb0, tap_c, nb);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= mma_warp) {
    // ------------------------------ MMA issuers (whole warp loops, one elected lane issues) -----
    // Two issuers on alternate tiles when p.dual (see tc_conv3_kernel): the 1x1 / 2x2 layers have 1-8 MMAs per tile,
    // so a single issuer spends most of a tile in barrier round trips.
    {
      const int mw = warp - mma_warp;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int it = 0;
      const uint32_t idesc = umma_idesc_bf16((uint32_t)p.BN);
      const uint64_t dbase = umma_desc_kmajor(0, row_bytes);
      const uint32_t dhi = (uint32_t)(dbase >> 32), dlo = (uint32_t)dbase;
      const int ksteps = p.KC / 16;
      const uint32_t a16 = a_bytes >> 4;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        if (!p.dual) { if (mw) break; }
        else if ((it & 1) != mw) {                  // the other issuer's tile: step over its ring slots
          stage += k_iters; while (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
          if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int kit = 0; kit < k_iters; ++kit) {
          // (no tcgen05.fence here: the operands were written by TMA, whose completion the mbarrier itself orders
          //  before this thread's MMAs; the fence is only needed after waits on barriers that OTHER THREADS'
          //  tcgen05.ld signalled, i.e. the accumulator hand-back above)
          ptx::mbar_wait(full_bar(stage), phase);
          if (ptx::elect_one()) {
            // +32 B per K step = +2 in the address field of the descriptors' low words (see umma_bf16_lo)
            const uint32_t a_lo = dlo + ((smem_base + (uint32_t)stage * stage_bytes) >> 4);
            const uint32_t first = kit != 0 ? 1u : 0u;
            if (p.kpair) {          // (KC = 64) two chunks: A tiles a_tx bytes apart, B tiles behind both A slots
              tc_mma_tap<4>(d_tmem, 0u, false, a_lo, 0u, a_lo + 2u * a16, dhi, idesc, first);
              tc_mma_tap<4>(d_tmem, 0u, false, a_lo + (a_tx >> 4), 0u, a_lo + 2u * a16 + (b_bytes >> 4), dhi, idesc, 1u);
            } else
            if (ksteps == 4) tc_mma_tap<4>(d_tmem, 0u, false, a_lo, 0u, a_lo + a16, dhi, idesc, first);
            else if (ksteps == 2) tc_mma_tap<2>(d_tmem, 0u, false, a_lo, 0u, a_lo + a16, dhi, idesc, first);
            else tc_mma_tap<1>(d_tmem, 0u, false, a_lo, 0u, a_lo + a16, dhi, idesc, first);
            if (p.mc) ptx::umma_commit_mc(empty_bar(stage), (uint16_t)3);   // (cluster) the stage is refilled for both CTAs at once
            else ptx::umma_commit(empty_bar(stage));     // frees the smem stage when these MMAs retire
            if (kit == k_iters - 1) ptx::umma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------ epilogue (G groups of 4 warps / 128 threads) ------------------------------
    const int grp = warp >> 2;        // which group; it drains tiles grp, grp + G, ... of this CTA
    const int bar_id = 1 + grp;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // accumulator row = pixel slot of the tile
    const int et = threadIdx.x & 127;   // 0..127 within the group
    const uint32_t pitch = (uint32_t)p.CS * 2u;
    const uint32_t swz_mask = pitch == 128 ? 7u : (pitch == 64 ? 3u : 1u);
    const uint32_t sub_bytes = 128u * pitch;
    const uint32_t grp_staging_off = staging_off + (uint32_t)(grp * nstg) * staging_bytes;
    uint8_t* staging = smem + grp_staging_off;
    uint32_t staging_addr = smem_base + grp_staging_off;
    int sbuf = 0;
    int acc = grp; uint32_t acc_phase = 0;
    int tb = 0; uint32_t tph = 0;           // `t` tile ring of this group (t_smem)
    const int rows_in_tile = p.tw * p.th * p.tn;
    const int wi = row % p.tw, hi = (row / p.tw) % p.th, ni = row / (p.tw * p.th);
    const float* vec = reinterpret_cast<const float*>(smem + vec_off);
    // statistics ownership: column pair cp, row group rg (BN/2 rows each); deterministic flush (see tc_conv3_kernel)
    const int cpairs = p.BN >> 1;
    const int cp = et % cpairs, rg = et / cpairs;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    int s_nb = -1;
    auto flush_stats = [&]() {
      if (p.stat && s_nb >= 0) {
        float* park = reinterpret_cast<float*>(staging);
        const int rgroups = 128 / cpairs;
        if (et == 0) ptx::tma_store_wait_read();     // no store may still be reading the tile we park in
        ptx::named_bar_sync(bar_id, 128);
        park[(rg * 2 + 0) * p.BN + 2 * cp] = s0; park[(rg * 2 + 0) * p.BN + 2 * cp + 1] = s1;
        park[(rg * 2 + 1) * p.BN + 2 * cp] = q0; park[(rg * 2 + 1) * p.BN + 2 * cp + 1] = q1;
        s0 = s1 = q0 = q1 = 0.f;
        ptx::named_bar_sync(bar_id, 128);
        for (int i = et; i < 2 * p.BN; i += 128) {
          const int which = i / p.BN, c = i - which * p.BN;
          float v = 0.f;
          for (int r = 0; r < rgroups; ++r) v += park[(r * 2 + which) * p.BN + c];
          atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
        }
        ptx::named_bar_sync(bar_id, 128);
      }
    };
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < total_tiles; tile += G * (int)gridDim.x) {
      int w0, h0, n0, nb;
      decode(tile, w0, h0, n0, nb);
      const bool valid = row < rows_in_tile && (w0 + wi) < p.W && (h0 + hi) < p.H && (n0 + ni) < p.B;
      long long pix = ((long long)(n0 + ni) * p.H + (h0 + hi)) * p.W + (w0 + wi);
      // Scattered (pixel-shuffle) output: GEMM column = (a,b,co).  An N tile may span several (a,b) blocks (BN a
      // multiple of Cst: the A tile is then loaded once for all four sub-pixels), so the channel base and the
      // output pixel are functions of the 32-column chunk, not of the tile.
      const long long pix_c5 = (long long)(h0 + hi) * 2 * (2 * p.W) + 2 * (w0 + wi);   // sub-pixel (0,0) of this row's pixel
      auto chunk_cb = [&](int col) { return p.c5 ? (col & (p.Cst - 1)) : col; };     // channel base in bias / t / output
      auto chunk_pix = [&](int col) {
        if (!p.c5) return pix;
        const int ab = col >> p.cst_shift;
        return pix_c5 + (long long)(ab >> 1) * (2 * p.W) + (ab & 1);
      };
      if constexpr (F32) {
        // ---- fp32 output (parity mode): 32-column chunks through a double-buffered 16 KB staging tile ----
        float* park = reinterpret_cast<float*>(smem + park_off) + grp * 8 * p.BN;     // [4 row groups][2][BN]
        if (p.stat && nb != s_nb) {
          if (s_nb >= 0) {
            ptx::named_bar_sync(bar_id, 128);
            for (int i = et; i < 2 * p.BN; i += 128) {
              const int which = i / p.BN, c = i - which * p.BN;
              float v = 0.f;
#pragma unroll
              for (int r = 0; r < 4; ++r) { v += park[(r * 2 + which) * p.BN + c]; park[(r * 2 + which) * p.BN + c] = 0.f; }
              atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
            }
            ptx::named_bar_sync(bar_id, 128);
          }
          s_nb = nb;
        }
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
        const uint32_t t_base = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(q * 32) << 16);
        const bool t_on = p.tf != nullptr && valid;
        for (int j = 0; j < p.BN / 32; ++j) {
          uint32_t v[32];
          ptx::tmem_ld32(t_base + (uint32_t)(j * 32), v);
          ptx::tmem_ld_wait();
          const int col = nb + j * 32;
          const int c0 = chunk_cb(col);
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(vec + c0 + i);
            f[i] = __uint_as_float(v[i]) + b4.x; f[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + b4.z; f[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
          }
          if (t_on) {
            const float4* tp4 = reinterpret_cast<const float4*>(p.tf + chunk_pix(col) * p.t_ld + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 t4 = tp4[i];
              const int c = 4 * i;
              f[c] += vec[vlen + c0 + c] * t4.x + vec[2 * vlen + c0 + c];
              f[c + 1] += vec[vlen + c0 + c + 1] * t4.y + vec[2 * vlen + c0 + c + 1];
              f[c + 2] += vec[vlen + c0 + c + 2] * t4.z + vec[2 * vlen + c0 + c + 2];
              f[c + 3] += vec[vlen + c0 + c + 3] * t4.w + vec[2 * vlen + c0 + c + 3];
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (p.post) {          // eval-mode BatchNorm behind the ReLU: z = a * relu(conv) + b (unet.py:213-215)
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaf(vec[vlen + c0 + i], f[i], vec[2 * vlen + c0 + i]);
          }
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = 0.f;
          }
          uint8_t* stg = smem + grp_staging_off + (uint32_t)sbuf * staging_bytes;
          const uint32_t stg_addr = smem_base + grp_staging_off + (uint32_t)sbuf * staging_bytes;
          if (et == 0) ptx::tma_store_wait_read1();       // the store issued two chunks ago has finished reading this tile
          ptx::named_bar_sync(bar_id, 128);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(stg + (uint32_t)row * 128u + (uint32_t)((g ^ (row & 7)) << 4)) =
                make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (et == 0) {
            if (p.c5) {
              const int ab = col >> p.cst_shift;
              ptx::tma_store_5d(&tmC, stg_addr, col & (p.Cst - 1), ab & 1, w0, ab >> 1, h0);
            } else {
              ptx::tma_store_4d(&tmC, stg_addr, col, w0, h0, n0);
            }
            ptx::tma_store_commit();
          }
          if (p.stat) {
            // column c of this chunk over 32 rows (rows outside the image hold zeros)
            const int c = et & 31, rgp = et >> 5;
            float a0 = 0.f, b0 = 0.f;
#pragma unroll 8
            for (int r = rgp * 32; r < rgp * 32 + 32; ++r) {
              const float x = *reinterpret_cast<const float*>(stg + (uint32_t)r * 128u + (uint32_t)((((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4));
              a0 += x; b0 = fmaf(x, x, b0);
            }
            park[(rgp * 2 + 0) * p.BN + j * 32 + c] += a0;
            park[(rgp * 2 + 1) * p.BN + j * 32 + c] += b0;
          }
          sbuf ^= 1;
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(tempty_bar(acc));
        acc += G; if (acc >= NACC) { acc -= NACC; acc_phase ^= 1u; }
        continue;
      }
      if (p.stat && nb != s_nb) { flush_stats(); s_nb = nb; }
      // the second epilogue operand (residual / accumulate) is fetched before waiting for the accumulator,
      // so its global-memory latency hides behind the MMAs (BN <= 128: 16 x 16 B per thread)
      if (nstg == 2) {
        // double-buffered staging: only the store issued two tiles ago must have finished reading
        staging = smem + grp_staging_off + (uint32_t)sbuf * staging_bytes;
        staging_addr = smem_base + grp_staging_off + (uint32_t)sbuf * staging_bytes;
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        ptx::named_bar_sync(bar_id, 128);
        sbuf ^= 1;
      }
      uint4 tnext[4];
      const bool t_sm = p.t_smem != 0;
      // (the G = 4 instantiation runs at the 96-register limit: there `t` only ever comes through shared memory --
      //  the launcher falls back to G = 2 otherwise -- and the register-resident global-load pipeline below is not compiled)
      const bool t_on = G < 4 && p.t != nullptr && valid && !t_sm;
      const uint8_t* tbuf = smem + tbuf_off + (uint32_t)(grp * kTcTBufs + tb) * staging_bytes;
      if (t_sm) ptx::mbar_wait(tt_full(grp * kTcTBufs + tb), tph);
      if (t_on) {
        const uint4* tp4 = reinterpret_cast<const uint4*>(p.t + chunk_pix(nb) * p.t_ld + chunk_cb(nb));
#pragma unroll
        for (int i = 0; i < 4; ++i) tnext[i] = tp4[i];
      }
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(q * 32) << 16);
      // (packed fp32 adds / FMAs and the ReLU on the packed bf16 pairs: see tc_conv3_kernel)
      const bool relu_packed = p.relu && !p.post;
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      for (int j = 0; j < p.BN / 32; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(t_base + (uint32_t)(j * 32), v);
        ptx::tmem_ld_wait();
        const int c0 = chunk_cb(nb + j * 32);
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(vec + c0 + i);
          const float2 r0 = fu_add2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(b4.x, b4.y));
          const float2 r1 = fu_add2(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), make_float2(b4.z, b4.w));
          f[i] = r0.x; f[i + 1] = r0.y; f[i + 2] = r1.x; f[i + 3] = r1.y;
        }
        if (t_on) {
          uint4 tcur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) tcur[i] = tnext[i];
          if (j + 1 < p.BN / 32) {            // software pipeline: next chunk's operand is in flight during this one
            const int ncol = nb + (j + 1) * 32;
            const uint4* tp4 = reinterpret_cast<const uint4*>(p.t + chunk_pix(ncol) * p.t_ld + chunk_cb(ncol));
#pragma unroll
            for (int i = 0; i < 4; ++i) tnext[i] = tp4[i];
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const uint4 u = tcur[i / 8];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const float2 tf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k2]));
              const int c = i + 2 * k2;
              f[c] += vec[vlen + c0 + c] * tf.x + vec[2 * vlen + c0 + c];
              f[c + 1] += vec[vlen + c0 + c + 1] * tf.y + vec[2 * vlen + c0 + c + 1];
            }
          }
        }
        // 32 channels <-> 4 x 16 bytes of the swizzled staging tile (rows outside the image: zeros)
        const int colt = j * 32;                 // column within the BN tile
        const int sub = colt >> p.cs_shift;      // store box this chunk belongs to
        const uint32_t row_off = (uint32_t)row * pitch + (uint32_t)(colt & (p.CS - 1)) * 2u;
        if (t_sm) {
          // second operand from its shared-memory tile (same layout as the staging tile): f += a * t + b
          const uint8_t* src = tbuf + (uint32_t)sub * sub_bytes;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t logical = row_off + (uint32_t)g * 16u;
            const uint32_t phys = logical ^ (((logical >> 7) & swz_mask) << 4);
            const uint4 u = *reinterpret_cast<const uint4*>(src + phys);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const float2 tf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k2]));
              const int c = g * 8 + 2 * k2;
              f[c] += vec[vlen + c0 + c] * tf.x + vec[2 * vlen + c0 + c];
              f[c + 1] += vec[vlen + c0 + c + 1] * tf.y + vec[2 * vlen + c0 + c + 1];
            }
          }
        }
        if (p.post) {            // eval-mode BatchNorm behind the ReLU: z = a * relu(conv) + b (unet.py:213-215)
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaf(vec[vlen + c0 + i], f[i], vec[2 * vlen + c0 + i]);
        }
        uint8_t* dst = staging + (uint32_t)sub * sub_bytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
          if (valid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[g * 8 + 2 * k], f[g * 8 + 2 * k + 1]);
              if (relu_packed) h2 = __hmax2(h2, zero2);
              pk[k] = *reinterpret_cast<uint32_t*>(&h2);
            }
          } else {
            pk[0] = pk[1] = pk[2] = pk[3] = 0u;
          }
          const uint32_t logical = row_off + (uint32_t)g * 16u;
          const uint32_t phys = logical ^ (((logical >> 7) & swz_mask) << 4);
          *reinterpret_cast<uint4*>(dst + phys) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar(acc));
      acc += G; if (acc >= NACC) { acc -= NACC; acc_phase ^= 1u; }
      if (t_sm) {                     // `t` tile consumed: the producer may refill it (tile i + kTcTBufs * G)
        ptx::mbar_arrive(tt_empty(grp * kTcTBufs + tb));
        if (++tb == kTcTBufs) { tb = 0; tph ^= 1u; }
      }
      // staging tile complete -> TMA store
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(bar_id, 128);
      if (et == 0) {
        for (int s = 0; s < p.BN / p.CS; ++s) {
          if (p.c5) {
            const int col = nb + s * p.CS, ab = col >> p.cst_shift;
            ptx::tma_store_5d(&tmC, staging_addr + (uint32_t)s * sub_bytes, col & (p.Cst - 1), ab & 1, w0, ab >> 1, h0);
          }
          else ptx::tma_store_4d(&tmC, staging_addr + (uint32_t)s * sub_bytes, nb + s * p.CS, w0, h0, n0);
        }
        ptx::tma_store_commit();
      }
      if (p.stat) {
        // per-channel statistics of the stored (bf16-rounded) tile; invalid rows hold zeros.
        // thread = (column pair, row group of BN/2 rows)
        const int c = 2 * cp;
        const int sub = c / p.CS;
        const uint32_t bir = (uint32_t)(c % p.CS) * 2u;
        const int r0 = rg * cpairs;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 4
        for (int r = r0; r < r0 + cpairs; ++r) {
          const uint32_t logical = (uint32_t)r * pitch + bir;
          const uint32_t phys = logical ^ (((logical >> 7) & swz_mask) << 4);
          const float2 x2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(staging + (uint32_t)sub * sub_bytes + phys));
          a0 += x2.x; a1 += x2.y; b0 = fmaf(x2.x, x2.x, b0); b1 = fmaf(x2.y, x2.y, b1);
        }
        s0 += a0; s1 += a1; q0 += b0; q1 += b1;
      }
      if (nstg == 1) {
        if (et == 0) ptx::tma_store_wait_read();     // staging may be overwritten after this
        ptx::named_bar_sync(bar_id, 128);
      } else if (p.stat) {
        ptx::named_bar_sync(bar_id, 128);                 // statistics readers are done with this tile
      }
    }
    if (nstg == 2) { if (et == 0) ptx::tma_store_wait_read(); ptx::named_bar_sync(bar_id, 128); }
    if constexpr (F32) {
      if (p.stat && s_nb >= 0) {
        float* park = reinterpret_cast<float*>(smem + park_off) + grp * 8 * p.BN;
        for (int i = et; i < 2 * p.BN; i += 128) {
          const int which = i / p.BN, c = i - which * p.BN;
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) v += park[(r * 2 + which) * p.BN + c];
          atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
        }
      }
    } else {
      flush_stats();
    }
    if (et == 0) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (p.mc) ptx::cluster_sync();          // no CTA leaves while its peer may still signal its barriers
  if (warp == mma_warp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ===========================================================================
// 3x3 convolution, second generation ("halo" kernel): forward and data gradient at levels whose width is
// >= 24.  v1 above re-loads the input tile for each of the 9 taps and the weights for each pixel tile and is
// bound by L2->SM traffic (profiles/r01_*).  Here
//  * the input of a tile is loaded ONCE per 64-channel chunk as a halo box (twb x (th+2) pixels, twb = two+2)
//    whose rows are the flattened padded-width pixel grid; tap (kh,kw) is the same shared-memory tile read
//    through an MMA descriptor whose start address is advanced by kh*twb + kw rows (the swizzle is a
//    function of the absolute shared-memory address, so a row-shifted view stays consistent).  The two
//    padded columns of every row produce garbage accumulator rows that are never stored;
//  * two pixel tiles share every weight tile (2 accumulators per TMEM stage);
//  * for thin layers (all taps fit in shared memory) the weights are loaded once per CTA and stay resident.
// ===========================================================================
// Per-channel sum / sum of squares of a bf16 staging tile: 128 rows x (nsub sub-boxes of CS channels), rows of PITCH =
// 2*CS bytes swizzled the way the TMA store expects.  Warp `q` scans rows [32q, 32q + 32) of every sub-box; one load
// instruction of the warp covers 128 consecutive bytes (one 128-byte row, or two 64-byte rows), so it is bank-conflict
// free, and the swizzle term of load i is a compile-time constant (i & 7 resp. i & 3): 8 instructions per load instead
// of the 28 of a generic address computation -- this loop was 36 % of all instructions the thin forward layers executed
// (ncu source page, profiles/r02_ncu_thin_conv_before.txt).  The lane keeps one column pair per sub-box:
// acc[sub] = {sum even ch, sum odd ch, sum of squares even, odd}.  Rows that are never written must hold zeros.
__host__ __device__ inline uint32_t tc3_park_floats(int BN, int CS) {
  const uint32_t a = 8u * (uint32_t)BN, b = 512u * (uint32_t)(BN / CS);
  return a > b ? a : b;
}
template <int PITCH>
__device__ __forceinline__ void tc_stats_scan(const uint8_t* tile, uint32_t sub_bytes, int nsub, int q, int lane,
                                              float (&acc)[4][4]) {
  constexpr int RPL = 128 / PITCH;               // rows per load instruction
  constexpr int CPW = PITCH / 4;                 // column pairs per row
  constexpr uint32_t SMASK = PITCH == 128 ? 7u : 3u;
  const uint32_t base = (uint32_t)(32 * q + lane / CPW) * PITCH + (uint32_t)(lane % CPW) * 4u;
  // (eight loads per trip = one swizzle period; not unrolled further: code size, see tc_mma_taps_resident)
#pragma unroll
  for (int sub = 0; sub < 4; ++sub) {
    if (sub < nsub) {
      const uint8_t* tp = tile + (uint32_t)sub * sub_bytes;
      float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int i0 = 0; i0 < 32 / RPL; i0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (base ^ (((uint32_t)k & SMASK) << 4)) + (uint32_t)(i0 + k) * 128u;
          const float2 x2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(tp + off));
          a = fu_add2(a, x2); b = fu_fma2(x2, x2, b);          // packed fp32: two instructions per column pair
        }
      }
      acc[sub][0] += a.x; acc[sub][1] += a.y; acc[sub][2] += b.x; acc[sub][3] += b.y;
    }
  }
}

struct TcConv3Params {
  int B, H, W, K, N;
  int KC, BN, CS;
  int twb, two, th;             // box width, valid output width (twb-2), tile height; twb*th <= 128
  int tiles_w, tiles_h, n_tiles;
  int halo1;                    // 1: one halo load per chunk (kw by row shift); 0: one load per (chunk, kw)
  int npair;                    // pixel tiles per weight tile (1 or 2)
  int resident;                 // weights resident in shared memory (n_tiles == 1)
  int res;                      // 1: a second source (tmA2 / tmB2) is accumulated as one more, centre-tap-only, set of K chunks:
                                //    dX = conv3x3^T(dY) + conv1x1^T(G), the data gradients of a residual block's first 3x3 conv and
                                //    of its 1x1 shortcut (unet.py:229-231) in ONE pass over dX instead of a write + read-modify-write
  int a_stages, b_stages;
  unsigned a_tile_bytes;        // one halo tile, rounded up to 1024 B
  int relu;
  const float* bias;
  const bf16* t; int t_ld;
  const float* bn_a; const float* bn_b;
  double* stat;
  long long* dbg;               // optional timeline buffer (FU_TC_DBG=1): [role][super][event] clock64 of CTA 0
  // split-bf16 x3 parity mode (template F32 = true, see TcConvParams): lo-half offsets of the A / weight maps of
  // both sources, fp32 second epilogue operand
  int a_lo, b_lo, a2_lo, b2_lo;
  const float* tf;
  int post;                     // see TcConvParams
  int dual;                     // 1: two MMA-issuing warps take alternate super tiles.  Only when an issuer that runs one
                                // super tile ahead can never be a whole ring round ahead of the other (mbarrier parity
                                // waits cannot tell round r from round r - 2): resident weights and a_stages >= 2 x the A
                                // slots of one super tile
  int nstg;                     // bf16 staging tiles per epilogue group (2: the TMA store of tile i drains under tile i + 1)
  FastDiv fd_ntiles, fd_timg, fd_tw;   // divisions by n_tiles, tiles_w * tiles_h, tiles_w (tile decode, every role, every tile)
  int cs_shift;                 // log2(CS)
  int baton_kh;                 // filter row (0..2) of a super tile's last section at which the issuer passes the baton
  int stack;                    // 1: the three kw taps of a filter row are ONE MMA with N = 3 * BN (the resident weight tiles of a
                                //    filter row are adjacent in shared memory, i.e. already one [3 * BN][KC] operand) against the
                                //    kw = 0 view of the halo tile: accumulator block kw of box position m holds X[m + kh*twb] * W[kh][kw],
                                //    and the epilogue forms out[m] = blk0[m] + blk1[m + 1] + blk2[m + 2] with two warp shuffles per
                                //    column.  The A operand (128 rows re-read by every MMA whatever its N) is fetched three times
                                //    less often.  Needs twb | 32 (valid columns never reach across a warp's 32 lanes), BN = 32.
  int nacc;                     // TMEM accumulator stages (2 * S; S when stacked: 3 * BN columns per tile)
  int planeC;                   // > 0: the output is PLANAR -- channels [k*planeC, (k+1)*planeC) form a contiguous (pixels, planeC)
                                // tensor k (5-D store map with the plane index as its last coordinate); CS == planeC
};

// S = epilogue sets.  A set is one group of 4 warps per pixel tile of the pair; super tile i of a CTA is drained
// by set i % S from TMEM stage i % (2S).  The thin layers (BN <= 64: 18-72 MMAs per tile) are bound by the
// epilogue's per-warp latency chain, not by the tensor pipe (FU_TC_DBG timeline: ~3600 cycles per super tile of
// which the MMAs take 2000), so they run two sets.
// STK: the kw-stacked form (TcConv3Params::stack) -- a template parameter because the two-set instantiation runs at its
// register cap: with the stacked epilogue merely compiled in, every other thin layer paid 120 instead of 32 bytes of spills
// (same-box A/B of the whole step: 5.235 vs 5.194 ms).
template <int S, bool F32 = false, bool STK = false>
__global__ void __launch_bounds__(96 + 256 * S, 1)
tc_conv3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB2, const TcConv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw);
  // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so role-local scalars can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // Warp roles.  The epilogue warps come FIRST and the TMA producer / MMA issuer LAST: the SM sub-partition schedulers
  // prefer the highest warp id among eligible warps, and the single MMA-issuing thread is the critical path of the
  // thin layers -- as warp 1 it was starved by the (instruction-heavy) epilogue warps sharing its scheduler (ncu:
  // ~1000 idle tensor-pipe cycles between tiles, profiles/r02_ncu_thin_conv_before.txt).
  constexpr int prod_warp = 8 * S, mma_warp = 8 * S + 1;
  const uint32_t row_bytes = (uint32_t)p.KC * 2u;
  const uint32_t b_bytes = (uint32_t)p.BN * row_bytes;
  const int rchunks = p.K / p.KC;                                  // real channel chunks
  // split mode: virtual chunk v = pass * rchunks + c; pass 0: (A_hi, W_hi), 1: (A_hi, W_lo), 2: (A_lo, W_hi)
  const int cchunks = F32 ? 3 * rchunks : rchunks;
  const int wchunks = F32 ? 2 * rchunks : rchunks;                 // distinct weight chunks (hi | lo)
  auto a_coord = [&](int v, int lo) { return F32 ? (v % rchunks) * p.KC + (v / rchunks == 2 ? lo : 0) : v * p.KC; };
  auto w_index = [&](int v) { return F32 ? (v / rchunks == 1 ? rchunks + v % rchunks : v % rchunks) : v; };
  auto w_coord = [&](int wi, int lo) { return F32 ? (wi % rchunks) * p.KC + (wi >= rchunks ? lo : 0) : wi * p.KC; };
  const uint32_t a_stage_bytes = (uint32_t)p.npair * p.a_tile_bytes;
  const uint32_t a_off = 0;
  const uint32_t b_off = a_off + (uint32_t)p.a_stages * a_stage_bytes;
  const int wtiles = wchunks * 9 + (p.res ? wchunks : 0);          // resident weight tiles: 9 taps (+ the 1x1) per chunk
  const uint32_t b_region = p.resident ? (uint32_t)wtiles * b_bytes : (uint32_t)p.b_stages * b_bytes;
  const uint32_t staging_off = b_off + b_region;
  // bf16: 128 x BN tile, double buffered when BN <= 64 (the TMA store of tile i drains while tile i + 1 is converted);
  // F32: 128 rows x 32 fp32 columns (16 KB), two per epilogue group
  const uint32_t tile_bytes = 128u * (uint32_t)p.BN * 2u;
  const int nstg = p.nstg;
  const uint32_t staging_bytes = F32 ? 32768u : (uint32_t)nstg * tile_bytes;      // per epilogue group
  const int NACC = p.nacc;                 // TMEM accumulator stages (each: npair tiles of acc_cols columns)
  const uint32_t acc_cols = (uint32_t)p.BN * (STK ? 3u : 1u);
  const uint32_t vec_off = staging_off + (uint32_t)(S * p.npair) * staging_bytes;   // bias | bn_a | bn_b, [3][N] floats
  // statistics partials per epilogue group (tc3_park_floats): F32 [4][2][BN] floats; bf16 one float4 per (thread, sub-box)
  const uint32_t stat_off = vec_off + 3u * (uint32_t)p.N * 4u;
  // (bf16: the partials are parked in the group's own staging tile at flush time, no separate region)
  const uint32_t park_floats = F32 ? tc3_park_floats(p.BN, p.CS) : 0u;
  const uint32_t bar_off = (stat_off + (uint32_t)(S * p.npair) * park_floats * 4u + 7u) & ~7u;
  const uint32_t bar_base = smem_base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(8 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (uint32_t)(16 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (uint32_t)(24 + s); };
  auto t_full = [&](int a) { return bar_base + 8u * (uint32_t)(32 + a); };
  auto t_empty = [&](int a) { return bar_base + 8u * (uint32_t)(36 + a); };
  const uint32_t res_bar = bar_base + 8u * 40u;
  auto baton_bar = [&](int x) { return bar_base + 8u * (uint32_t)(42 + x); };
  const uint32_t slot_addr = bar_base + 8u * 41u;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * 41u);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(NACC * p.npair) * acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) {
      ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1);
      ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < NACC; ++a) { ptx::mbar_init(t_full(a), 1); ptx::mbar_init(t_empty(a), 128u * (uint32_t)p.npair); }
    ptx::mbar_init(res_bar, 1);
    ptx::mbar_init(baton_bar(0), 1); ptx::mbar_init(baton_bar(1), 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmB); ptx::prefetch_tmap(&tmC);
    if (p.res) { ptx::prefetch_tmap(&tmA2); ptx::prefetch_tmap(&tmB2); }
  }
  if (warp == mma_warp) { ptx::tmem_alloc(slot_addr, tmem_cols); ptx::tmem_relinquish(); }
  pdl_wait(); pdl_trigger();          // (see tc_conv_kernel)
  {
    // per-channel epilogue vectors live in shared memory for the whole (persistent) CTA: the epilogue warps
    // run one per scheduler, so a global/L1 load in their dependency chain is an exposed long-scoreboard stall
    float* vec = reinterpret_cast<float*>(smem + vec_off);
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      vec[i] = p.bias ? p.bias[i] : 0.f;
      vec[p.N + i] = p.bn_a ? p.bn_a[i] : 1.f;
      vec[2 * p.N + i] = p.bn_a ? p.bn_b[i] : 0.f;
    }
    if (F32) {
      float* park = reinterpret_cast<float*>(smem + stat_off);
      for (int i = threadIdx.x; i < S * p.npair * (int)park_floats; i += blockDim.x) park[i] = 0.f;
    } else if (p.stat) {
      // rows of the staging tiles that no valid pixel maps to are never written: they must read as zeros in the
      // statistics scan
      uint4* z = reinterpret_cast<uint4*>(smem + staging_off);
      for (int i = threadIdx.x; i < (int)((uint32_t)(S * p.npair) * staging_bytes / 16u); i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

// (timeline stamps of CTA 0, FU_TC_DBG=1: compiled in only with -DFU_TC_TIMELINE=1 -- the checks sat in every role's per-tile
//  path of a kernel whose two-set instantiation runs at its register cap)
#if FU_TC_TIMELINE
#define FU_DBG(role, idx, ev)                                                                       \
  do {                                                                                              \
    if (p.dbg && blockIdx.x == 0 && (idx) < 24) p.dbg[((role) * 24 + (idx)) * 4 + (ev)] = clock64(); \
  } while (0)
#else
#define FU_DBG(role, idx, ev) do { } while (0)
#endif
  const int tiles_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_img * p.B;
  const int supers_m = (m_tiles + p.npair - 1) / p.npair;
  const int total_super = supers_m * p.n_tiles;
  const int groups = p.halo1 ? 1 : 3;          // A loads per chunk
  const int taps_per_group = p.halo1 ? 9 : 3;
  const uint32_t box_bytes = (uint32_t)(p.twb * (p.th + 2)) * row_bytes;

  auto decode_m = [&](int mt, int& w0, int& h0, int& n) {
    int r, hq, wq;
    p.fd_timg.divmod(mt, n, r);
    p.fd_tw.divmod(r, hq, wq);
    h0 = hq * p.th;
    w0 = wq * p.two;
  };

  if (warp == prod_warp) {
    // ------------------------------ TMA producer (whole warp loops, one elected lane issues) -----
    {
      if (p.resident && ptx::elect_one()) {
        ptx::mbar_expect_tx(res_bar, (uint32_t)wtiles * b_bytes);
        for (int c = 0; c < wchunks; ++c)
          for (int tap = 0; tap < 9; ++tap)
            ptx::tma_load_3d(smem_base + b_off + (uint32_t)(c * 9 + tap) * b_bytes, &tmB, res_bar, w_coord(c, p.b_lo), tap, 0);
        if (p.res)
          for (int c = 0; c < wchunks; ++c)
            ptx::tma_load_3d(smem_base + b_off + (uint32_t)(wchunks * 9 + c) * b_bytes, &tmB2, res_bar, w_coord(c, p.b2_lo), 0, 0);
      }
      __syncwarp();
      int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
      for (int st = blockIdx.x; st < total_super; st += gridDim.x) {
        int nt, sp;
        p.fd_ntiles.divmod(st, sp, nt);
        const int nb = nt * p.BN;
        for (int c = 0; c < cchunks; ++c) {
          for (int g = 0; g < groups; ++g) {
            ptx::mbar_wait(a_empty(as), aph ^ 1u);
            if (ptx::elect_one()) {
              if (c == 0 && g == 0) FU_DBG(0, (st - (int)blockIdx.x) / (int)gridDim.x, 0);
              int nvalid = 0;
              for (int i = 0; i < p.npair; ++i) nvalid += (sp * p.npair + i) < m_tiles ? 1 : 0;
              ptx::mbar_expect_tx(a_full(as), (uint32_t)nvalid * box_bytes);
              for (int i = 0; i < p.npair; ++i) {
                const int mt = sp * p.npair + i;
                if (mt >= m_tiles) break;
                int w0, h0, n;
                decode_m(mt, w0, h0, n);
                ptx::tma_load_4d(smem_base + a_off + (uint32_t)as * a_stage_bytes + (uint32_t)i * p.a_tile_bytes, &tmA,
                                 a_full(as), a_coord(c, p.a_lo), w0 - 1 + (p.halo1 ? 0 : g), h0 - 1, n);
              }
            }
            __syncwarp();
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            if (!p.resident) {
              for (int tt = 0; tt < taps_per_group; ++tt) {
                const int tap = p.halo1 ? tt : tt * 3 + g;      // (kh = tt, kw = g) when one load per kw
                ptx::mbar_wait(b_empty(bs), bph ^ 1u);
                if (ptx::elect_one()) {
                  ptx::mbar_expect_tx(b_full(bs), b_bytes);
                  ptx::tma_load_3d(smem_base + b_off + (uint32_t)bs * b_bytes, &tmB, b_full(bs), w_coord(w_index(c), p.b_lo), tap, nb);
                }
                __syncwarp();
                if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
              }
            }
          }
        }
        if (p.res) {
          // second source: the same halo boxes of G (only the centre view is used), one weight tile per chunk
          for (int c = 0; c < cchunks; ++c) {
            ptx::mbar_wait(a_empty(as), aph ^ 1u);
            if (ptx::elect_one()) {
              int nvalid = 0;
              for (int i = 0; i < p.npair; ++i) nvalid += (sp * p.npair + i) < m_tiles ? 1 : 0;
              ptx::mbar_expect_tx(a_full(as), (uint32_t)nvalid * box_bytes);
              for (int i = 0; i < p.npair; ++i) {
                const int mt = sp * p.npair + i;
                if (mt >= m_tiles) break;
                int w0, h0, n;
                decode_m(mt, w0, h0, n);
                ptx::tma_load_4d(smem_base + a_off + (uint32_t)as * a_stage_bytes + (uint32_t)i * p.a_tile_bytes, &tmA2,
                                 a_full(as), a_coord(c, p.a2_lo), w0 - 1, h0 - 1, n);
              }
            }
            __syncwarp();
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            if (!p.resident) {
              ptx::mbar_wait(b_empty(bs), bph ^ 1u);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx(b_full(bs), b_bytes);
                ptx::tma_load_3d(smem_base + b_off + (uint32_t)bs * b_bytes, &tmB2, b_full(bs), w_coord(w_index(c), p.b2_lo), 0, nb);
              }
              __syncwarp();
              if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp >= mma_warp) {
    // ------------------------------ MMA issuers (two warps, alternating super tiles) ------------------------------
    // The whole warp runs the loop; one elected lane issues.  Between two super tiles an issuer spends ~2000 cycles in
    // serial barrier round trips (accumulator free -> fence -> operands landed -> ... -> commit; each already-complete
    // mbarrier wait costs 300-500 cycles through the busy shared-memory pipe) while the tensor pipe, whose queue holds
    // only ~4 MMAs, runs dry: FU_TC_DBG timeline of 32->32 @192x192: 1750 cycles of MMAs per 4050-cycle super tile.
    // Two issuers take alternate super tiles (each tracks the rings of the tiles it skips), so one issuer's round trips
    // overlap the other's MMAs; the MMAs of consecutive super tiles go to different accumulators and are independent.
    // Left to themselves the two issuers fall into lock step (both issue into the one tensor pipe at the same time, both
    // finish together, both epilogue sets wake together, and both issuers then crawl through their ~1200-4000 cycle
    // control path while the pipe idles: FU_TC_DBG timeline of 32->32 @192x192, round 2: MMA phase 3000 cycles, then
    // 3900 idle).  p.dual == 2: a baton (two mbarriers) makes them alternate strictly -- an issuer does its waits for
    // super tile i + 1 while the other issues tile i, starts only when the other has reached its last filter row, and
    // the epilogue sets are de-phased with them.
    {
      const int mw = warp - mma_warp;               // issuer 0 / 1
      const bool use_baton = p.dual == 2;
      uint32_t bat_phase = 0; bool bat_first = mw == 0;
      int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc1 = umma_idesc_bf16((uint32_t)p.BN);                 // one tap (and the 1x1 second source)
      const uint32_t idesc = STK ? umma_idesc_bf16(3u * (uint32_t)p.BN) : idesc1;
      const uint64_t dbase = umma_desc_kmajor(0, row_bytes);
      const int ksteps = p.KC / 16;
      const int npair = p.npair, resident = p.resident, halo1 = p.halo1, a_stages = p.a_stages, b_stages = p.b_stages;
      const uint32_t a_tile16 = p.a_tile_bytes >> 4, b16 = b_bytes >> 4;
      // offset (in 16-byte units) of a tap's view into the halo tile: kh image rows + kw pixels
      const uint32_t kh16 = ((uint32_t)p.twb * row_bytes) >> 4, kw16 = halo1 ? (row_bytes >> 4) : 0u;
      const uint32_t dhi = (uint32_t)(dbase >> 32), dlo = (uint32_t)dbase;      // descriptor words shared by every tile
      if (resident) { ptx::mbar_wait(res_bar, 0); ptx::tc_fence_after(); }
      // ring slots one super tile consumes (the same for every super tile)
      const int a_per_tile = cchunks * groups + (p.res ? cchunks : 0);
      const int b_per_tile = resident ? 0 : cchunks * groups * taps_per_group + (p.res ? cchunks : 0);
      int it = 0;
      for (int st = blockIdx.x; st < total_super; st += gridDim.x, ++it) {
        if (!p.dual) { if (mw) break; }
        else if ((it & 1) != mw) {                  // the other issuer's super tile: step over its ring slots
          as += a_per_tile; while (as >= a_stages) { as -= a_stages; aph ^= 1u; }
          bs += b_per_tile; while (bs >= b_stages) { bs -= b_stages; bph ^= 1u; }
          if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
        const int sp = p.fd_ntiles.div(st);
        const bool two = npair == 2 && (sp * 2 + 1) < m_tiles;
        if (lane == 0) FU_DBG(0, (st - (int)blockIdx.x) / (int)gridDim.x, 3);     // (dbg) arrived at the accumulator wait
        ptx::mbar_wait(t_empty(acc), acc_phase ^ 1u);
        if (lane == 0) FU_DBG(0, (st - (int)blockIdx.x) / (int)gridDim.x, 1);     // (dbg) accumulator free, before the fence
        ptx::tc_fence_after();
        if (lane == 0) FU_DBG(1, (st - (int)blockIdx.x) / (int)gridDim.x, 0);
        const uint32_t d0 = tmem_base + (uint32_t)(acc * npair) * acc_cols, d1 = d0 + acc_cols;
        uint32_t accum = 0;                       // 0 only for the very first MMA into each accumulator
        for (int c = 0; c < cchunks; ++c) {
          for (int g = 0; g < groups; ++g) {
            ptx::mbar_wait(a_full(as), aph);
            if (lane == 0 && c == 0 && g == 0) FU_DBG(1, (st - (int)blockIdx.x) / (int)gridDim.x, 1);
            if (use_baton && c == 0 && g == 0) {          // my turn at the tensor pipe (issuer 0 opens)
              if (!bat_first) { ptx::mbar_wait(baton_bar(mw), bat_phase); bat_phase ^= 1u; }
              bat_first = false;
            }
            const uint32_t a_lo0 = dlo + ((smem_base + a_off + (uint32_t)as * a_stage_bytes) >> 4);
            const bool last_group = c == cchunks - 1 && g == groups - 1 && !p.res;
            if (resident) {
              // all 9 taps straight from the resident weight region: one elected section per A stage
              if (ptx::elect_one()) {
                if (c == 0 && g == 0) FU_DBG(0, (st - (int)blockIdx.x) / (int)gridDim.x, 2);   // (dbg) first MMA about to issue
                const uint32_t b_lo0 = dlo + ((smem_base + b_off + (uint32_t)(w_index(c) * 9) * b_bytes) >> 4);
                // (halo1: all nine taps; otherwise one A load per kw and only the taps with kw == g)
                // (stacked: one N = 3 * BN MMA per filter row = the "taps with kw == 0" walk over three-tile-wide weight rows)
                tc_mma_taps_resident(ksteps, d0, d1, two, a_lo0, a_tile16, b_lo0, b16, kh16, kw16, STK ? 0 : (halo1 ? -1 : g), dhi, idesc, accum,
                                     (use_baton && last_group) ? baton_bar(mw ^ 1) : 0u, p.baton_kh);
                ptx::umma_commit(a_empty(as));
                if (last_group) { ptx::umma_commit(t_full(acc)); FU_DBG(1, (st - (int)blockIdx.x) / (int)gridDim.x, 2); }
              }
              __syncwarp();
              accum = 1;
            } else {
              for (int tt = 0; tt < taps_per_group; ++tt) {
                const int tap = halo1 ? tt : tt * 3 + g;
                ptx::mbar_wait(b_full(bs), bph);
                if (ptx::elect_one()) {
                  const uint32_t kh = (uint32_t)tap / 3u, kw = (uint32_t)tap - 3u * kh;
                  const uint32_t a_lo = a_lo0 + kh * kh16 + kw * kw16;
                  const uint32_t b_lo = dlo + ((smem_base + b_off + (uint32_t)bs * b_bytes) >> 4);
                  if (ksteps == 4) tc_mma_tap<4>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc, accum);
                  else if (ksteps == 2) tc_mma_tap<2>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc, accum);
                  else tc_mma_tap<1>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc, accum);
                  ptx::umma_commit(b_empty(bs));
                  if (tt == taps_per_group - 1) {
                    ptx::umma_commit(a_empty(as));
                    if (last_group) { ptx::umma_commit(t_full(acc)); FU_DBG(1, (st - (int)blockIdx.x) / (int)gridDim.x, 2); }
                  }
                }
                __syncwarp();
                accum = 1;
                if (++bs == b_stages) { bs = 0; bph ^= 1u; }
              }
            }
            if (++as == a_stages) { as = 0; aph ^= 1u; }
          }
        }
        if (p.res) {
          // + conv1x1^T(G): centre view of G's halo tile times the shortcut's weights, one K chunk at a time
          for (int c = 0; c < cchunks; ++c) {
            ptx::mbar_wait(a_full(as), aph);
            if (!resident) ptx::mbar_wait(b_full(bs), bph);
            if (ptx::elect_one()) {
              const uint32_t a_lo = dlo + ((smem_base + a_off + (uint32_t)as * a_stage_bytes) >> 4) +
                                    (((uint32_t)(p.twb + 1) * row_bytes) >> 4);               // view (kh,kw) = (1,1)
              const uint32_t b_addr = resident ? smem_base + b_off + (uint32_t)(wchunks * 9 + w_index(c)) * b_bytes
                                               : smem_base + b_off + (uint32_t)bs * b_bytes;
              const uint32_t b_lo = dlo + (b_addr >> 4);
              if (use_baton && c == cchunks - 1) ptx::mbar_arrive(baton_bar(mw ^ 1));     // last section of this super tile
              // (stacked: into block 0 of the accumulator, which the epilogue reads unshifted)
              if (ksteps == 4) tc_mma_tap<4>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc1, accum);
              else if (ksteps == 2) tc_mma_tap<2>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc1, accum);
              else tc_mma_tap<1>(d0, d1, two, a_lo, a_lo + a_tile16, b_lo, dhi, idesc1, accum);
              if (!resident) ptx::umma_commit(b_empty(bs));
              ptx::umma_commit(a_empty(as));
              if (c == cchunks - 1) { ptx::umma_commit(t_full(acc)); FU_DBG(1, (st - (int)blockIdx.x) / (int)gridDim.x, 2); }
            }
            __syncwarp();
            accum = 1;
            if (!resident) { if (++bs == b_stages) { bs = 0; bph ^= 1u; } }
            if (++as == a_stages) { as = 0; aph ^= 1u; }
          }
        }
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------ epilogue: S sets of one group of 4 warps per pixel tile of the pair -----
    const int gi = warp >> 2;           // group index 0 .. 2S-1
    const int grp = gi & 1;                   // which pixel tile of the super tile
    const int set = gi >> 1;                  // this set drains super tiles set, set + S, ... of the CTA
    if (grp < p.npair) {
      const int q = warp & 3;                 // TMEM lane quadrant this warp may access
      const int row = q * 32 + lane;
      const int et = threadIdx.x & 127;   // thread index within the group
      const int bar_id = 1 + gi;
      const uint32_t pitch = (uint32_t)p.CS * 2u;
      const uint32_t swz_mask = pitch == 128 ? 7u : (pitch == 64 ? 3u : 1u);
      const uint32_t sub_bytes = 128u * pitch;
      uint8_t* staging = smem + staging_off + (uint32_t)gi * staging_bytes;
      const uint32_t staging_addr = smem_base + staging_off + (uint32_t)gi * staging_bytes;
      const float* vec = reinterpret_cast<const float*>(smem + vec_off);
      int acc = set; uint32_t acc_phase = 0;
      const int hi = row / p.twb, wq = row - hi * p.twb;
      const bool row_ok = wq < p.two && hi < p.th;
      const int mr = hi * p.two + wq;               // compacted staging row (valid rows only)
      const int rows_valid = p.th * p.two;
      // statistics: this lane's column pair of every sub-box over the rows its warp scans (tc_stats_scan)
      const int nsub = p.BN / p.CS;
      float sacc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.f; }
      int s_nb = -1;
      int sbuf = 0;                                 // staging buffer of the next tile (bf16, BN <= 64: double buffered)
      // Deterministic flush: every thread parks its partial sums, then one thread per (statistic, channel) adds the
      // partials of the 4 warps (and, for 64-byte rows, of the two lanes that share a column pair) in a fixed order.
      // Float atomics here would make BN statistics differ by ~1e-7 run to run, which bf16 rounding + the network
      // amplify into visibly different gradients (tools/diag_repeat2.py).
      auto flush_stats = [&]() {
        if (p.stat && s_nb >= 0) {
          // the partials are parked in this group's staging tile ([128 threads][nsub][4] floats <= 4 KB of its >= 8 KB):
          // a flush happens once per CTA (or per N tile), and a dedicated region cost 8-16 KB of shared memory that
          // some layers need for their fourth operand stage (and with it the second MMA issuer)
          float* park = reinterpret_cast<float*>(staging);
          if (et == 0) ptx::tma_store_wait_read();       // no store may still be reading the tile
          ptx::named_bar_sync(bar_id, 128);
#pragma unroll
          for (int sub = 0; sub < 4; ++sub)
            if (sub < nsub) {
              *reinterpret_cast<float4*>(park + (et * nsub + sub) * 4) = make_float4(sacc[sub][0], sacc[sub][1], sacc[sub][2], sacc[sub][3]);
              sacc[sub][0] = sacc[sub][1] = sacc[sub][2] = sacc[sub][3] = 0.f;
            }
          ptx::named_bar_sync(bar_id, 128);
          const int cpw = p.CS / 2, rpl = 64 / p.CS;          // column pairs per row, rows per load (see tc_stats_scan)
          for (int i = et; i < 2 * p.BN; i += 128) {
            const int which = i / p.BN, c = i - which * p.BN;
            const int sub = c / p.CS, cc = c - sub * p.CS;
            const int comp = (cc & 1) + 2 * which;
            float v = 0.f;
            for (int w = 0; w < 4; ++w)
              for (int ro = 0; ro < rpl; ++ro)
                v += park[((w * 32 + ro * cpw + (cc >> 1)) * nsub + sub) * 4 + comp];
            atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
          }
          ptx::named_bar_sync(bar_id, 128);
          // the statistics scan relies on never-written staging rows reading as zeros: put them back
#pragma unroll
          for (int sub = 0; sub < 4; ++sub)
            if (sub < nsub) *reinterpret_cast<float4*>(park + (et * nsub + sub) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          ptx::named_bar_sync(bar_id, 128);
        }
      };
      for (int st = blockIdx.x + set * (int)gridDim.x; st < total_super; st += S * (int)gridDim.x) {
        int nt, sp;
        p.fd_ntiles.divmod(st, sp, nt);
        const int nb = nt * p.BN;
        const int mt = sp * p.npair + grp;
        if constexpr (F32) {
          // ---- fp32 output (parity mode): 32-column chunks through two 16 KB staging tiles ----
          float* park = reinterpret_cast<float*>(smem + stat_off) + gi * (int)park_floats;      // [4 row groups][2][BN]
          if (p.stat && nb != s_nb) {
            if (s_nb >= 0) {
              ptx::named_bar_sync(bar_id, 128);
              for (int i = et; i < 2 * p.BN; i += 128) {
                const int which = i / p.BN, c = i - which * p.BN;
                float v = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) { v += park[(r * 2 + which) * p.BN + c]; park[(r * 2 + which) * p.BN + c] = 0.f; }
                atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
              }
              ptx::named_bar_sync(bar_id, 128);
            }
            s_nb = nb;
          }
          ptx::mbar_wait(t_full(acc), acc_phase);
          ptx::tc_fence_after();
          if (mt < m_tiles) {
            int w0, h0, n;
            decode_m(mt, w0, h0, n);
            const bool valid = row_ok && (w0 + wq) < p.W && (h0 + hi) < p.H;
            const long long pix = ((long long)n * p.H + (h0 + hi)) * p.W + (w0 + wq);
            const uint32_t t_base = tmem_base + (uint32_t)(acc * p.npair + grp) * acc_cols + ((uint32_t)(q * 32) << 16);
            for (int j = 0; j < p.BN / 32; ++j) {
              uint32_t v[32];
              ptx::tmem_ld32(t_base + (uint32_t)(j * 32), v);
              ptx::tmem_ld_wait();
              const int c0 = nb + j * 32;
              float f[32];
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(vec + c0 + k);
                f[k] = __uint_as_float(v[k]) + b4.x; f[k + 1] = __uint_as_float(v[k + 1]) + b4.y;
                f[k + 2] = __uint_as_float(v[k + 2]) + b4.z; f[k + 3] = __uint_as_float(v[k + 3]) + b4.w;
              }
              if (p.tf && valid) {
                const float4* tp4 = reinterpret_cast<const float4*>(p.tf + pix * p.t_ld + c0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float4 t4 = tp4[k];
                  const int c = 4 * k;
                  f[c] += vec[p.N + c0 + c] * t4.x + vec[2 * p.N + c0 + c];
                  f[c + 1] += vec[p.N + c0 + c + 1] * t4.y + vec[2 * p.N + c0 + c + 1];
                  f[c + 2] += vec[p.N + c0 + c + 2] * t4.z + vec[2 * p.N + c0 + c + 2];
                  f[c + 3] += vec[p.N + c0 + c + 3] * t4.w + vec[2 * p.N + c0 + c + 3];
                }
              }
              if (p.relu) {
#pragma unroll
                for (int k = 0; k < 32; ++k) f[k] = fmaxf(f[k], 0.f);
              }
              if (p.post) {
#pragma unroll
                for (int k = 0; k < 32; ++k) f[k] = fmaf(vec[p.N + c0 + k], f[k], vec[2 * p.N + c0 + k]);
              }
              uint8_t* stg = staging + (uint32_t)(j & 1) * 16384u;
              if (et == 0) ptx::tma_store_wait_read1();
              ptx::named_bar_sync(bar_id, 128);
              if (row_ok) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4)
                  *reinterpret_cast<float4*>(stg + (uint32_t)mr * 128u + (uint32_t)((g4 ^ (mr & 7)) << 4)) =
                      valid ? make_float4(f[4 * g4], f[4 * g4 + 1], f[4 * g4 + 2], f[4 * g4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
              ptx::fence_proxy_async_smem();
              ptx::named_bar_sync(bar_id, 128);
              if (et == 0) {
                ptx::tma_store_4d(&tmC, staging_addr + (uint32_t)(j & 1) * 16384u, c0, w0, h0, n);
                ptx::tma_store_commit();
              }
              if (p.stat) {
                const int c = et & 31, rgp = et >> 5;
                float a0 = 0.f, b0 = 0.f;
#pragma unroll 8
                for (int r = rgp * 32; r < rgp * 32 + 32; ++r) {
                  if (r < rows_valid) {
                    const float x = *reinterpret_cast<const float*>(stg + (uint32_t)r * 128u + (uint32_t)((((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4));
                    a0 += x; b0 = fmaf(x, x, b0);
                  }
                }
                park[(rgp * 2 + 0) * p.BN + j * 32 + c] += a0;
                park[(rgp * 2 + 1) * p.BN + j * 32 + c] += b0;
              }
            }
            // an odd chunk count would leave the buffer parity out of step with `j & 1`; BN is 32, 64 or 128 and a single
            // chunk (BN = 32) always uses buffer 0: drain before the next tile reuses it
            if (p.BN == 32) { if (et == 0) ptx::tma_store_wait_read(); ptx::named_bar_sync(bar_id, 128); }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(t_empty(acc));
          acc += S; if (acc >= NACC) { acc -= NACC; acc_phase ^= 1u; }
          continue;
        }
        if (p.stat && nb != s_nb) { flush_stats(); s_nb = nb; }
        if (et == 0) FU_DBG(2 + grp, (st - (int)blockIdx.x) / (int)gridDim.x, 0);
        uint8_t* stg = staging + (uint32_t)sbuf * tile_bytes;
        const uint32_t stg_addr = staging_addr + (uint32_t)sbuf * tile_bytes;
        if (nstg == 2 && mt < m_tiles) {
          // double-buffered staging: only the store issued two tiles ago must have finished reading this buffer
          if (et == 0) ptx::tma_store_wait_read1();
          ptx::named_bar_sync(bar_id, 128);
          sbuf ^= 1;
        }
        // (tile coordinates before the accumulator wait: the index arithmetic hides behind the MMAs)
        int w0 = 0, h0 = 0, n = 0;
        if (mt < m_tiles) decode_m(mt, w0, h0, n);
        const bool valid = row_ok && (w0 + wq) < p.W && (h0 + hi) < p.H;
        ptx::mbar_wait(t_full(acc), acc_phase);
        ptx::tc_fence_after();
        if (et == 0) FU_DBG(2 + grp, (st - (int)blockIdx.x) / (int)gridDim.x, 1);
        if (mt < m_tiles) {
          const long long pix = ((long long)n * p.H + (h0 + hi)) * p.W + (w0 + wq);
          const uint32_t t_base = tmem_base + (uint32_t)(acc * p.npair + grp) * acc_cols + ((uint32_t)(q * 32) << 16);
          // ReLU without a BatchNorm behind it is applied to the packed bf16 pairs (16 instead of 32 instructions per
          // chunk; rounding is monotonic and keeps zero, so relu(bf16(x)) == bf16(relu(x)))
          const bool relu_packed = p.relu && !p.post;
          for (int j = 0; j < p.BN / 32; ++j) {
            uint32_t v[32];
            ptx::tmem_ld32(t_base + (uint32_t)(j * 32), v);
            ptx::tmem_ld_wait();
            if constexpr (STK) {
              // out[m] = blk0[m] + blk1[m + 1] + blk2[m + 2]: box positions m + 1, m + 2 of a valid output column are lanes
              // of the same warp (twb divides 32); the lanes whose neighbours would be another warp's are halo columns
              uint32_t u[32];
              ptx::tmem_ld32(t_base + (uint32_t)(p.BN + j * 32), u);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k)
                v[k] = __float_as_uint(__uint_as_float(v[k]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[k]), 1));
              ptx::tmem_ld32(t_base + (uint32_t)(2 * p.BN + j * 32), u);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k)
                v[k] = __float_as_uint(__uint_as_float(v[k]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[k]), 2));
            }
            const int c0 = nb + j * 32;
            float f[32];
#pragma unroll
            for (int k = 0; k < 32; k += 4) {      // bias: two packed fp32 adds per four channels
              const float4 b4 = *reinterpret_cast<const float4*>(vec + c0 + k);
              const float2 r0 = fu_add2(make_float2(__uint_as_float(v[k]), __uint_as_float(v[k + 1])), make_float2(b4.x, b4.y));
              const float2 r1 = fu_add2(make_float2(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])), make_float2(b4.z, b4.w));
              f[k] = r0.x; f[k + 1] = r0.y; f[k + 2] = r1.x; f[k + 3] = r1.y;
            }
            if (S < 2 && p.t && valid) {          // (the two-set instantiation runs at its register cap and does not carry this path)
              const bf16* tp = p.t + pix * p.t_ld + c0;
#pragma unroll
              for (int k = 0; k < 32; k += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(tp + k);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const float2 tf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k2]));
                  const int c = k + 2 * k2;
                  f[c] += vec[p.N + c0 + c] * tf.x + vec[2 * p.N + c0 + c];
                  f[c + 1] += vec[p.N + c0 + c + 1] * tf.y + vec[2 * p.N + c0 + c + 1];
                }
              }
            }
            if (p.post) {        // eval-mode BatchNorm behind the ReLU: z = a * relu(conv) + b (unet.py:213-215)
              if (p.relu) {
#pragma unroll
                for (int k = 0; k < 32; ++k) f[k] = fmaxf(f[k], 0.f);
              }
#pragma unroll
              for (int k = 0; k < 32; ++k) f[k] = fmaf(vec[p.N + c0 + k], f[k], vec[2 * p.N + c0 + k]);
            }
            if (row_ok) {
              // rows that fall outside the image are written as zeros so the statistics can scan the tile
              const int colt = j * 32;
              const int sub = colt >> p.cs_shift;
              const uint32_t byte_in_row = (uint32_t)(colt & (p.CS - 1)) * 2u;
              const uint32_t row_off = (uint32_t)mr * pitch + byte_in_row;
              uint8_t* dst = stg + (uint32_t)sub * sub_bytes;
              const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                uint32_t pk[4];
                if (valid) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[g4 * 8 + 2 * k], f[g4 * 8 + 2 * k + 1]);
                    if (relu_packed) h2 = __hmax2(h2, zero2);
                    pk[k] = *reinterpret_cast<uint32_t*>(&h2);
                  }
                } else {
                  pk[0] = pk[1] = pk[2] = pk[3] = 0u;
                }
                const uint32_t logical = row_off + (uint32_t)g4 * 16u;
                const uint32_t phys = logical ^ (((logical >> 7) & swz_mask) << 4);
                *reinterpret_cast<uint4*>(dst + phys) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        // accumulator drained: hand the TMEM stage back to the MMA warp
        ptx::tc_fence_before();
        ptx::mbar_arrive(t_empty(acc));
        if (et == 0) FU_DBG(2 + grp, (st - (int)blockIdx.x) / (int)gridDim.x, 2);
        acc += S; if (acc >= NACC) { acc -= NACC; acc_phase ^= 1u; }
        if (mt < m_tiles) {
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (et == 0) {
            for (int s2 = 0; s2 < p.BN / p.CS; ++s2) {
              if (p.planeC) ptx::tma_store_5d(&tmC, stg_addr + (uint32_t)s2 * sub_bytes, 0, w0, h0, n, (nb + s2 * p.CS) / p.planeC);
              else ptx::tma_store_4d(&tmC, stg_addr + (uint32_t)s2 * sub_bytes, nb + s2 * p.CS, w0, h0, n);
            }
            ptx::tma_store_commit();
          }
          if (p.stat) {
            // per-channel sum / sum of squares of the stored (bf16) tile; rows no pixel maps to hold zeros
            if (pitch == 128) tc_stats_scan<128>(stg, sub_bytes, nsub, q, lane, sacc);
            else tc_stats_scan<64>(stg, sub_bytes, nsub, q, lane, sacc);
          }
          if (nstg == 1) {
            if (et == 0) ptx::tma_store_wait_read();   // staging may be overwritten after this
            ptx::named_bar_sync(bar_id, 128);
          }
          if (et == 0) FU_DBG(2 + grp, (st - (int)blockIdx.x) / (int)gridDim.x, 3);
        }
      }
      if constexpr (F32) {
        if (p.stat && s_nb >= 0) {
          float* park = reinterpret_cast<float*>(smem + stat_off) + gi * (int)park_floats;
          ptx::named_bar_sync(bar_id, 128);
          for (int i = et; i < 2 * p.BN; i += 128) {
            const int which = i / p.BN, c = i - which * p.BN;
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) v += park[(r * 2 + which) * p.BN + c];
            atomicAdd(p.stat + which * p.N + s_nb + c, (double)v);
          }
        }
      } else {
        flush_stats();
      }
      if (et == 0) ptx::tma_store_wait_all();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == mma_warp) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, tmem_cols); }
}

// ===========================================================================
// weight gradient: dW[tap][co][ci] += sum over pixels dY[p][co] * X[p @ tap][ci]
// A = dY tile (M = 128 out-channels), B = X tile shifted by the tap (N = 64/128 in-channels), K = 64 pixels
// per pipeline stage.  Both operands are the NHWC tiles exactly as TMA delivers them (pixel rows of 64
// channels) = MN-major for the MMA.  One CTA owns one (co tile, ci tile, filter row) and a range of pixel
// tiles (split-K); the 3 taps of the filter row share the dY tile and accumulate in 3 TMEM regions.
// ===========================================================================
// Epilogue of the weight-gradient kernels: thread = accumulator row (out-channel) `co`.  For every tap the row leaves
// TMEM through a private shared-memory row buffer (the operand stages are free once the last MMA has retired) and
// is added to the global [tap][co][ci] accumulator by ONE bulk reduction (`cp.reduce.async.bulk ... add.f32`, up to
// 512 B) instead of N/4 16-byte `red.global` instructions: the split-K epilogues cost 0.37 ms of a 7.5 ms step
// (FU_TC_WGRAD_NOEPI diagnostic) and most of that was issue and L2-atomic traffic.  Rows are private to their
// thread, so the only ordering needed is the thread's own proxy fence and bulk-group waits (double buffered).
// m64 != 0: the MMAs ran with M = 64 (layers with <= 64 out-channels; an M = 128 tile would be half or three quarters
// padding).  `m64` selects how accumulator row r maps to a TMEM lane: 2 (what the hardware does, tools/m64_probe.py):
// lane (r % 16) + 32 * (r / 16), i.e. 16 lanes of each warp's quadrant; 1 (wrong, kept for the probe): lane r.
// Measured gain is small (93 -> 78 us at 32->32 @192x192): these layers are not bound by tensor-pipe time.
__device__ __forceinline__ void tc_wgrad_epilogue(uint32_t tmem_base, uint32_t smem_base, uint32_t avail_bytes, int N, int ntaps,
                                                  int tap0, int tap_stride, int co0, int Cout, int ci0, int Cin, int row,
                                                  int q, float* dw_acc, int skip, int m64) {
  const uint32_t pitch = (uint32_t)N * 4u + 16u;            // +16 B: 8 consecutive rows cover all 32 banks
  const int nbuf = avail_bytes >= 2u * 128u * pitch ? 2 : 1;
  const int ncols = min(N, Cin - ci0);                      // valid in-channels of this tile
  int acc_row = row;                                        // accumulator row held by this thread's TMEM lane
  if (m64 == 1) acc_row = row < 64 ? row : -1;
  else if (m64 == 2) acc_row = (row & 31) < 16 ? (row >> 5) * 16 + (row & 15) : -1;
  else if (m64 == 3) {
    // filter rows stacked along M (tc_wgrad3_kernel, mstack): accumulator row = (2 - kh) * 32 + co, columns (kw, ci)
    acc_row = row < 96 ? (row & 31) : -1;
    tap0 = (2 - (row >> 5)) * 3;
  }
  const int co = co0 + acc_row;
  const bool live = acc_row >= 0 && co < Cout && ncols > 0 && !skip;
  for (int t = 0; t < ntaps; ++t) {
    const uint32_t soff = (uint32_t)(row * nbuf + (t % nbuf)) * pitch;
    if (t >= nbuf) { if (nbuf == 2) ptx::tma_store_wait_read1(); else ptx::tma_store_wait_read(); }
    for (int j = 0; j < N / 32; ++j) {
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + (uint32_t)(t * N + j * 32) + ((uint32_t)(q * 32) << 16), v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + soff + (uint32_t)(j * 32 + i) * 4u),
                     "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3]) : "memory");
    }
    if (live) {
      ptx::fence_proxy_async_smem();
      const int tap = tap0 + t * tap_stride;
      ptx::bulk_reduce_add_f32(dw_acc + ((long long)tap * Cout + co) * Cin + ci0, smem_base + soff, (uint32_t)ncols * 4u);
    }
    ptx::tma_store_commit();
  }
  ptx::tma_store_wait_all();       // every reduction has completed before the CTA (and its smem) goes away
}

struct TcWgradParams {
  int B, H, W, Cin, Cout;
  int ksz, pad, taps_per_cta, groups;
  int N;                        // in-channel tile (64 or 128)
  int tw, th, tn;               // pixel tile, tw*th*tn == 64
  int tiles_w, tiles_h, tiles_b;
  int co_tiles, ci_tiles, splits, stages;
  float* dw_acc;                // [taps][Cout][Cin] fp32, zeroed by the caller
  int skip_epi;                 // diagnostics only (FU_TC_WGRAD_NOEPI=1): drop the accumulators instead of adding them
  int m64;                      // 0: M = 128 MMAs; 1/2: M = 64 (out-channels <= 64), value = TMEM row layout (see epilogue)
  int b5;                       // B operand gathered with stride 2 (5-D map), taps = 2x2, pixel grid (W, H), B = 1
  int kpix;                     // pixels (K) per pipeline stage: 64, or 128 where three such stages fit shared memory (the
                                // kernel spends a roughly constant ~800-1000 cycles per stage on barrier round trips and TMA
                                // issue, which bounds the 1x1 / 2x2 layers: 4 MMAs per 64-pixel stage)
  // split-bf16 x3 parity mode: every pixel tile is visited three times, (dY_hi, X_hi), (dY_hi, X_lo), (dY_lo, X_hi);
  // y_lo / x_lo are the channel offsets of the lo halves in the two (twin) maps
  int split, y_lo, x_lo;
};

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw);
  // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so role-local scalars can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t kBox = (uint32_t)p.kpix * 128u;        // kpix pixels x 64 channels bf16
  const int nblk_b = p.N / 64;
  const uint32_t a_bytes = 2u * kBox;
  const uint32_t b_bytes = (uint32_t)(p.taps_per_cta * nblk_b) * kBox;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar_off = (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kTcMaxStages + s); };
  const uint32_t done_bar = bar_base + 8u * (uint32_t)(2 * kTcMaxStages);
  const uint32_t slot_addr = bar_base + 8u * (uint32_t)(2 * kTcMaxStages + 1);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * kTcMaxStages + 1));
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.taps_per_cta * p.N)) tmem_cols <<= 1;

  // work decode
  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int grp = bid % p.groups; bid /= p.groups;
  const int ci_t = bid % p.ci_tiles;
  const int co_t = bid / p.ci_tiles;
  const int co0 = co_t * 128, ci0 = ci_t * p.N;
  const int total_pt = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per = (total_pt + p.splits - 1) / p.splits;
  const int pt_begin = split * per;
  const int pt_end = min(total_pt, pt_begin + per);
  const int npass = p.split ? 3 : 1;
  const int n_iters = max(pt_end - pt_begin, 0) * npass;
  const int nblk_a = (p.Cout - co0 > 64) ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmY); ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1) { ptx::tmem_alloc(slot_addr, tmem_cols); ptx::tmem_relinquish(); }
  pdl_wait(); pdl_trigger();          // (see tc_conv_kernel)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    // whole warp loops, one elected lane issues the TMA loads
    int stage = 0; uint32_t phase = 0;
    // pixel-tile coordinates advanced incrementally (no integer divisions in the per-stage loop)
    int wi, hi, ni;
    { int pt = pt_begin; wi = pt % p.tiles_w; pt /= p.tiles_w; hi = pt % p.tiles_h; ni = pt / p.tiles_h; }
    int pass = 0;
    for (int it = 0; it < n_iters; ++it) {
      const int yoff = co0 + (pass == 2 ? p.y_lo : 0), xoff = ci0 + (pass == 1 ? p.x_lo : 0);
      const int w0 = wi * p.tw, h0 = hi * p.th, n0 = ni * p.tn;
      if (++pass == npass) {
        pass = 0;
        if (++wi == p.tiles_w) { wi = 0; if (++hi == p.tiles_h) { hi = 0; ++ni; } }
      }
      ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t a_dst = smem_base + (uint32_t)stage * stage_bytes;
        ptx::mbar_expect_tx(full_bar(stage), (uint32_t)nblk_a * kBox + b_bytes);
        for (int blk = 0; blk < nblk_a; ++blk)
          ptx::tma_load_4d(a_dst + (uint32_t)blk * kBox, &tmY, full_bar(stage), yoff + blk * 64, w0, h0, n0);
        for (int t = 0; t < p.taps_per_cta; ++t) {
          const int kh = p.ksz > 1 ? grp : 0, kw = p.ksz > 1 ? t : 0;
          for (int blk = 0; blk < nblk_b; ++blk) {
            const uint32_t dst = a_dst + a_bytes + (uint32_t)(t * nblk_b + blk) * kBox;
            if (p.b5) ptx::tma_load_5d(dst, &tmX, full_bar(stage), xoff + blk * 64, kw, w0, kh, h0);
            else ptx::tma_load_4d(dst, &tmX, full_bar(stage), xoff + blk * 64, w0 + kw - p.pad, h0 + kh - p.pad, n0);
          }
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    const uint32_t idesc = umma_idesc_bf16_mn((uint32_t)p.N, p.m64 ? 64u : 128u);
    const uint64_t dbase = umma_desc_mnmajor(0, kBox, 1024u);
    for (int it = 0; it < n_iters; ++it) {
      ptx::mbar_wait(full_bar(stage), phase);
      if (ptx::elect_one()) {
        const uint32_t a_addr = smem_base + (uint32_t)stage * stage_bytes;
        const uint64_t ad0 = umma_desc_at(dbase, a_addr);
        // kpix pixels = kpix / 16 K steps; 16 pixel rows = 2048 B = +128 in the address field.  K step outermost so
        // that consecutive MMAs accumulate into different taps' accumulators.
        const int ksteps = p.kpix >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          for (int t = 0; t < p.taps_per_cta; ++t) {
            const uint64_t bd0 = umma_desc_at(dbase, a_addr + a_bytes + (uint32_t)(t * nblk_b) * kBox);
            ptx::umma_bf16(tmem_base + (uint32_t)(t * p.N), ad0 + (uint64_t)(128 * ks), bd0 + (uint64_t)(128 * ks), idesc,
                           (it | ks) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(empty_bar(stage));
        if (it == n_iters - 1) ptx::umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    if (n_iters == 0 && ptx::elect_one()) ptx::umma_commit(done_bar);
    __syncwarp();
  } else if (n_iters > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    tc_wgrad_epilogue(tmem_base, smem_base, (uint32_t)p.stages * stage_bytes, p.N, p.taps_per_cta,
                      p.ksz > 1 ? p.ksz * grp : 0, 1, co0, p.Cout, ci0, p.Cin, row, q, p.dw_acc, p.skip_epi, p.m64);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, tmem_cols); }
}

// ===========================================================================
// weight gradient of the 3x3 convolutions, second generation (levels with W >= 32).
// K tiles are single image-row segments of tw pixels.  The dY tile (A, M = out-channels) is loaded once per
// stage and the X operand (B, N = in-channels) comes from ONE halo box of (tw+2) x nrows pixels; tap (kh,kw)
// is that tile read from row (kh-kh0)*(tw+2)+kw on (a start-address shift of the MN-major descriptor: the
// pixel rows are the K dimension, and a single image row never wraps, so no garbage enters the reduction).
// Thin layers (Cin = 32) use 32-channel / 64-byte boxes so all nine taps (9 x 32 columns) fit one CTA's
// TMEM; wider layers keep one filter row (3 taps) per CTA.
// Tap stacking (N == cb, i.e. Cin <= 64): the kw = 0,1,2 views of a halo row differ by one pixel row
// (= `brow` bytes) of the MN-major tile, so they are addressed as three consecutive N blocks of ONE operand
// whose leading-dimension byte offset is `brow`: one MMA with N = 3*Cin replaces three with N = Cin.  The
// accumulator columns come out as (kw, ci), exactly the per-tap layout the epilogue already reads.  The thin
// layers are bound by MMA count (every MMA re-reads the 128-row dY operand from shared memory), so this
// is a ~2x cut of their tensor-pipe time.
// ===========================================================================
struct TcWgrad3Params {
  int B, H, W, Cin, Cout;
  int tw, segs;                 // K pixels per stage (32/48/64), row segments per image row
  int N, cb;                    // in-channel tile (32/64/128) and channels per B box (32 -> SW64, 64 -> SW128)
  int tpg, groups;              // taps per CTA (9 or 3) and tap groups (1 or 3)
  int co_tiles, ci_tiles, splits, stages;
  unsigned b_stage_bytes;       // bytes of the B region of one stage
  int stack;                    // 1: the three kw taps of a filter row are ONE MMA with N = 3*N (see above)
  int mstack;                   // 1 (Cout <= 32, with stack): the three FILTER ROWS are stacked along M as well.  A K tile is the
                                //    row segment (h, w0..w0+tw) of X row h+1 against dY rows h, h+1, h+2 loaded as three
                                //    32-channel (SWIZZLE_64B) blocks one leading-dimension offset apart: accumulator rows
                                //    [32 s, 32 s + 32) hold sum dY[h+s] (x) X[h+1] = filter row kh = 2 - s, so ONE M = 128 MMA
                                //    per 16 pixels replaces three M = 64 ones (which run at half rate with half their rows
                                //    unused: the tensor pipe was 54 % busy doing 25 % useful work, profiles/r02_ncu_thin_wgrad3_before.txt).
                                //    Row bands run over h = -2 .. H-1; TMA zero-fills the rows outside the image.
  int skip_epi;                 // diagnostics only (FU_TC_WGRAD_NOEPI=1)
  int m64;                      // see TcWgradParams
  float* dw_acc;                // [9][Cout][Cin] fp32, zeroed by the caller
  int split, y_lo, x_lo;        // see TcWgradParams
};

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad3_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const TcWgrad3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw);
  // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so role-local scalars can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t kABox = (uint32_t)p.tw * 128u;            // one 64-channel box of tw pixels (a multiple of 1024: tw % 16 == 0)
  // A region of a stage: two 64-channel boxes, or (mstack) four 32-channel blocks; tw <= 128 pixels
  const uint32_t a_bytes = p.mstack ? 4u * (uint32_t)p.tw * 64u : 2u * kABox;
  const uint32_t brow = (uint32_t)p.cb * 2u;                // bytes per pixel row of a B box
  const int nrows = p.tpg == 9 ? 3 : 1;                     // halo rows held per stage
  const int nblk_b = p.N / p.cb;
  const uint32_t b_box_bytes = ((uint32_t)((p.tw + 2) * nrows) * brow + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + p.b_stage_bytes;
  const uint32_t bar_off = (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kTcMaxStages + s); };
  const uint32_t done_bar = bar_base + 8u * (uint32_t)(2 * kTcMaxStages);
  const uint32_t slot_addr = bar_base + 8u * (uint32_t)(2 * kTcMaxStages + 1);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * kTcMaxStages + 1));
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.tpg * p.N)) tmem_cols <<= 1;

  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int grp = bid % p.groups; bid /= p.groups;
  const int ci_t = bid % p.ci_tiles;
  const int co_t = bid / p.ci_tiles;
  const int co0 = co_t * 128, ci0 = ci_t * p.N;
  const int Hk = p.mstack ? p.H + 2 : p.H;                    // row bands per image
  const int total_kt = p.B * Hk * p.segs;
  const int per = (total_kt + p.splits - 1) / p.splits;
  const int kt_begin = split * per;
  const int kt_end = min(total_kt, kt_begin + per);
  const int npass = p.split ? 3 : 1;
  const int n_iters = max(kt_end - kt_begin, 0) * npass;
  const int nblk_a = (p.Cout - co0 > 64) ? 2 : 1;
  const int kh0 = p.mstack ? 2 : (p.tpg == 9 ? 0 : grp);    // first filter row held by this CTA (mstack: X row h + 1)

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmY); ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1) { ptx::tmem_alloc(slot_addr, tmem_cols); ptx::tmem_relinquish(); }
  pdl_wait(); pdl_trigger();          // (see tc_conv_kernel)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    const uint32_t a_tx = p.mstack ? 3u * (uint32_t)p.tw * 64u : (uint32_t)nblk_a * (uint32_t)p.tw * 128u;
    const uint32_t b_tx = (uint32_t)nblk_b * (uint32_t)((p.tw + 2) * nrows) * brow;
    // (segment, row band, image) of the K tile, advanced incrementally: three integer divisions per stage were most of
    //  what this warp executed, and with 2-4 MMAs per stage the producer's issue rate bounds the thin layers
    int seg_i, hb_i, n_i;
    { int kt = kt_begin; seg_i = kt % p.segs; kt /= p.segs; hb_i = kt % Hk; n_i = kt / Hk; }
    int pass = 0;
    for (int it = 0; it < n_iters; ++it) {
      const int yoff = co0 + (pass == 2 ? p.y_lo : 0), xoff = ci0 + (pass == 1 ? p.x_lo : 0);
      const int w0 = seg_i * p.tw;
      const int h = hb_i - (p.mstack ? 2 : 0);
      const int n = n_i;
      if (++pass == npass) {
        pass = 0;
        if (++seg_i == p.segs) { seg_i = 0; if (++hb_i == Hk) { hb_i = 0; ++n_i; } }
      }
      ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t a_dst = smem_base + (uint32_t)stage * stage_bytes;
        ptx::mbar_expect_tx(full_bar(stage), a_tx + b_tx);
        if (p.mstack) {
          for (int sft = 0; sft < 3; ++sft)       // dY rows h, h+1, h+2 -> M blocks 0, 1, 2 (block 3 is never read back)
            ptx::tma_load_4d(a_dst + (uint32_t)sft * (uint32_t)p.tw * 64u, &tmY, full_bar(stage), yoff, w0, h + sft, n);
        } else
        for (int blk = 0; blk < nblk_a; ++blk)
          ptx::tma_load_4d(a_dst + (uint32_t)blk * kABox, &tmY, full_bar(stage), yoff + blk * 64, w0, h, n);
        for (int blk = 0; blk < nblk_b; ++blk)
          ptx::tma_load_4d(a_dst + a_bytes + (uint32_t)blk * b_box_bytes, &tmX, full_bar(stage), xoff + blk * p.cb, w0 - 1,
                           h - 1 + kh0, n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    const uint32_t idesc = umma_idesc_bf16_mn((uint32_t)p.N, p.m64 ? 64u : 128u);
    const uint64_t da = umma_desc_mnmajor(0, kABox, 1024u);                       // A: 64-channel SW128 blocks
    uint64_t db;                                                                   // B: SW128 or SW64 blocks
    {
      uint64_t d = 0;
      d |= (uint64_t)((b_box_bytes >> 4) & 0x3FFF) << 16;                          // LBO: next channel block
      d |= (uint64_t)(((8u * brow) >> 4) & 0x3FFF) << 32;                          // SBO: next 8 pixel rows
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(p.cb == 64 ? 2 : 4) << 61;
      db = d;
    }
    // stacked operand: N block b = the tile shifted by b pixel rows
    const uint64_t dbs = (db & ~((uint64_t)0x3FFF << 16)) | ((uint64_t)((brow >> 4) & 0x3FFF) << 16);
    const uint32_t idesc_s = umma_idesc_bf16_mn((uint32_t)(3 * p.N), p.m64 ? 64u : 128u);
    const int ksteps = p.tw / 16;
    const int tpg = p.tpg, Nn = p.N;
    const uint32_t step_a16 = (16u * 128u) >> 4, step_b16 = (16u * brow) >> 4;
    // mstack: A = four 32-channel SWIZZLE_64B blocks (dY rows h, h+1, h+2, unused) tw * 64 bytes apart, M = 128
    uint64_t dam;
    {
      uint64_t d = 0;
      d |= (uint64_t)((((uint32_t)p.tw * 64u) >> 4) & 0x3FFF) << 16;               // LBO: next M block (next dY row)
      d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;                                 // SBO: next 8 pixel rows
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)4 << 61;                                                      // SWIZZLE_64B
      dam = d;
    }
    const uint32_t idesc_ms = umma_idesc_bf16_mn((uint32_t)(3 * p.N), 128u);
    const uint32_t step_am16 = (16u * 64u) >> 4;
    uint32_t tapoff16[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int kh = tpg == 9 ? t / 3 : 0, kw = tpg == 9 ? t % 3 : t;
      tapoff16[t] = ((uint32_t)(kh * (p.tw + 2) + kw) * brow) >> 4;
    }
    for (int it = 0; it < n_iters; ++it) {
      ptx::mbar_wait(full_bar(stage), phase);
      if (ptx::elect_one()) {
        const uint32_t a_addr = smem_base + (uint32_t)stage * stage_bytes;
        const uint64_t ad0 = da + (uint64_t)(a_addr >> 4);
        const uint64_t bd0 = db + (uint64_t)((a_addr + a_bytes) >> 4);
        const uint64_t bd0s = dbs + (uint64_t)((a_addr + a_bytes) >> 4);
        const uint32_t accum = it != 0 ? 1u : 0u;
        // K step outermost: consecutive MMAs go to DIFFERENT accumulators, so a dependent accumulation is
        // never issued back to back
        if (p.mstack) {
          const uint64_t adm = dam + (uint64_t)(a_addr >> 4);
          for (int ks = 0; ks < ksteps; ++ks)
            ptx::umma_bf16(tmem_base, adm + (uint64_t)(ks * step_am16), bd0s + (uint64_t)(ks * step_b16), idesc_ms, accum | (uint32_t)ks);
        } else if (p.stack) {
          for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              if (3 * r < tpg)
                ptx::umma_bf16(tmem_base + (uint32_t)(3 * r * Nn), ad0 + (uint64_t)(ks * step_a16),
                               bd0s + (uint64_t)(tapoff16[3 * r] + ks * step_b16), idesc_s, accum | (uint32_t)ks);
            }
          }
        } else {
          for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < tpg)
                ptx::umma_bf16(tmem_base + (uint32_t)(t * Nn), ad0 + (uint64_t)(ks * step_a16),
                               bd0 + (uint64_t)(tapoff16[t] + ks * step_b16), idesc, accum | (uint32_t)ks);
            }
          }
        }
        ptx::umma_commit(empty_bar(stage));
        if (it == n_iters - 1) ptx::umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (n_iters > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    tc_wgrad_epilogue(tmem_base, smem_base, (uint32_t)p.stages * stage_bytes, p.N, p.tpg, p.tpg == 9 ? 0 : grp * 3, 1,
                      co0, p.Cout, ci0, p.Cin, row, q, p.dw_acc, p.skip_epi, p.mstack ? 3 : p.m64);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------
// Split-bf16 twin of an fp32 NHWC tensor (parity mode): hi = bf16(v) at dst[pix*dld + c], lo = bf16(v - hi) at
// dst[pix*dld + dlo + c].  The twin of a buffer with fp32 pixel stride ld has dld = 2*ld and dlo = ld, i.e. it
// occupies the same bytes per pixel, and a channel slice of the buffer maps to the same slice of either half.
// hi + lo carries 16 mantissa bits of v; the three products hi*hi' + hi*lo' + lo*hi' of two split operands are
// exact in the tensor core's fp32 accumulation, so a GEMM over twins differs from the fp32 one by ~2^-16 per term.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ src, int ld, int C, bf16* __restrict__ dst,
                                                        int dld, int dlo, long long P) {
  pdl_wait(); pdl_trigger();
  const int cv = C >> 2;                       // 4-channel vectors per pixel
  const long long total = P * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c = (int)(i - pix * cv) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + pix * ld + c);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    uint2 hh, ll;
    hh.x = *reinterpret_cast<const uint32_t*>(&h0); hh.y = *reinterpret_cast<const uint32_t*>(&h1);
    ll.x = *reinterpret_cast<const uint32_t*>(&l0); ll.y = *reinterpret_cast<const uint32_t*>(&l1);
    bf16* d = dst + pix * dld + c;
    *reinterpret_cast<uint2*>(d) = hh;
    *reinterpret_cast<uint2*>(d + dlo) = ll;
  }
}
inline int tc_split(const float* src, int ld, int C, bf16* dst, int dld, int dlo, long long P, int sms, cudaStream_t stream,
                    fu_counters* cnt) {
  long long blocks = (P * (C / 4) + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (blocks < 1) blocks = 1;
  fu_launch(split_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, fu_pdl_enabled(), src, ld, C, dst, dld, dlo, P);
  if (cnt) cnt->kernel_launches++;
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------
// Weight repacking, batched: ONE launch per step packs every tensor-core layer's fp32 torch weights into the
// two bf16 GEMM layouts, and ONE launch at the end of backward turns every [taps][M][N] fp32 weight-gradient
// accumulator into the torch layout.  (The per-layer versions were 92 launches and 0.64 ms per step, most of
// it uncoalesced 2-byte scatter: the transposes now go through shared-memory tiles.)
// ---------------------------------------------------------------------------
// One pack job = one fp32 torch weight viewed as [R][Cc][taps] (rows, columns, taps; taps fastest) and its two bf16 GEMM
// layouts, both written through 32 x 32 x taps shared-memory tiles:
//   out1[r*a1 + t*b1 + c]          -- for fixed (r, t) the columns are contiguous  ("row-major" output)
//   out2[base2 + c*a2 + t*b2 + r]  -- for fixed (c, t) the rows are contiguous     ("transposed" output)
// Conv2d 3x3/1x1 (R=Cout, Cc=Cin): out1 = forward [Cout][tap][Cin], out2 = data gradient [Cin][flip(tap)][Cout].
// Conv2d 2x2/s2  (R=Cout, Cc=Cin): out1 = forward (gather conv) [Cout][ab][Cin], out2 = data gradient (scatter GEMM) [(ab,ci)][Cout].
// ConvT  2x2/s2  (R=Cin, Cc=Cout): out1 = data gradient (gather conv over dY) [Cin][ab][Cout], out2 = forward (scatter GEMM) [(ab,co)][Cin].
struct TcPackJob {
  int block0;                   // first block of this job in the batched grid
  int nblocks;
  const float* w; bf16* out1; bf16* out2;
  int R, Cc, taps;
  long long a1, b1, a2, b2, base2;
  long long lo1, lo2;           // split-bf16 parity mode (0: off): the residual bf16(w - hi) is written lo1 / lo2 elements
                                // after the hi value (the GEMM K dimension of both layouts is [hi | lo])
};
struct TcUnpackJob {
  const float* acc;             // [taps][M][N] fp32
  long long dw_off;             // torch layout (M, N, k, k) at base + dw_off floats; the base (the step's flat gradient
                                // buffer) is a launch argument, so the table does not change when PyTorch hands the
                                // engine a different buffer
  int M, N, taps;
  int block0, nblocks;
};
template <typename Job>
__device__ __forceinline__ int tc_find_job(const Job* jobs, int njobs, int* s_job) {
  // last job with block0 <= blockIdx.x: every thread tests one job (one round trip to the table instead of the
  // six dependent ones of a single-thread binary search at the start of every block)
  if (threadIdx.x == 0) *s_job = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < njobs; i += blockDim.x)
    if (jobs[i].block0 <= (int)blockIdx.x && (i + 1 == njobs || jobs[i + 1].block0 > (int)blockIdx.x)) *s_job = i;
  __syncthreads();
  return *s_job;
}

// one 32 (rows) x 32 (columns) x TAPS tile; the loops are warp = row, lane = column, so no division by a run-time value
template <int TAPS>
__device__ __forceinline__ void tc_pack_tile(const TcPackJob& J, int lb, float* tile) {
  const int tiles_c = J.Cc / 32;
  const int r0 = (lb / tiles_c) * 32, c0 = (lb % tiles_c) * 32;
  constexpr int ROW = 32 * TAPS;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int r = wrp; r < 32; r += 8) {
    const float* src = J.w + ((long long)(r0 + r) * J.Cc + c0) * TAPS;
#pragma unroll
    for (int c = lane; c < ROW; c += 32) tile[r * 289 + c] = src[c];
  }
  __syncthreads();
  for (int r = wrp; r < 32; r += 8) {           // out1: (t, c) per row, c contiguous
    bf16* dst = J.out1 + (long long)(r0 + r) * J.a1 + c0 + lane;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      const float v = tile[r * 289 + lane * TAPS + t];
      const bf16 h = __float2bfloat16_rn(v);
      dst[(long long)t * J.b1] = h;
      if (J.lo1) dst[(long long)t * J.b1 + J.lo1] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
  for (int c = wrp; c < 32; c += 8) {           // out2: (t, r) per column, r contiguous
    bf16* dst = J.out2 + J.base2 + (long long)(c0 + c) * J.a2 + r0 + lane;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      const float v = tile[lane * 289 + c * TAPS + t];
      const bf16 h = __float2bfloat16_rn(v);
      dst[(long long)t * J.b2] = h;
      if (J.lo2) dst[(long long)t * J.b2 + J.lo2] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

__global__ void __launch_bounds__(256) tc_pack_batched_kernel(const TcPackJob* jobs, int njobs) {
  pdl_wait(); pdl_trigger();
  __shared__ int s_job;
  __shared__ float tile[32 * 289];      // [r][c*taps + t], odd row stride: conflict-free row- and column-wise
  const TcPackJob& J = jobs[tc_find_job(jobs, njobs, &s_job)];
  const int lb = (int)blockIdx.x - J.block0;
  if (J.taps == 9) tc_pack_tile<9>(J, lb, tile);
  else if (J.taps == 4) tc_pack_tile<4>(J, lb, tile);
  else tc_pack_tile<1>(J, lb, tile);
}

// dw[(m*N + n)*taps + t] = acc[(t*M + m)*N + n]: tiles of 8 m x 32 n, reads and writes both contiguous.
// READ AND CLEAR: every accumulator element is zeroed right after it is read, so the accumulators are clean for the next
// step's split-K reductions and the engine does not memset them (136 MB at the head of every backward, on the critical
// path) -- the zeros are written here instead, on the side stream beside the data-gradient chain.
__global__ void __launch_bounds__(256) tc_unpack_batched_kernel(const TcUnpackJob* jobs, int njobs, float* base) {
  pdl_wait(); pdl_trigger();
  __shared__ int s_job;
  __shared__ float tile[8][32 * 9 + 1];
  const TcUnpackJob& J = jobs[tc_find_job(jobs, njobs, &s_job)];
  const int lb = (int)blockIdx.x - J.block0;
  const int taps = J.taps, tiles_n = J.N / 32;
  const int m0 = (lb / tiles_n) * 8, n0 = (lb % tiles_n) * 32;
  const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;      // warp = one m row, lane = n column
  if (m0 + m < J.M) {
    float* src = const_cast<float*>(J.acc) + ((long long)(m0 + m)) * J.N + n0 + lane;
    for (int t = 0; t < taps; ++t) { tile[m][lane * taps + t] = src[(long long)t * J.M * J.N]; src[(long long)t * J.M * J.N] = 0.f; }
  }
  __syncwarp();
  if (m0 + m < J.M) {
    float* dst = base + J.dw_off + ((long long)(m0 + m) * J.N + n0) * taps;
    for (int c = lane; c < 32 * taps; c += 32) dst[c] = tile[m][c];
  }
}

// Jobs are collected by the engine over a step (tc_batch() != nullptr) and flushed in one launch; without a
// collector (kernel test hook) every job is flushed immediately through a one-entry table.
struct TcBatch {
  std::vector<TcPackJob> pack;
  std::vector<TcUnpackJob> unpack;
  float* flat = nullptr;        // base of the gradient buffer the unpack offsets are relative to
};
inline TcBatch*& tc_batch() {
  static thread_local TcBatch* b = nullptr;
  return b;
}
// upload `jobs` when they differ from what the device table already holds, then launch.  `pinned` (optional,
// dev_cap entries of page-locked host memory) is the staging copy the upload reads: a copy from pageable memory
// cannot be captured into a CUDA graph, and a captured copy re-reads its host source at every replay.
template <typename Job, typename Launch>
inline int tc_flush_jobs(std::vector<Job>& jobs, std::vector<Job>& uploaded, Job* dev_tbl, int dev_cap, Launch launch,
                         cudaStream_t stream, fu_counters* cnt, Job* pinned = nullptr) {
  if (jobs.empty()) return 0;
  if ((int)jobs.size() > dev_cap) return -1;
  int blocks = 0;
  for (auto& j : jobs) { j.block0 = blocks; blocks += j.nblocks; }
  if (uploaded.size() != jobs.size() || memcmp(uploaded.data(), jobs.data(), jobs.size() * sizeof(Job)) != 0) {
    const void* src = jobs.data();
    if (pinned) {
      // an earlier upload may still be reading the staging copy (eager mode; tables change only when a tensor
      // address changes, so this wait is rare).  While capturing nothing from eager mode is in flight on this
      // stream and synchronising would invalidate the capture.
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
      if (cap == cudaStreamCaptureStatusNone && cudaStreamSynchronize(stream) != cudaSuccess) return -1;
      memcpy(pinned, jobs.data(), jobs.size() * sizeof(Job));
      src = pinned;
    }
    if (cudaMemcpyAsync(dev_tbl, src, jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -1;
    uploaded = jobs;
  }
  launch(blocks, dev_tbl, (int)jobs.size());
  if (cnt) cnt->kernel_launches++;
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}
inline auto tc_pack_launcher(cudaStream_t stream) {
  return [stream](int blocks, const TcPackJob* tbl, int n) { fu_launch(tc_pack_batched_kernel, dim3(blocks), dim3(256), 0, stream, fu_pdl_enabled(), tbl, n); };
}
inline auto tc_unpack_launcher(cudaStream_t stream, float* base) {
  return [stream, base](int blocks, const TcUnpackJob* tbl, int n) { fu_launch(tc_unpack_batched_kernel, dim3(blocks), dim3(256), 0, stream, fu_pdl_enabled(), tbl, n, base); };
}
template <typename Job, typename Launch>
inline int tc_run_job_now(Job job, Launch launch, cudaStream_t stream, fu_counters* cnt) {
  Job* d = nullptr;
  if (cudaMalloc(&d, sizeof(Job)) != cudaSuccess) return -1;
  std::vector<Job> jobs(1, job), up;
  const int rc = tc_flush_jobs(jobs, up, d, 1, launch, stream, cnt);
  cudaStreamSynchronize(stream);
  cudaFree(d);
  return rc;
}
inline int tc_unpack(const float* acc, float* dw, int M, int N, int taps, cudaStream_t stream, fu_counters* cnt) {
  TcUnpackJob j;
  memset(&j, 0, sizeof(j));
  j.acc = acc; j.M = M; j.N = N; j.taps = taps;
  j.nblocks = ((M + 7) / 8) * (N / 32);
  if (tc_batch()) { j.dw_off = dw - tc_batch()->flat; tc_batch()->unpack.push_back(j); return 0; }
  j.dw_off = 0;
  return tc_run_job_now(j, tc_unpack_launcher(stream, dw), stream, cnt);
}

// ===========================================================================
// host side
// ===========================================================================
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a flag per (call site, device), so an
// engine created on a second GPU of the same process also gets the opt-in
inline bool tc_attr_needed(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

inline std::string& tc_err() {
  static thread_local std::string s;
  return s;
}
inline const char* tc_last_error() { return tc_err().c_str(); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

inline CUtensorMapSwizzle tc_swizzle_for(int inner_bytes) {
  return inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 tensor map over up to 4 dims; dims[0] is the contiguous (channel) dim, strides in elements
inline int tc_make_map(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides_elems,
                       const int* box, int inner_bytes, bool f32 = false) {
  PFN_encodeTiled fn = tc_encode_fn();
  if (!fn) { tc_err() = "cuTensorMapEncodeTiled entry point not available"; return -1; }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = (cuuint64_t)strides_elems[i] * (f32 ? 4ull : 2ull);
  }
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle_for(inner_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims %lld,%lld,%lld,%lld box %d,%d,%d,%d", (int)r,
             rank, dims[0], dims[1], rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0, box[0], box[1], rank > 2 ? box[2] : 0,
             rank > 3 ? box[3] : 0);
    tc_err() = buf;
    return -1;
  }
  return 0;
}

// pixel tile (tw, th, tn), tw*th*tn <= 128, minimising the number of tiles
inline void tc_pick_tile(int B, int H, int W, int& tw, int& th, int& tn) {
  long long best = -1;
  tw = 1; th = 1; tn = 1;
  for (int a = 1; a <= 128 && a <= W; ++a) {
    for (int b = 1; a * b <= 128 && b <= H; ++b) {
      int c = 128 / (a * b);
      if (c > B) c = B;
      if (c < 1) continue;
      const long long tiles = (long long)((W + a - 1) / a) * ((H + b - 1) / b) * ((B + c - 1) / c);
      const long long key = tiles * 1024 - a;   // fewer tiles first, then longer rows
      if (best < 0 || key < best) { best = key; tw = a; th = b; tn = c; }
    }
  }
}

struct TcConv {
  bool enabled = false;
  bool split = false;        // split-bf16 x3 parity mode: fp32 tensors, operands read from their [hi | lo] bf16 twins
                             // (pointer = hi half, pixel stride = 2 * fp32 stride, lo half = stride / 2 elements on)
  int Cin = 0, Cout = 0, k = 0;
  int kind = 0;              // 0: Conv2d 3x3/1x1 stride 1; 1: Conv2d 2x2 stride 2; 2: ConvTranspose2d 2x2 stride 2
  bf16* w_fwd = nullptr;     // [Cout][taps][Cin]
  bf16* w_dgrad = nullptr;   // [Cin][taps][Cout]
  float* dw_acc = nullptr;   // [taps][Cout][Cin] fp32 weight-gradient accumulator
  struct WCached {
    const void *x, *dy; int x_ld, dy_ld, B, H, W;
    CUtensorMap y, xm; TcWgradParams p; int grid; size_t smem;
  };
  std::vector<WCached> wcache;
  struct W3Cached {
    const void *x, *dy; int x_ld, dy_ld, B, H, W;
    CUtensorMap y, xm; TcWgrad3Params p; int grid; size_t smem;
  };
  std::vector<W3Cached> w3cache;
  struct Cached3 {
    const void *x, *y, *x2; int x_ld, y_ld, x2_ld, B, H, W, dir;     // x2: second source of the fused residual dgrad (or null)
    CUtensorMap a, b, c, a2, b2; TcConv3Params p; int grid; size_t smem; int S; int f32;
  };
  std::vector<Cached3> cache3;
  struct Cached {
    const void *x, *y; int x_ld, y_ld, B, H, W, dir;   // H, W: spatial dims of the GEMM's pixel grid
    CUtensorMap a, b, c; TcConvParams p; int grid; size_t smem; int G; int f32;
    // second-epilogue-operand map (t_smem), built at launch for the tensor the launch names
    CUtensorMap tm_t; const void* tm_t_ptr; int tm_t_ld; size_t smem_base_bytes; int stages_base;
  };
  std::vector<Cached> cache;
};

inline int tc_bn_max() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("FU_TC_BN_MAX");
    v = s ? atoi(s) : 128;
    if (v != 32 && v != 64 && v != 128 && v != 256) v = 128;
  }
  return v;
}

inline void tc_carve(TcConv& t, int Cin, int Cout, int k, bool transposed, bool enabled, Bump& w, Bump& ws, bool split = false) {
  t.Cin = Cin; t.Cout = Cout; t.k = k; t.split = split;
  t.kind = transposed ? 2 : (k == 2 ? 1 : 0);
  t.enabled = enabled && (k == 3 || k == 1 || k == 2) && (Cin % 32 == 0) && (Cout % 32 == 0) &&
              getenv("FU_TC_DISABLE") == nullptr;
  if (t.kind != 0 && getenv("FU_TC_NO_UPDOWN") != nullptr) t.enabled = false;
  if (!t.enabled) return;
  const size_t n = (size_t)Cin * Cout * k * k;
  t.w_fwd = w.take<bf16>(split ? 2 * n : n);
  t.w_dgrad = w.take<bf16>(split ? 2 * n : n);
  t.dw_acc = ws.take<float>(n);
  t.cache.clear();
  t.cache3.clear();
  t.wcache.clear();
  t.w3cache.clear();
}

inline int tc_pack(TcConv& t, const float* w, cudaStream_t stream, fu_counters* cnt) {
  if (!t.enabled) return 0;
  TcPackJob j;
  memset(&j, 0, sizeof(j));
  j.w = w;
  // K widths of the two GEMMs: doubled to [hi | lo] in split mode
  const long long s = t.split ? 2 : 1;
  const long long Ci = t.Cin * s, Co = t.Cout * s;
  if (t.kind == 0) {            // W[co][ci][tap]
    const int taps = t.k * t.k;
    j.R = t.Cout; j.Cc = t.Cin; j.taps = taps;
    j.out1 = t.w_fwd; j.a1 = taps * Ci; j.b1 = Ci;                                   // [co][tap][ci]
    j.out2 = t.w_dgrad; j.a2 = taps * Co; j.b2 = -Co; j.base2 = (taps - 1) * Co;     // [ci][flip tap][co]
    if (t.split) { j.lo1 = t.Cin; j.lo2 = t.Cout; }
  } else if (t.kind == 1) {     // W[co][ci][ab]
    j.R = t.Cout; j.Cc = t.Cin; j.taps = 4;
    j.out1 = t.w_fwd; j.a1 = 4 * Ci; j.b1 = Ci;                                      // gather conv: [co][ab][ci]
    j.out2 = t.w_dgrad; j.a2 = Co; j.b2 = (long long)t.Cin * Co; j.base2 = 0;        // scatter GEMM: [(ab,ci)][co]
    if (t.split) { j.lo1 = t.Cin; j.lo2 = t.Cout; }
  } else {                      // W[ci][co][ab]
    j.R = t.Cin; j.Cc = t.Cout; j.taps = 4;
    j.out1 = t.w_dgrad; j.a1 = 4 * Co; j.b1 = Co;                                    // gather conv over dY: [ci][ab][co]
    j.out2 = t.w_fwd; j.a2 = Ci; j.b2 = (long long)t.Cout * Ci; j.base2 = 0;         // scatter GEMM: [(ab,co)][ci]
    if (t.split) { j.lo1 = t.Cout; j.lo2 = t.Cin; }
  }
  j.nblocks = (j.R / 32) * (j.Cc / 32);
  if (tc_batch()) { tc_batch()->pack.push_back(j); return 0; }
  return tc_run_job_now(j, tc_pack_launcher(stream), stream, cnt);
}

inline int tc_env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// Split-K factor of the weight-gradient kernels.  One CTA is resident per SM, so CTAs = units x splits should
// fill whole waves: among 1..4 target waves (starting at `first`) take the one with the best last-wave fill,
// split counts rounded DOWN (rounding up made 3 units x 99 splits = 297 CTAs = three waves on 148 SMs).
inline long long tc_pick_splits(long long units, long long max_splits, int sms, int first) {
  long long best = 1; double best_eff = -1.0;
  for (int w = first; w <= 4; ++w) {
    long long sp = (long long)w * sms / units;
    if (sp > max_splits) sp = max_splits;
    if (sp < 1) sp = 1;
    const long long ctas = units * sp;
    const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
    if (eff > best_eff + 0.03) { best_eff = eff; best = sp; }
  }
  return best;
}

// epilogue groups of the conv kernel: 4 for the thin, shallow tiles (few MMAs per tile: the epilogue is the
// bottleneck), 2 up to 128 columns, 1 for 256-column tiles (TMEM holds 2G accumulators of BN columns)
inline int tc_pick_groups(int BN, int k_iters, int ksteps) {
  const int g_env = tc_env_int("FU_TC_EPI_GROUPS", 0);
  // (four groups only for 32-column tiles: at 64 columns their staging + `t` tiles leave 1-2 operand stages --
  //  residual 1x1 128->64 @96x96: 48 us with four groups, 33 with two)
  int G = (BN <= 32 && k_iters * ksteps <= 48) ? 4 : (BN <= 128 ? 2 : 1);
  if (g_env == 1 || g_env == 2 || g_env == 4) G = g_env;
  while (2 * G * BN > 512) G >>= 1;
  return G < 1 ? 1 : G;
}

inline bool tc_ptr_ok(const void* p, int ld) { return p && (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 8 == 0); }
// fp32 tensor (output / second epilogue operand of the split mode)
inline bool tc_ptr_ok_f32(const void* p, int ld) { return p && (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0); }
// set the second epilogue operand of a launch description for either storage mode
template <typename P>
inline void tc_set_t(const TcConv& t, P& p, const void* tp, int t_ld) {
  p.t = t.split ? nullptr : reinterpret_cast<const bf16*>(tp);
  p.tf = t.split ? reinterpret_cast<const float*>(tp) : nullptr;
  p.t_ld = t_ld;
}

// Build (or fetch) the launch description: dir 0 = forward (A has Cin channels, N = Cout), dir 1 = data
// gradient (A has Cout channels, N = Cin).
inline TcConv::Cached* tc_prepare(TcConv& t, int dir, const void* x, int x_ld, void* y, int y_ld, int B, int H, int W) {
  for (auto& c : t.cache)
    if (c.x == x && c.y == y && c.x_ld == x_ld && c.y_ld == y_ld && c.B == B && c.H == H && c.W == W && c.dir == dir)
      return &c;
  TcConv::Cached c;
  memset(&c, 0, sizeof(c));
  c.x = x; c.y = y; c.x_ld = x_ld; c.y_ld = y_ld; c.B = B; c.H = H; c.W = W; c.dir = dir; c.f32 = t.split ? 1 : 0;
  TcConvParams& p = c.p;
  const int K = dir == 0 ? t.Cin : t.Cout;
  const int N = dir == 0 ? t.Cout : t.Cin;
  p.B = B; p.H = H; p.W = W; p.Cin = K; p.N = N; p.ksz = t.k; p.pad = t.k / 2;
  p.KC = (K % 64 == 0) ? 64 : 32;
  { const int kc = tc_env_int("FU_TC_KC", 0); if ((kc == 16 || kc == 32 || kc == 64) && K % kc == 0) p.KC = kc; }
  tc_pick_tile(B, H, W, p.tw, p.th, p.tn);
  p.tiles_w = (W + p.tw - 1) / p.tw; p.tiles_h = (H + p.th - 1) / p.th; p.tiles_b = (B + p.tn - 1) / p.tn;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int bn = tc_bn_max();
  while (N % bn) bn >>= 1;
  // 128 x 256 tiles halve the shared-memory operand traffic per FLOP (an M=128, N=128 MMA reads 8 KB per
  // 64 cycles = the full 128 B/clk of the SM; measured 707 -> 977 TFLOP/s at 256->256 @24x24), but only pay
  // when they still fill the machine
  if (getenv("FU_TC_BN_MAX") == nullptr && N % 256 == 0 &&
      (long long)p.tiles_w * p.tiles_h * p.tiles_b * (N / 256) * 10 >= 9ll * sms)
    bn = 256;
  // ... and the other way round on the smallest levels (6x6: 9 pixel tiles): narrower N tiles until at least ~60 % of
  // the SMs have one (72 CTAs of 128 columns leave half the machine idle; 144 of 64 columns re-read A twice as often
  // but finish sooner)
  // (32-column tiles only when 64-column ones would leave three quarters of the machine idle: every N tile re-reads the same A
  //  tiles through the L2, and at 1024 -> 512 @6x6 (data gradient, 144 k-iterations) 144 CTAs of 32 columns took 46.3 us where 72 CTAs
  //  of 64 columns take 30.9 us and 36 of 128 columns 35.5 us)
  if (getenv("FU_TC_BN_MAX") == nullptr)
    while (bn > 32 && (long long)p.tiles_w * p.tiles_h * p.tiles_b * (N / bn) * 10 < (bn > 64 ? 6ll : 2ll) * sms) bn >>= 1;
  p.BN = bn;
  p.CS = (bn >= 64 && !t.split) ? 64 : 32;
  p.n_tiles = N / p.BN;
  p.fd_ntiles = FastDiv(p.n_tiles); p.fd_tw = FastDiv(p.tiles_w); p.fd_th = FastDiv(p.tiles_h);
  p.cs_shift = p.CS == 64 ? 6 : 5;
  // two 64-channel K chunks per pipeline stage on the deep layers (see TcConvParams::kpair)
  p.kpair = (!t.split && p.KC == 64 && (K / 64) % 2 == 0 && ((p.tw * p.th * p.tn) % 8) == 0 && t.k * t.k * (K / 64) >= 16 &&
             p.BN >= tc_env_int("FU_TC_KPAIR_MINBN", 64) && tc_env_int("FU_TC_KPAIR", 1)) ? 1 : 0;
  p.nstaging = p.BN <= 64 ? 2 : 1;
  if (p.kpair) {      // only where three double stages fit (beside, or overlaid by, the staging tiles): 128-column tiles
    const long long tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b * (N / p.BN);
    const size_t stg = (size_t)tc_pick_groups(p.BN, t.k * t.k * (K / p.KC), p.KC / 16) * p.nstaging * 128 * p.BN * 2;
    const size_t avail = (size_t)227 * 1024 - 4096 - (size_t)12 * N - (tiles <= sms ? 0 : stg);
    if (avail < (size_t)3 * 2 * (128 + p.BN) * p.KC * 2) p.kpair = 0;
  }
  const size_t stage_bytes = (size_t)(128 + p.BN) * p.KC * 2 * (p.kpair ? 2 : 1);
  c.G = tc_pick_groups(p.BN, t.k * t.k * (K / p.KC) * (t.split ? 3 : 1), p.KC / 16);
  if (t.split && c.G > 2) c.G = 2;
  const size_t staging = t.split ? (size_t)c.G * 2 * 16384 + (size_t)c.G * 8 * p.BN * 4 : (size_t)c.G * p.nstaging * 128 * p.BN * 2;
  const size_t fixed = 1024 /*alignment slack*/ + staging + (size_t)12 * N + 16 + 8 * (2 * kTcMaxStages + 18 + 16);
  const size_t budget = 227 * 1024;
  p.a_lo = t.split ? x_ld / 2 : 0; p.b_lo = t.split ? K : 0;
  const long long total_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;
  c.grid = (int)(total_tiles < sms ? total_tiles : sms);
  int stages = (int)((budget - fixed) / stage_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages < 2) { tc_err() = "tile does not fit shared memory"; return nullptr; }
  size_t fixed_run = fixed;
  if (total_tiles <= sms && !t.split && stages < kTcMaxStages && tc_env_int("FU_TC_ALIAS_STAGING", 1)) {
    // one tile per CTA: the staging tiles overlay the (by then idle) operand stages
    int st2 = (int)((budget - (fixed - staging)) / stage_bytes);
    if (st2 > kTcMaxStages) st2 = kTcMaxStages;
    if (st2 > stages && (size_t)st2 * stage_bytes >= staging) { p.alias_staging = 1; stages = st2; fixed_run = fixed - staging; }
  }
  p.stages = stages;
  c.smem = fixed_run + (size_t)stages * stage_bytes;
  if (p.kpair) {
    // the channel dimension split as (64, K/64) with the chunk index as the SLOWEST box dimension: a box of two chunks
    // arrives as two back-to-back [pixels][64] (resp. [BN][64]) tiles
    const int taps = t.k * t.k;
    long long da[5] = {64, W, H, B, K / 64};
    long long sa[5] = {1, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, 64};
    int ba[5] = {64, p.tw, p.th, p.tn, 2};
    if (tc_make_map(&c.a, x, 5, da, sa, ba, 128)) return nullptr;
    long long db[4] = {64, taps, N, K / 64};
    long long sb[4] = {1, K, (long long)taps * K, 64};
    int bb[4] = {64, 1, p.BN, 2};
    if (tc_make_map(&c.b, dir == 0 ? t.w_fwd : t.w_dgrad, 4, db, sb, bb, 128)) return nullptr;
  } else {
  // A: activation (K channels, W, H, B); split mode: the twin's [hi | lo] halves are one channel range
  {
    long long dims[4] = {K + p.a_lo, W, H, B};
    long long str[4] = {1, x_ld, (long long)W * x_ld, (long long)H * W * x_ld};
    int box[4] = {p.KC, p.tw, p.th, p.tn};
    if (tc_make_map(&c.a, x, 4, dims, str, box, p.KC * 2)) return nullptr;
  }
  // B: weights [N][taps][K]  (split: [N][taps][hi K | lo K])
  {
    const int taps = t.k * t.k;
    const long long Kw = t.split ? 2 * K : K;
    long long dims[3] = {Kw, taps, N};
    long long str[3] = {1, Kw, (long long)taps * Kw};
    int box[3] = {p.KC, 1, p.BN};
    if (tc_make_map(&c.b, dir == 0 ? t.w_fwd : t.w_dgrad, 3, dims, str, box, p.KC * 2)) return nullptr;
  }
  }
  // C: output (N channels, W, H, B), bf16 or (split) fp32
  {
    long long dims[4] = {N, W, H, B};
    long long str[4] = {1, y_ld, (long long)W * y_ld, (long long)H * W * y_ld};
    int box[4] = {p.CS, p.tw, p.th, p.tn};
    if (tc_make_map(&c.c, y, 4, dims, str, box, t.split ? 128 : p.CS * 2, t.split)) return nullptr;
  }
  t.cache.push_back(c);
  return &t.cache.back();
}

// 5-D view of an NHWC tensor (B, 2*Hg, 2*Wg, C) as (C, 2, Wg, 2, B*Hg): element (c, b, j, a, r) is
// pixel (row 2*(r % Hg) + a of image r / Hg, column 2j + b).  Rows of consecutive images are contiguous.
inline int tc_make_s2_map(CUtensorMap* m, const void* base, int C, int ld, int Wg, long long Rg, int box_c, int tw, int th,
                          bool f32 = false) {
  long long dims[5] = {C, 2, Wg, 2, Rg};
  long long str[5] = {1, ld, 2ll * ld, 2ll * Wg * ld, 4ll * Wg * ld};
  int box[5] = {box_c, 1, tw, 1, th};
  return tc_make_map(m, base, 5, dims, str, box, box_c * (f32 ? 4 : 2), f32);
}

// Strided layers.  gather = 1: y(Wg,Rg) = sum over the 2x2 taps of x(2Wg,2Rg)  (Conv2d k2s2 forward, ConvT dgrad)
//                  gather = 0: y(2Wg,2Rg) scattered from x(Wg,Rg)               (ConvT forward, Conv2d k2s2 dgrad)
// (Wg, Hg) is the COARSE grid; Rg = B*Hg merged rows.  K / N are the GEMM depth / columns.
inline TcConv::Cached* tc_prepare_s2(TcConv& t, int dir, int gather, const bf16* wmat, int K, int N, int Cst,
                                      const void* x, int x_ld, void* y, int y_ld, int B, int Hg, int Wg, int merge = 1) {
  for (auto& c : t.cache)
    if (c.x == x && c.y == y && c.x_ld == x_ld && c.y_ld == y_ld && c.B == B && c.H == Hg && c.W == Wg && c.dir == dir)
      return &c;
  TcConv::Cached c;
  memset(&c, 0, sizeof(c));
  c.x = x; c.y = y; c.x_ld = x_ld; c.y_ld = y_ld; c.B = B; c.H = Hg; c.W = Wg; c.dir = dir; c.f32 = t.split ? 1 : 0;
  TcConvParams& p = c.p;
  const long long Rg = (long long)B * Hg;
  p.B = 1; p.H = (int)Rg; p.W = Wg; p.Cin = K; p.N = N;
  p.ksz = gather ? 2 : 1; p.pad = 0;
  p.a5 = gather ? 1 : 0; p.c5 = gather ? 0 : 1; p.Cst = Cst;
  if (!gather) {
    if (Cst <= 0 || (Cst & (Cst - 1)) != 0) { tc_err() = "scatter layer: channel count must be a power of two"; return nullptr; }
    for (p.cst_shift = 0; (1 << p.cst_shift) < Cst; ++p.cst_shift) {}
  }
  p.KC = (K % 64 == 0) ? 64 : 32;
  int bn = tc_bn_max();
  p.a_lo = t.split ? x_ld / 2 : 0; p.b_lo = t.split ? K : 0;
  if (gather) {
    while (N % bn) bn >>= 1;
    p.BN = bn;
    p.CS = (bn >= 64 && !t.split) ? 64 : 32;
  } else {
    // scatter: N = 4*Cst columns (a,b,co).  One N tile covers as many whole (a,b) blocks as fit 128 columns, so the
    // A tile is read once for all of them; a store box (CS channels) never straddles two blocks.
    while (Cst % bn) bn >>= 1;                                   // bn <= Cst, divides it
    p.CS = (bn >= 64 && !t.split) ? 64 : 32;
    // (not when the epilogue also reads the old output, i.e. the accumulating downsample data gradient: there the
    //  wider tile's per-chunk second-operand fetches cost more than the A re-reads save -- measured in the step)
    if (merge && tc_env_int("FU_TC_SCATTER_MERGE", 1)) while (bn * 2 <= 128 && N % (bn * 2) == 0) bn *= 2;
    p.BN = bn;
  }
  tc_pick_tile(1, (int)Rg, Wg, p.tw, p.th, p.tn);
  p.tn = 1;
  p.tiles_w = (Wg + p.tw - 1) / p.tw; p.tiles_h = (int)((Rg + p.th - 1) / p.th); p.tiles_b = 1;
  p.n_tiles = N / p.BN;
  p.fd_ntiles = FastDiv(p.n_tiles); p.fd_tw = FastDiv(p.tiles_w); p.fd_th = FastDiv(p.tiles_h);
  p.cs_shift = p.CS == 64 ? 6 : 5;
  const size_t stage_bytes = (size_t)(128 + p.BN) * p.KC * 2;
  p.nstaging = p.BN <= 64 ? 2 : 1;
  c.G = tc_pick_groups(p.BN, (gather ? 4 : 1) * (K / p.KC) * (t.split ? 3 : 1), p.KC / 16);
  if (t.split && c.G > 2) c.G = 2;
  const size_t staging = t.split ? (size_t)c.G * 2 * 16384 + (size_t)c.G * 8 * p.BN * 4 : (size_t)c.G * p.nstaging * 128 * p.BN * 2;
  const size_t fixed = 1024 + staging + (size_t)12 * (gather ? N : Cst) + 16 + 8 * (2 * kTcMaxStages + 18 + 16);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  p.stages = stages;
  c.smem = fixed + (size_t)stages * stage_bytes;
  const long long total_tiles = (long long)p.tiles_w * p.tiles_h * p.n_tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  c.grid = (int)(total_tiles < sms ? total_tiles : sms);
  if (gather) {
    if (tc_make_s2_map(&c.a, x, K + p.a_lo, x_ld, Wg, Rg, p.KC, p.tw, p.th)) return nullptr;
    long long dims[4] = {N, Wg, Rg, 1};
    long long str[4] = {1, y_ld, (long long)Wg * y_ld, Rg * Wg * y_ld};
    int box[4] = {p.CS, p.tw, p.th, 1};
    if (tc_make_map(&c.c, y, 4, dims, str, box, t.split ? 128 : p.CS * 2, t.split)) return nullptr;
  } else {
    long long dims[4] = {K + p.a_lo, Wg, Rg, 1};
    long long str[4] = {1, x_ld, (long long)Wg * x_ld, Rg * Wg * x_ld};
    int box[4] = {p.KC, p.tw, p.th, 1};
    if (tc_make_map(&c.a, x, 4, dims, str, box, p.KC * 2)) return nullptr;
    if (tc_make_s2_map(&c.c, y, Cst, y_ld, Wg, Rg, p.CS, p.tw, p.th, t.split)) return nullptr;
  }
  {
    const int taps = gather ? 4 : 1;
    const long long Kw = t.split ? 2 * K : K;
    long long dims[3] = {Kw, taps, N};
    long long str[3] = {1, Kw, (long long)taps * Kw};
    int box[3] = {p.KC, 1, p.BN};
    if (tc_make_map(&c.b, wmat, 3, dims, str, box, p.KC * 2)) return nullptr;
  }
  t.cache.push_back(c);
  return &t.cache.back();
}

// ---------------------------------------------------------------------------
// halo kernel (tc_conv3_kernel) host side
// ---------------------------------------------------------------------------

inline bool tc_use_v2(const TcConv& t, int H, int W) {
  // measured (tools/conv_time.py, B=32): 32->32@192 73 vs 214 us, 64->64@96 34 vs 61, 128->128@48 33 vs 34, 256->256@24 40 vs 32
  return t.kind == 0 && t.k == 3 && W >= tc_env_int("FU_TC_V2_MINW", 48) && H >= 8 && tc_env_int("FU_TC_V2", 1) != 0;
}

// warp_rows: 0 any box width; 1 prefer a box width that divides a warp (16 / 32: what the kw-stacked MMAs need) when
// that costs at most 4 % more tiles; 2 only such widths
inline void tc_pick_halo_tile(int H, int W, bool halo1, int& twb, int& th, int warp_rows = 0) {
  long long best = -1, best_w = -1, tiles_best = 0, tiles_w = 0;
  int twb_w = 16, th_w = 8;
  twb = 16; th = 8;
  for (int a = 8; a <= 66; a += (halo1 ? 1 : 8)) {
    int b = 128 / a;
    if (b > H) b = H;
    if (b < 1) continue;
    const int two = a - 2;
    if (two < 1) continue;
    const long long tiles = (long long)((W + two - 1) / two) * ((H + b - 1) / b);
    const long long key = tiles * 4096 + (long long)(b + 2) * a;
    if (best < 0 || key < best) { best = key; twb = a; th = b; tiles_best = tiles; }
    if ((a == 16 || a == 32) && (best_w < 0 || key < best_w)) { best_w = key; twb_w = a; th_w = b; tiles_w = tiles; }
  }
  if (best_w >= 0 && (warp_rows == 2 || (warp_rows == 1 && tiles_w * 100 <= tiles_best * 104))) { twb = twb_w; th = th_w; }
}

// res / x2: (dir 1 only) the block's 1x1 shortcut and the gradient G of the block output: dX = conv3x3^T(x) + conv1x1^T(x2)
inline TcConv::Cached3* tc_prepare3(TcConv& t, int dir, const void* x, int x_ld, void* y, int y_ld, int B, int H, int W,
                                     const TcConv* res = nullptr, const void* x2 = nullptr, int x2_ld = 0,
                                     long long y_plane = 0, int y_planeC = 0) {
  for (auto& c : t.cache3)
    if (c.x == x && c.y == y && c.x_ld == x_ld && c.y_ld == y_ld && c.B == B && c.H == H && c.W == W && c.dir == dir &&
        c.x2 == x2 && c.x2_ld == x2_ld && c.p.planeC == y_planeC)
      return &c;
  TcConv::Cached3 c;
  memset(&c, 0, sizeof(c));
  c.x = x; c.y = y; c.x_ld = x_ld; c.y_ld = y_ld; c.B = B; c.H = H; c.W = W; c.dir = dir; c.x2 = x2; c.x2_ld = x2_ld;
  c.f32 = t.split ? 1 : 0;
  TcConv3Params& p = c.p;
  const int K = dir == 0 ? t.Cin : t.Cout;
  const int N = dir == 0 ? t.Cout : t.Cin;
  p.B = B; p.H = H; p.W = W; p.K = K; p.N = N;
  p.KC = (K % 64 == 0) ? 64 : 32;
  { const int kc = tc_env_int("FU_TC_KC", 0); if ((kc == 16 || kc == 32 || kc == 64) && K % kc == 0) p.KC = kc; }
  int bn = 128;
  while (N % bn) bn >>= 1;
  p.BN = bn; p.CS = (bn >= 64 && !t.split) ? 64 : 32;
  p.planeC = 0;
  if (y_planeC) {               // planar output: one store box per plane slice
    if (t.split || y_planeC != 32 || N % y_planeC) { tc_err() = "planar output: 32-channel planes, bf16 storage only"; return nullptr; }
    p.planeC = y_planeC; p.CS = 32;
  }
  p.n_tiles = N / bn;
  p.a_lo = t.split ? x_ld / 2 : 0; p.b_lo = t.split ? K : 0;
  p.a2_lo = t.split ? x2_ld / 2 : 0; p.b2_lo = p.b_lo;
  p.halo1 = tc_env_int("FU_TC_HALO1", 1) ? 1 : 0;
  p.npair = tc_env_int("FU_TC_PAIR", 1) ? 2 : 1;
  // (a layer that can run the kw-stacked MMAs -- 32 output columns, weights that will be resident -- prefers box widths 16 / 32)
  const int stack_env = tc_env_int("FU_TC_STACK", 0);      // 0: off (default, see below), 1: K >= 64 layers, 2: every 32-column layer
  // Measured (B = 32 @192x192, us per launch, stacked vs not): K = 64: forward 67.8 vs 78.6, data gradient 63.2 vs 73.7 (@96x96 19.3
  // vs 25.4); K = 32: 64.2 vs 51.2 / 47.9 vs 43.1 -- with 18 MMAs per tile those layers are bound by their epilogue, which the
  // stacked form makes heavier (three accumulator blocks to read, 64 shuffles per thread).  So: K >= 64 only (FU_TC_STACK=2: all).
  // In the graph-replayed step the K >= 64 gains do not show (same-box A/B against the build without this code: 5.1589 vs 5.1582 ms
  // per step), so the mode is opt-in.  What the experiment established: per-MMA cost of these kernels is ~40 + 1.2 * N cycles
  // (N = 32: 55-75, N = 64: 100-150, N = 96 stacked: ~140), i.e. proportional to N, not dominated by re-reading the A rows.
  const bool stack_cand = stack_env && p.halo1 && !t.split && p.BN == 32 && p.n_tiles == 1 && (K >= 64 || stack_env >= 2);
  tc_pick_halo_tile(H, W, p.halo1 != 0, p.twb, p.th, stack_cand ? 2 : 0);
  p.two = p.twb - 2;
  p.tiles_w = (W + p.two - 1) / p.two; p.tiles_h = (H + p.th - 1) / p.th;
  p.fd_ntiles = FastDiv(p.n_tiles); p.fd_timg = FastDiv(p.tiles_w * p.tiles_h); p.fd_tw = FastDiv(p.tiles_w);
  p.cs_shift = p.CS == 64 ? 6 : 5;
  const size_t row_bytes = (size_t)p.KC * 2;
  size_t rows = (size_t)(p.th + 2) * p.twb + 2;
  if (rows < (size_t)(2 * p.twb + 2 + 128)) rows = (size_t)(2 * p.twb + 2 + 128);
  p.a_tile_bytes = (unsigned)((rows * row_bytes + 1023) / 1024 * 1024);
  const size_t a_stage = (size_t)p.npair * p.a_tile_bytes;
  const size_t b_bytes = (size_t)p.BN * row_bytes;
  const size_t budget = 227 * 1024;
  p.res = (res && x2) ? 1 : 0;
  const size_t wbytes = (size_t)(9 + p.res) * K * p.BN * 2 * (t.split ? 2 : 1);
  // two epilogue sets for the thin layers when everything (resident weights, >= 2 A stages) still fits
  // (measured: 32-column layers 75 -> 70 us, 32->64 dgrad @192 95 -> 77 us; 64->64 @96 35.5 -> 37 us: no gain at K >= 576)
  // The forward launches (dir 0) also compute the BN statistics in their epilogue, which roughly doubles its cost
  // (64->64 @96x96: 35 us without, 56 us with statistics): there two sets pay at every thin shape that fits.
  c.S = (p.BN <= 64 && (dir == 0 || (long long)K * p.BN <= 64 * 32) && p.npair == 2 && tc_env_int("FU_TC_EPI_SETS", 2) >= 2) ? 2 : 1;
  if (t.split) c.S = 1;            // fp32 staging: 32 KB per epilogue group
  // preference order when shared memory is short: (2 sets, 2 staging tiles) -> (2 sets, 1) -> (1 set, 2) -> (1 set, 1)
  p.nstg = (p.BN <= 64 && !t.split && tc_env_int("FU_TC_STAGING2", 1)) ? 2 : 1;
  const int want_S = c.S, want_nstg = p.nstg;
  // candidate (epilogue sets, staging tiles) in order of preference
  int cand[4][2]; int ncand = 0;
  cand[ncand][0] = want_S; cand[ncand++][1] = want_nstg;
  if (want_S == 2 && want_nstg == 2) { cand[ncand][0] = 2; cand[ncand++][1] = 1; }
  if (want_S == 2) { cand[ncand][0] = 1; cand[ncand++][1] = want_nstg; }
  if (want_S != 1 || want_nstg != 1) { cand[ncand][0] = 1; cand[ncand++][1] = 1; }
  auto fixed_of = [&](int S_, int nstg_) {
    const size_t staging = t.split ? (size_t)S_ * p.npair * 32768 : (size_t)S_ * p.npair * 128 * p.BN * 2 * nstg_;
    return 1024 + staging + (size_t)12 * N + (t.split ? (size_t)S_ * p.npair * tc3_park_floats(p.BN, p.CS) * 4 : 0) + 16 + 8 * 48;
  };
  // A stages one super tile consumes: with fewer than twice that many the two MMA issuers (and their baton) are off,
  // which costs more than a second staging tile gains (32->64 data gradient @192x192: 89 us with 4 stages and one staging
  // tile, 112 us with 2 stages and two).  Pass 0 only accepts resident configurations that keep them, pass 1 any resident one.
  const int a_per_tile_h = (K / p.KC) * (t.split ? 3 : 1) * (p.halo1 ? 1 : 3) * (p.res ? 2 : 1);
  const bool dual_possible = tc_env_int("FU_TC_DUAL", 1) != 0 && 2 * a_per_tile_h <= 4;
  p.resident = 0;
  for (int pass = dual_possible ? 0 : 1; pass < 2 && !p.resident; ++pass)
    for (int ci = 0; ci < ncand && !p.resident; ++ci) {
      const size_t fixed = fixed_of(cand[ci][0], cand[ci][1]);
      if (p.n_tiles != 1 || fixed + wbytes + 2 * a_stage > budget || !tc_env_int("FU_TC_RESIDENT", 1)) continue;
      int as = (int)((budget - fixed - wbytes) / a_stage);
      if (as > 4) as = 4;
      if (pass == 0 && as < 2 * a_per_tile_h) continue;
      p.resident = 1; c.S = cand[ci][0]; p.nstg = cand[ci][1];
      p.a_stages = as; p.b_stages = 1;
      c.smem = fixed + wbytes + (size_t)as * a_stage;
    }
  if (!p.resident) {
    c.S = 1; p.nstg = 1;
    const size_t fixed = fixed_of(1, 1);
    p.a_stages = 2;
    if (fixed + 3 * a_stage + 6 * b_bytes <= budget) p.a_stages = 3;
    if (fixed + (size_t)p.a_stages * a_stage + 2 * b_bytes > budget) { tc_err() = "halo tile does not fit shared memory"; return nullptr; }
    int bs = (int)((budget - fixed - (size_t)p.a_stages * a_stage) / b_bytes);
    if (bs > 8) bs = 8;
    if (bs < 2) { tc_err() = "halo tile does not fit shared memory"; return nullptr; }
    p.b_stages = bs;
    c.smem = fixed + (size_t)p.a_stages * a_stage + (size_t)bs * b_bytes;
  }
  // kw-stacked MMAs (TcConv3Params::stack): 32-column layers with resident weights whose box width divides a warp
  p.stack = (p.resident && p.halo1 && !t.split && p.BN == 32 && (p.twb == 16 || p.twb == 32) && p.twb * p.th <= 128 && stack_cand && c.S == 2) ? 1 : 0;
  p.nacc = p.stack ? 2 : 2 * c.S;
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * B;
  const long long total_super = (m_tiles + p.npair - 1) / p.npair * p.n_tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  c.grid = (int)(total_super < sms ? total_super : sms);
  {
    const int vchunks = (K / p.KC) * (t.split ? 3 : 1);
    const int a_per_tile = vchunks * (p.halo1 ? 1 : 3) + (p.res ? vchunks : 0);
    p.dual = (p.resident && p.a_stages >= 2 * a_per_tile && tc_env_int("FU_TC_DUAL", 1)) ? 1 : 0;
    if (p.dual && tc_env_int("FU_TC_BATON", 1)) p.dual = 2;
    p.baton_kh = tc_env_int("FU_TC_BATON_KH", 2);   // measured (32->32 @192x192 dgrad): row 0: 55.0, 1: 50.3, 2: 47.6 us; no baton 56.4
  }
  const long long Kw = t.split ? 2 * K : K;       // K width of the weight layouts ([hi | lo] in split mode)
  {
    long long dims[4] = {K + p.a_lo, W, H, B};
    long long str[4] = {1, x_ld, (long long)W * x_ld, (long long)H * W * x_ld};
    int box[4] = {p.KC, p.twb, p.th + 2, 1};
    if (tc_make_map(&c.a, x, 4, dims, str, box, p.KC * 2)) return nullptr;
  }
  {
    long long dims[3] = {Kw, 9, N};
    long long str[3] = {1, Kw, 9ll * Kw};
    int box[3] = {p.KC, 1, p.BN};
    if (tc_make_map(&c.b, dir == 0 ? t.w_fwd : t.w_dgrad, 3, dims, str, box, p.KC * 2)) return nullptr;
  }
  if (p.planeC) {
    long long dims[5] = {p.planeC, W, H, B, N / p.planeC};
    long long str[5] = {1, p.planeC, (long long)W * p.planeC, (long long)H * W * p.planeC, y_plane};
    int box[5] = {p.CS, p.two, p.th, 1, 1};
    if (tc_make_map(&c.c, y, 5, dims, str, box, p.CS * 2)) return nullptr;
  } else {
    long long dims[4] = {N, W, H, B};
    long long str[4] = {1, y_ld, (long long)W * y_ld, (long long)H * W * y_ld};
    int box[4] = {p.CS, p.two, p.th, 1};
    if (tc_make_map(&c.c, y, 4, dims, str, box, t.split ? 128 : p.CS * 2, t.split)) return nullptr;
  }
  c.a2 = c.a; c.b2 = c.b;
  if (p.res) {
    long long dims[4] = {K + p.a2_lo, W, H, B};
    long long str[4] = {1, x2_ld, (long long)W * x2_ld, (long long)H * W * x2_ld};
    int box[4] = {p.KC, p.twb, p.th + 2, 1};
    if (tc_make_map(&c.a2, x2, 4, dims, str, box, p.KC * 2)) return nullptr;
    long long wd[3] = {Kw, 1, N};
    long long ws[3] = {1, Kw, Kw};
    int wb[3] = {p.KC, 1, p.BN};
    if (tc_make_map(&c.b2, res->w_dgrad, 3, wd, ws, wb, p.KC * 2)) return nullptr;      // [Cin][1][Cout] of the 1x1
  }
  if (tc_env_int("FU_TC_VERBOSE", 0))
    fprintf(stderr, "[tc_conv3] dir %d B %d %dx%d K %d N %d: KC %d BN %d box %dx%d (valid %dx%d) tiles %lld S %d nstg %d resident %d a_stages %d "
            "b_stages %d dual %d stack %d nacc %d res %d grid %d smem %zu\n", dir, B, H, W, K, N, p.KC, p.BN, p.twb, p.th + 2, p.two, p.th,
            m_tiles, c.S, p.nstg, p.resident, p.a_stages, p.b_stages, p.dual, p.stack, p.nacc, p.res, c.grid, (size_t)c.smem);
  t.cache3.push_back(c);
  return &t.cache3.back();
}

inline int tc_launch3(TcConv::Cached3* c, cudaStream_t stream, fu_counters* cnt) {
  static bool attr_done[64] = {};
  if (tc_attr_needed(attr_done)) {
    if (cudaFuncSetAttribute(tc_conv3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv3_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      tc_err() = "cudaFuncSetAttribute(max dynamic smem, conv3) failed";
      return -1;
    }
  }
  static long long* dbg_buf = nullptr;
  const bool dbg = tc_env_int("FU_TC_DBG", 0) != 0;
  if (dbg) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 4 * 24 * 4 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4 * 24 * 4 * sizeof(long long), stream);
  }
  c->p.dbg = dbg ? dbg_buf : nullptr;
  const bool pdl = fu_pdl_enabled() && !dbg;
  if (c->f32) fu_launch(tc_conv3_kernel<1, true>, dim3(c->grid), dim3(96 + 256), c->smem, stream, pdl, c->a, c->b, c->c, c->a2, c->b2, c->p);
  else if (c->S == 2 && !c->p.t && c->p.stack) fu_launch(tc_conv3_kernel<2, false, true>, dim3(c->grid), dim3(96 + 256 * 2), c->smem, stream, pdl, c->a, c->b, c->c, c->a2, c->b2, c->p);
  else if (c->S == 2 && !c->p.t) fu_launch(tc_conv3_kernel<2, false>, dim3(c->grid), dim3(96 + 256 * 2), c->smem, stream, pdl, c->a, c->b, c->c, c->a2, c->b2, c->p);
  else fu_launch(tc_conv3_kernel<1, false>, dim3(c->grid), dim3(96 + 256), c->smem, stream, pdl, c->a, c->b, c->c, c->a2, c->b2, c->p);
  if (dbg) {
    long long h[4 * 24 * 4];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    long long t0 = h[0];
    fprintf(stderr, "[tc_conv3 timeline, CTA 0, cycles rel. to first producer issue] twb=%d th=%d BN=%d KC=%d pair=%d resident=%d a_stages=%d b_stages=%d grid=%d\n",
            c->p.twb, c->p.th, c->p.BN, c->p.KC, c->p.npair, c->p.resident, c->p.a_stages, c->p.b_stages, c->grid);
    for (int i = 0; i < 12; ++i)
      fprintf(stderr, "super %2d: prod %7lld | mma at_wait %7lld acc_free %7lld fenced %7lld a_full %7lld first_mma %7lld committed %7lld | epi0 start %7lld t_full %7lld drained %7lld stored %7lld | epi1 t_full %7lld stored %7lld\n",
              i, h[(0 * 24 + i) * 4] - t0, h[(0 * 24 + i) * 4 + 3] - t0, h[(0 * 24 + i) * 4 + 1] - t0, h[(1 * 24 + i) * 4] - t0,
              h[(1 * 24 + i) * 4 + 1] - t0, h[(0 * 24 + i) * 4 + 2] - t0, h[(1 * 24 + i) * 4 + 2] - t0,
              h[(2 * 24 + i) * 4] - t0, h[(2 * 24 + i) * 4 + 1] - t0, h[(2 * 24 + i) * 4 + 2] - t0, h[(2 * 24 + i) * 4 + 3] - t0,
              h[(3 * 24 + i) * 4 + 1] - t0, h[(3 * 24 + i) * 4 + 3] - t0);
  }
  if (cnt) { cnt->kernel_launches++; cnt->tc_kernel_launches++; }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { tc_err() = cudaGetErrorString(e); return -1; }
  return 0;
}

inline int tc_launch(TcConv::Cached* c, cudaStream_t stream, fu_counters* cnt) {
  static bool attr_done[64] = {};
  if (tc_attr_needed(attr_done)) {
    if (cudaFuncSetAttribute(tc_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tc_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      tc_err() = "cudaFuncSetAttribute(max dynamic smem) failed";
      return -1;
    }
  }
  const bool pdl = fu_pdl_enabled();
  // second epilogue operand through shared memory (bf16 storage, thin tiles): two `t` tiles per epilogue group are
  // carved from the operand stages; the map is the output's own (accumulate) or one of the same geometry over `t`
  if (c->stages_base == 0) { c->stages_base = c->p.stages; c->smem_base_bytes = c->smem; }
  c->p.t_smem = 0; c->p.stages = c->stages_base; c->smem = c->smem_base_bytes;
  if (c->p.t && !c->f32 && c->p.BN <= 64 && tc_env_int("FU_TC_T_SMEM", 1)) {
    const size_t stage_bytes = (size_t)(128 + c->p.BN) * c->p.KC * 2;
    const size_t tbytes = (size_t)c->G * kTcTBufs * 128 * c->p.BN * 2;
    const size_t fixed = c->smem_base_bytes - (size_t)c->stages_base * stage_bytes;
    int st = fixed + tbytes < (size_t)227 * 1024 ? (int)(((size_t)227 * 1024 - fixed - tbytes) / stage_bytes) : 0;
    if (st > c->stages_base) st = c->stages_base;
    bool ok = st >= 2;
    if (ok && c->p.t == reinterpret_cast<const bf16*>(c->y) && c->p.t_ld == c->y_ld) {
      c->tm_t = c->c;                                   // accumulate: the old output is read through the store map
    } else if (ok && !c->p.c5) {
      if (c->tm_t_ptr != c->p.t || c->tm_t_ld != c->p.t_ld) {
        long long dims[4] = {c->p.N, c->p.W, c->p.H, c->p.B};
        long long str[4] = {1, c->p.t_ld, (long long)c->p.W * c->p.t_ld, (long long)c->p.H * c->p.W * c->p.t_ld};
        int box[4] = {c->p.CS, c->p.tw, c->p.th, c->p.tn};
        if (tc_make_map(&c->tm_t, c->p.t, 4, dims, str, box, c->p.CS * 2)) return -1;
        c->tm_t_ptr = c->p.t; c->tm_t_ld = c->p.t_ld;
      }
    } else {
      ok = false;
    }
    if (ok) { c->p.t_smem = 1; c->p.stages = st; c->smem = fixed + tbytes + (size_t)st * stage_bytes; }
  }
  if (!c->p.t_smem) c->tm_t = c->c;
  const int G_run = (c->G == 4 && c->p.t && !c->p.t_smem) ? 2 : c->G;     // (G = 4 has no global-load path for `t`)
  {
    const int k_iters = c->p.ksz * c->p.ksz * (c->p.Cin / c->p.KC) * (c->f32 ? 3 : 1);
    c->p.dual = (c->p.stages >= 2 * k_iters && tc_env_int("FU_TC_DUAL", 1)) ? 1 : 0;
  }
  {
    // clusters of two one-tile CTAs sharing an operand through TMA multicast (see TcConvParams::mc)
    const long long total_tiles = (long long)c->p.tiles_w * c->p.tiles_h * c->p.tiles_b * c->p.n_tiles;
    const int k_iters = c->p.ksz * c->p.ksz * (c->p.Cin / c->p.KC);
    c->p.mc = 0;
    if (!c->f32 && !c->p.a5 && !c->p.c5 && !c->p.dual && !c->p.kpair && c->grid == total_tiles && (c->grid % 2) == 0 && k_iters >= 8 &&
        tc_env_int("FU_TC_MC", 0)) {      // measured 3-8 % SLOWER on the 12x12 / 24x24 layers: off unless asked for
      if (c->p.n_tiles % 2 == 0) c->p.mc = 1;
      else if (c->p.n_tiles == 1) c->p.mc = 2;
    }
  }
  if (c->p.mc) {
    if (G_run == 4) fu_launch_cluster(tc_conv_kernel<4, false>, dim3(c->grid), dim3(96 + 128 * 4), c->smem, stream, pdl, 2u, c->a, c->b, c->c, c->tm_t, c->p);
    else if (G_run == 2) fu_launch_cluster(tc_conv_kernel<2, false>, dim3(c->grid), dim3(96 + 128 * 2), c->smem, stream, pdl, 2u, c->a, c->b, c->c, c->tm_t, c->p);
    else fu_launch_cluster(tc_conv_kernel<1, false>, dim3(c->grid), dim3(96 + 128), c->smem, stream, pdl, 2u, c->a, c->b, c->c, c->tm_t, c->p);
  } else
  if (c->f32 && c->G == 2) fu_launch(tc_conv_kernel<2, true>, dim3(c->grid), dim3(96 + 128 * 2), c->smem, stream, pdl, c->a, c->b, c->c, c->tm_t, c->p);
  else if (c->f32) fu_launch(tc_conv_kernel<1, true>, dim3(c->grid), dim3(96 + 128), c->smem, stream, pdl, c->a, c->b, c->c, c->tm_t, c->p);
  else if (G_run == 4) fu_launch(tc_conv_kernel<4, false>, dim3(c->grid), dim3(96 + 128 * 4), c->smem, stream, pdl, c->a, c->b, c->c, c->tm_t, c->p);
  else if (G_run == 2) fu_launch(tc_conv_kernel<2, false>, dim3(c->grid), dim3(96 + 128 * 2), c->smem, stream, pdl, c->a, c->b, c->c, c->tm_t, c->p);
  else fu_launch(tc_conv_kernel<1, false>, dim3(c->grid), dim3(96 + 128), c->smem, stream, pdl, c->a, c->b, c->c, c->tm_t, c->p);
  if (cnt) { cnt->kernel_launches++; cnt->tc_kernel_launches++; }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { tc_err() = cudaGetErrorString(e); return -1; }
  return 0;
}

// (split mode: x / x_ld describe the operand's bf16 twin, y and tp are fp32 tensors)
inline bool tc_conv_eligible(const TcConv& t, const void* x, int x_ld, const void* y, int y_ld, const void* tp, int t_ld) {
  if (t.split) return t.enabled && tc_ptr_ok(x, x_ld) && tc_ptr_ok_f32(y, y_ld) && (!tp || tc_ptr_ok_f32(tp, t_ld));
  return t.enabled && tc_ptr_ok(x, x_ld) && tc_ptr_ok(y, y_ld) && (!tp || tc_ptr_ok(tp, t_ld));
}

// fin (optional, 1x1 / first-generation kernel only): finalise the BatchNorm whose folded scale / shift the epilogue applies
// to `tp` inside this launch (see TcBnFin) instead of reading bn_a / bn_b
inline int tc_conv_forward(TcConv& t, const void* x, int x_ld, void* y, int y_ld, int B, int H, int W, const float* bias,
                           int relu, double* stat, const void* tp, int t_ld, const float* bn_a, const float* bn_b,
                           int accumulate, cudaStream_t stream, fu_counters* cnt, const TcBnFin* fin = nullptr, int post = 0) {
  if (fin && tc_use_v2(t, H, W)) { tc_err() = "BN finalisation can only ride on the 1x1 kernel"; return -1; }
  if (tc_use_v2(t, H, W)) {
    TcConv::Cached3* c3 = tc_prepare3(t, 0, x, x_ld, y, y_ld, B, H, W);
    if (c3) {
      c3->p.bias = bias; c3->p.relu = relu; c3->p.stat = stat;
      tc_set_t(t, c3->p, tp, t_ld); c3->p.bn_a = bn_a; c3->p.bn_b = bn_b; c3->p.post = post;
      if (accumulate) { tc_set_t(t, c3->p, y, y_ld); c3->p.bn_a = nullptr; c3->p.bn_b = nullptr; }
      return tc_launch3(c3, stream, cnt);
    }   // else: the halo configuration does not fit shared memory for this shape -> first-generation kernel
  }
  TcConv::Cached* c = tc_prepare(t, 0, x, x_ld, y, y_ld, B, H, W);
  if (!c) return -1;
  c->p.bias = bias; c->p.relu = relu; c->p.stat = stat;
  tc_set_t(t, c->p, tp, t_ld); c->p.bn_a = bn_a; c->p.bn_b = bn_b; c->p.post = post;
  if (accumulate) { tc_set_t(t, c->p, y, y_ld); c->p.bn_a = nullptr; c->p.bn_b = nullptr; }
  if (fin) c->p.fin = *fin; else memset(&c->p.fin, 0, sizeof(c->p.fin));
  return tc_launch(c, stream, cnt);
}

inline bool tc_dgrad_eligible(const TcConv& t, const void* dy, int dy_ld, const void* dx, int dx_ld) {
  return t.enabled && tc_ptr_ok(dy, dy_ld) && (t.split ? tc_ptr_ok_f32(dx, dx_ld) : tc_ptr_ok(dx, dx_ld));
}

// stat (optional, [2*Cin] doubles, zeroed by the caller): per-channel sum / sum of squares of the STORED dx, i.e.
// of the final value when accumulate = 1 -- the bias gradient of the layer that produced dx's forward twin
// res / g: fuse the data gradient of the block's 1x1 shortcut (dX += conv1x1^T(g)) into this launch; only the halo
// kernel can (tc_dgrad_can_fuse_res)
inline bool tc_dgrad_can_fuse_res(const TcConv& t, const TcConv& res, int H, int W, const void* g, int g_ld) {
  return tc_use_v2(t, H, W) && res.enabled && res.kind == 0 && res.k == 1 && res.Cin == t.Cin && res.Cout == t.Cout &&
         res.split == t.split && tc_ptr_ok(g, g_ld) && tc_env_int("FU_TC_FUSE_RES", 1) != 0;
}
inline int tc_conv_dgrad(TcConv& t, const void* dy, int dy_ld, void* dx, int dx_ld, int B, int H, int W, int accumulate,
                         cudaStream_t stream, fu_counters* cnt, double* stat = nullptr, const TcConv* res = nullptr,
                         const void* g = nullptr, int g_ld = 0, long long dx_plane = 0, int dx_planeC = 0) {
  if (dx_planeC && (!tc_use_v2(t, H, W) || accumulate)) { tc_err() = "planar output needs the halo kernel"; return -1; }
  if (tc_use_v2(t, H, W)) {
    TcConv::Cached3* c3 = tc_prepare3(t, 1, dy, dy_ld, dx, dx_ld, B, H, W, res, g, g_ld, dx_plane, dx_planeC);
    if (!c3 && dx_planeC) return -1;
    if (c3) {
      c3->p.bias = nullptr; c3->p.relu = 0; c3->p.stat = stat; c3->p.bn_a = nullptr; c3->p.bn_b = nullptr;
      tc_set_t(t, c3->p, accumulate ? dx : nullptr, dx_ld);
      return tc_launch3(c3, stream, cnt);
    }
  }
  if (res) return -2;          // not fusable here (no halo configuration): the caller runs the two kernels
  TcConv::Cached* c = tc_prepare(t, 1, dy, dy_ld, dx, dx_ld, B, H, W);
  if (!c) return -1;
  c->p.bias = nullptr; c->p.relu = 0; c->p.stat = stat; c->p.bn_a = nullptr; c->p.bn_b = nullptr;
  tc_set_t(t, c->p, accumulate ? dx : nullptr, dx_ld);
  return tc_launch(c, stream, cnt);
}

// ---- Conv2d(C,C,2,stride 2) (unet.py:93) : x (B,H,W,Cin) -> y (B,H/2,W/2,Cout) ----
inline bool tc_down_eligible(const TcConv& t, const void* x, int x_ld, const void* y, int y_ld) {
  if (t.split) return t.enabled && t.kind == 1 && tc_ptr_ok(x, x_ld) && tc_ptr_ok_f32(y, y_ld);
  return t.enabled && t.kind == 1 && tc_ptr_ok(x, x_ld) && tc_ptr_ok(y, y_ld);
}
inline int tc_down_forward(TcConv& t, const void* x, int x_ld, void* y, int y_ld, int B, int H, int W, const float* bias,
                           cudaStream_t stream, fu_counters* cnt) {
  TcConv::Cached* c = tc_prepare_s2(t, 0, 1, t.w_fwd, t.Cin, t.Cout, 0, x, x_ld, y, y_ld, B, H / 2, W / 2);
  if (!c) return -1;
  c->p.bias = bias; c->p.relu = 0; c->p.stat = nullptr; c->p.t = nullptr; c->p.tf = nullptr; c->p.bn_a = nullptr; c->p.bn_b = nullptr;
  return tc_launch(c, stream, cnt);
}
// dx (B,H,W,Cin) (+)= scatter of dy (B,H/2,W/2,Cout)
inline int tc_down_dgrad(TcConv& t, const void* dy, int dy_ld, void* dx, int dx_ld, int B, int H, int W, int accumulate,
                         cudaStream_t stream, fu_counters* cnt) {
  TcConv::Cached* c = tc_prepare_s2(t, 1, 0, t.w_dgrad, t.Cout, 4 * t.Cin, t.Cin, dy, dy_ld, dx, dx_ld, B, H / 2, W / 2,
                                    accumulate ? 0 : 1);
  if (!c) return -1;
  c->p.bias = nullptr; c->p.relu = 0; c->p.stat = nullptr; c->p.bn_a = nullptr; c->p.bn_b = nullptr;
  tc_set_t(t, c->p, accumulate ? dx : nullptr, dx_ld);
  return tc_launch(c, stream, cnt);
}
// ---- ConvTranspose2d(Cin,Cout,2,stride 2) (unet.py:240) : x (B,h,w,Cin) -> y (B,2h,2w,Cout) ----
inline bool tc_up_eligible(const TcConv& t, const void* x, int x_ld, const void* y, int y_ld) {
  if (t.split) return t.enabled && t.kind == 2 && tc_ptr_ok(x, x_ld) && tc_ptr_ok_f32(y, y_ld);
  return t.enabled && t.kind == 2 && tc_ptr_ok(x, x_ld) && tc_ptr_ok(y, y_ld);
}
inline int tc_up_forward(TcConv& t, const void* x, int x_ld, void* y, int y_ld, int B, int h, int w, const float* bias,
                         cudaStream_t stream, fu_counters* cnt) {
  TcConv::Cached* c = tc_prepare_s2(t, 0, 0, t.w_fwd, t.Cin, 4 * t.Cout, t.Cout, x, x_ld, y, y_ld, B, h, w);
  if (!c) return -1;
  c->p.bias = bias; c->p.relu = 0; c->p.stat = nullptr; c->p.t = nullptr; c->p.tf = nullptr; c->p.bn_a = nullptr; c->p.bn_b = nullptr;
  return tc_launch(c, stream, cnt);
}
// dx (B,h,w,Cin) = gather of dy (B,2h,2w,Cout)
inline int tc_up_dgrad(TcConv& t, const void* dy, int dy_ld, void* dx, int dx_ld, int B, int h, int w, cudaStream_t stream,
                       fu_counters* cnt) {
  TcConv::Cached* c = tc_prepare_s2(t, 1, 1, t.w_dgrad, t.Cout, t.Cin, 0, dy, dy_ld, dx, dx_ld, B, h, w);
  if (!c) return -1;
  c->p.bias = nullptr; c->p.relu = 0; c->p.stat = nullptr; c->p.t = nullptr; c->p.tf = nullptr; c->p.bn_a = nullptr; c->p.bn_b = nullptr;
  return tc_launch(c, stream, cnt);
}

inline bool tc_wgrad_eligible(const TcConv& t, const void* x, int x_ld, const void* dy, int dy_ld) {
  return t.enabled && tc_ptr_ok(x, x_ld) && tc_ptr_ok(dy, dy_ld) && getenv("FU_TC_NO_WGRAD") == nullptr;
}

inline void tc_pick_tile64(int B, int H, int W, int& tw, int& th, int& tn, int npx = 64) {
  long long best = -1;
  tw = 8; th = 8; tn = npx / 64;
  for (int a = 1; a <= npx; a <<= 1)
    for (int b = 1; a * b <= npx; b <<= 1) {
      const int c = npx / (a * b);
      const long long tiles = (long long)((W + a - 1) / a) * ((H + b - 1) / b) * ((B + c - 1) / c);
      const long long key = tiles * 1024 - a;
      if (best < 0 || key < best) { best = key; tw = a; th = b; tn = c; }
    }
}

// Common weight-gradient launcher.  `a` = operand read at the pixel itself (M channels -> accumulator rows),
// `b` = operand read at the tap-shifted pixel (Nn channels -> accumulator columns).  s2 = 0: 3x3/1x1 taps on the
// same (B,H,W) grid; s2 = 1: `b` lives on the 2x finer grid and the 2x2 taps select its sub-pixels.
// Result: dw[(m*Nn + n)*taps + tap] (torch layouts (Cout,Cin,k,k) resp. ConvTranspose's (Cin,Cout,2,2)).
// t.dw_acc must have been zeroed since its last use.
inline int tc_wgrad_common(TcConv& t, const void* a, int a_ld, int M, const void* b, int b_ld, int Nn, int s2, int ksz,
                           int B, int H, int W, float* dw, cudaStream_t stream, fu_counters* cnt) {
  TcConv::WCached* c = nullptr;
  for (auto& k : t.wcache)
    if (k.x == b && k.dy == a && k.x_ld == b_ld && k.dy_ld == a_ld && k.B == B && k.H == H && k.W == W) { c = &k; break; }
  if (!c) {
    TcConv::WCached n;
    memset(&n, 0, sizeof(n));
    n.x = b; n.dy = a; n.x_ld = b_ld; n.dy_ld = a_ld; n.B = B; n.H = H; n.W = W;
    TcWgradParams& p = n.p;
    const long long Rg = (long long)B * H;
    p.Cin = Nn; p.Cout = M; p.ksz = ksz; p.pad = ksz == 3 ? 1 : 0; p.b5 = s2;
    p.taps_per_cta = ksz; p.groups = ksz;
    p.N = Nn > 64 ? 128 : 64;
    // 128-pixel stages when at least three of them fit and every CTA still gets a few
    p.kpix = ((2 + p.taps_per_cta * (p.N / 64)) * 16384 * 3 + 2048 <= 227 * 1024 && (long long)B * H * W >= 148ll * 4 * 128 &&
              tc_env_int("FU_TC_WGRAD_KPIX", 128) == 128) ? 128 : 64;
    if (s2) {
      p.B = 1; p.H = (int)Rg; p.W = W;
      tc_pick_tile64(1, (int)Rg, W, p.tw, p.th, p.tn, p.kpix); p.tn = 1;
      if (p.tw * p.th != p.kpix) { p.kpix = 64; tc_pick_tile64(1, (int)Rg, W, p.tw, p.th, p.tn); p.tn = 1; }
    } else { p.B = B; p.H = H; p.W = W; tc_pick_tile64(B, H, W, p.tw, p.th, p.tn, p.kpix); }
    if (s2 && p.tw * p.th != 64 && p.tw * p.th != p.kpix) { tc_err() = "strided wgrad: no 64-pixel tile"; return -1; }
    p.tiles_w = (p.W + p.tw - 1) / p.tw; p.tiles_h = (p.H + p.th - 1) / p.th; p.tiles_b = (p.B + p.tn - 1) / p.tn;
    p.co_tiles = (M + 127) / 128; p.ci_tiles = (Nn + p.N - 1) / p.N;
    const size_t stage_bytes = (size_t)(2 + p.taps_per_cta * (p.N / 64)) * 128 * p.kpix;
    const size_t fixed = 1024 + 8 * (2 * kTcMaxStages + 4);
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > kTcMaxStages) stages = kTcMaxStages;
    if (stages < 2) { tc_err() = "wgrad tile does not fit shared memory"; return -1; }
    p.stages = stages;
    n.smem = fixed + (size_t)stages * stage_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long units = (long long)p.co_tiles * p.ci_tiles * p.groups;
    const long long total_pt = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
    const long long max_splits = (total_pt + 3) / 4;          // at least 4 pixel tiles per CTA
    long long splits = tc_pick_splits(units, max_splits, sms, tc_env_int("FU_TC_WGRAD_WAVES", 1));
    const long long per = (total_pt + splits - 1) / splits;
    splits = (total_pt + per - 1) / per;
    p.splits = (int)splits;
    n.grid = (int)(units * splits);
    p.dw_acc = t.dw_acc;
    p.skip_epi = tc_env_int("FU_TC_WGRAD_NOEPI", 0);
    p.m64 = M <= 64 ? tc_env_int("FU_TC_M64", 2) : 0;
    p.split = t.split ? 1 : 0; p.y_lo = t.split ? a_ld / 2 : 0; p.x_lo = t.split ? b_ld / 2 : 0;
    {
      long long dims[4] = {M + p.y_lo, p.W, p.H, p.B};
      long long str[4] = {1, a_ld, (long long)p.W * a_ld, (long long)p.H * p.W * a_ld};
      int box[4] = {64, p.tw, p.th, p.tn};
      if (tc_make_map(&n.y, a, 4, dims, str, box, 128)) return -1;
    }
    if (s2) {
      if (tc_make_s2_map(&n.xm, b, Nn + p.x_lo, b_ld, W, Rg, 64, p.tw, p.th)) return -1;
    } else {
      long long dims[4] = {Nn + p.x_lo, W, H, B};
      long long str[4] = {1, b_ld, (long long)W * b_ld, (long long)H * W * b_ld};
      int box[4] = {64, p.tw, p.th, p.tn};
      if (tc_make_map(&n.xm, b, 4, dims, str, box, 128)) return -1;
    }
    t.wcache.push_back(n);
    c = &t.wcache.back();
  }
  static bool attr_done[64] = {};
  if (tc_attr_needed(attr_done)) {
    if (cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      tc_err() = "cudaFuncSetAttribute(max dynamic smem, wgrad) failed";
      return -1;
    }
  }
  // 1x1 convolutions: [1][M][N] IS the torch layout, so the kernel accumulates straight into the (zeroed)
  // gradient and nothing is unpacked
  const int taps = ksz * ksz;
  c->p.dw_acc = taps == 1 ? dw : t.dw_acc;
  fu_launch(tc_wgrad_kernel, dim3(c->grid), dim3(kTcThreads), c->smem, stream, fu_pdl_enabled(), c->y, c->xm, c->p);
  if (cnt) { cnt->kernel_launches++; cnt->tc_kernel_launches++; }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { tc_err() = cudaGetErrorString(e); return -1; }
  if (taps > 1 && tc_unpack(t.dw_acc, dw, M, Nn, taps, stream, cnt)) { tc_err() = "weight-gradient unpack failed"; return -1; }
  return 0;
}

// halo weight-gradient kernel (3x3, W >= 32)
inline int tc_wgrad3(TcConv& t, const void* x, int x_ld, const void* dy, int dy_ld, int B, int H, int W, float* dw,
                     cudaStream_t stream, fu_counters* cnt) {
  TcConv::W3Cached* c = nullptr;
  for (auto& k : t.w3cache)
    if (k.x == x && k.dy == dy && k.x_ld == x_ld && k.dy_ld == dy_ld && k.B == B && k.H == H && k.W == W) { c = &k; break; }
  if (!c) {
    TcConv::W3Cached n;
    memset(&n, 0, sizeof(n));
    n.x = x; n.dy = dy; n.x_ld = x_ld; n.dy_ld = dy_ld; n.B = B; n.H = H; n.W = W;
    TcWgrad3Params& p = n.p;
    p.B = B; p.H = H; p.W = W; p.Cin = t.Cin; p.Cout = t.Cout;
    // K pixels per stage: the multiple of 16 in {64,48,32} with the lowest cost per image row = (padded pixels + a fixed
    // per-stage cost worth ~24 pixels: barrier round trips and 2-5 TMA loads per stage.  W = 736 used to pick 32 (no
    // padding, 23 stages of 2 K steps) and ran 1.7x slower per pixel than W = 192 with 64)
    if (t.Cin <= 32) { p.N = 32; p.cb = 32; p.tpg = 9; p.groups = 1; }
    else { p.N = t.Cin > 64 ? 128 : 64; p.cb = 64; p.tpg = 3; p.groups = 3; }
    // filter rows stacked along M (see TcWgrad3Params::mstack): Cout <= 32, one in-channel block, bf16 storage
    p.mstack = (t.Cout <= 32 && p.N == p.cb && t.Cin <= p.N && !t.split && tc_env_int("FU_TC_W3_STACK", 1) &&
                tc_env_int("FU_TC_W3_MSTACK", 1)) ? 1 : 0;
    if (p.mstack) { p.tpg = 3; p.groups = 1; }
    // (a stage holds up to 128 pixels: these kernels run at a roughly constant ~800 cycles per stage -- barrier round
    //  trips, TMA issue -- whatever the stage holds, so on the wide levels fewer, larger stages win)
    int best_tw = 64; long long best_cost = -1;
    for (int tw = tc_env_int("FU_TC_W3_MAXTW", 128); tw >= 32; tw -= 16) {
      const long long cost = (long long)((W + tw - 1) / tw) * (tw + tc_env_int("FU_TC_W3_STAGE_COST", 24));
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tw = tw; }
    }
    p.tw = best_tw; p.segs = (W + p.tw - 1) / p.tw;
    const int nrows = p.tpg == 9 ? 3 : 1;
    const size_t b_box = ((size_t)(p.tw + 2) * nrows * p.cb * 2 + 1023) / 1024 * 1024;
    p.b_stage_bytes = (unsigned)(b_box * (p.N / p.cb));
    p.stack = (p.N == p.cb && tc_env_int("FU_TC_W3_STACK", 1)) ? 1 : 0;
    p.co_tiles = (t.Cout + 127) / 128; p.ci_tiles = (t.Cin + p.N - 1) / p.N;
    const size_t stage_bytes = (p.mstack ? (size_t)4 * p.tw * 64 : (size_t)2 * p.tw * 128) + p.b_stage_bytes;
    const size_t fixed = 1024 + 8 * (2 * kTcMaxStages + 4);
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > kTcMaxStages) stages = kTcMaxStages;
    p.stages = stages;
    n.smem = fixed + (size_t)stages * stage_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long units = (long long)p.co_tiles * p.ci_tiles * p.groups;
    const long long total_kt = (long long)B * (p.mstack ? H + 2 : H) * p.segs;
    const long long max_splits = (total_kt + 7) / 8;
    long long splits = tc_pick_splits(units, max_splits, sms, tc_env_int("FU_TC_WGRAD3_WAVES", 1));
    const long long per = (total_kt + splits - 1) / splits;
    splits = (total_kt + per - 1) / per;
    p.splits = (int)splits;
    n.grid = (int)(units * splits);
    p.dw_acc = t.dw_acc;
    p.skip_epi = tc_env_int("FU_TC_WGRAD_NOEPI", 0);
    p.m64 = t.Cout <= 64 ? tc_env_int("FU_TC_M64", 2) : 0;
    p.split = t.split ? 1 : 0; p.y_lo = t.split ? dy_ld / 2 : 0; p.x_lo = t.split ? x_ld / 2 : 0;
    {
      long long dims[4] = {t.Cout + p.y_lo, W, H, B};
      long long str[4] = {1, dy_ld, (long long)W * dy_ld, (long long)H * W * dy_ld};
      int box[4] = {p.mstack ? 32 : 64, p.tw, 1, 1};
      if (tc_make_map(&n.y, dy, 4, dims, str, box, p.mstack ? 64 : 128)) return -1;
    }
    {
      long long dims[4] = {t.Cin + p.x_lo, W, H, B};
      long long str[4] = {1, x_ld, (long long)W * x_ld, (long long)H * W * x_ld};
      int box[4] = {p.cb, p.tw + 2, nrows, 1};
      if (tc_make_map(&n.xm, x, 4, dims, str, box, p.cb * 2)) return -1;
    }
    t.w3cache.push_back(n);
    c = &t.w3cache.back();
  }
  static bool attr_done[64] = {};
  if (tc_attr_needed(attr_done)) {
    if (cudaFuncSetAttribute(tc_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      tc_err() = "cudaFuncSetAttribute(max dynamic smem, wgrad3) failed";
      return -1;
    }
  }
  fu_launch(tc_wgrad3_kernel, dim3(c->grid), dim3(kTcThreads), c->smem, stream, fu_pdl_enabled(), c->y, c->xm, c->p);
  if (cnt) { cnt->kernel_launches++; cnt->tc_kernel_launches++; }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { tc_err() = cudaGetErrorString(e); return -1; }
  if (tc_unpack(t.dw_acc, dw, t.Cout, t.Cin, 9, stream, cnt)) { tc_err() = "weight-gradient unpack failed"; return -1; }
  return 0;
}

// Conv2d 3x3 / 1x1: x, dy on the same (B,H,W) grid
inline int tc_conv_wgrad(TcConv& t, const void* x, int x_ld, const void* dy, int dy_ld, int B, int H, int W, float* dw,
                         cudaStream_t stream, fu_counters* cnt) {
  if (t.kind == 0 && t.k == 3 && W >= tc_env_int("FU_TC_W3_MINW", 32) && tc_env_int("FU_TC_W3", 1))
    return tc_wgrad3(t, x, x_ld, dy, dy_ld, B, H, W, dw, stream, cnt);
  return tc_wgrad_common(t, dy, dy_ld, t.Cout, x, x_ld, t.Cin, 0, t.k, B, H, W, dw, stream, cnt);
}
// Conv2d 2x2/s2: x (B,H,W,Cin) fine, dy (B,H/2,W/2,Cout) coarse -> dw (Cout,Cin,2,2)
inline int tc_down_wgrad(TcConv& t, const void* x, int x_ld, const void* dy, int dy_ld, int B, int H, int W, float* dw,
                         cudaStream_t stream, fu_counters* cnt) {
  return tc_wgrad_common(t, dy, dy_ld, t.Cout, x, x_ld, t.Cin, 1, 2, B, H / 2, W / 2, dw, stream, cnt);
}
// ConvTranspose2d 2x2/s2: x (B,h,w,Cin) coarse, dy (B,2h,2w,Cout) fine -> dw (Cin,Cout,2,2)
inline int tc_up_wgrad(TcConv& t, const void* x, int x_ld, const void* dy, int dy_ld, int B, int h, int w, float* dw,
                       cudaStream_t stream, fu_counters* cnt) {
  return tc_wgrad_common(t, x, x_ld, t.Cin, dy, dy_ld, t.Cout, 1, 2, B, h, w, dw, stream, cnt);
}

// kernel-level test hook (fu_test_conv, impl = 1): bf16 NHWC tensors, fp32 torch-layout weights
// split != 0: parity mode -- x / dy / y_or_dx are fp32 NHWC, the operands go through their bf16 twins
inline int tc_test_conv(int mode, int B, int H, int W, int Cin, int Cout, int k, int stride, int pad, int relu,
                        const void* x, const float* w, const float* bias, void* y_or_dx, const void* dy, float* dw,
                        double* stats, cudaStream_t stream, fu_counters* cnt, int split = 0) {
  // k=3/1, stride 1: Conv2d.  k=2, stride 2: Conv2d(Cin,Cout,2,2) on x (B,H,W,Cin).  k=2, stride -2:
  // ConvTranspose2d(Cin,Cout,2,2) on x (B,H,W,Cin) -> (B,2H,2W,Cout), w in its (Cin,Cout,2,2) layout.
  const bool plain = (k == 1 || k == 3) && stride == 1 && pad == k / 2;
  const bool down = k == 2 && stride == 2 && pad == 0;
  const bool up = k == 2 && stride == -2 && pad == 0;
  if (!plain && !down && !up) { tc_err() = "tc_test_conv: unsupported geometry"; return -1; }
  TcConv t;
  char *mem = nullptr, *mem2 = nullptr;
  Bump dry, dry2;
  tc_carve(t, Cin, Cout, k, up, true, dry, dry2, split != 0);
  if (!t.enabled) { tc_err() = "tc_test_conv: shape not eligible (channels must be multiples of 32)"; return -1; }
  if (cudaMalloc(&mem, dry.off + 256) != cudaSuccess || cudaMalloc(&mem2, dry2.off + 256) != cudaSuccess) {
    tc_err() = "cudaMalloc failed";
    return -1;
  }
  Bump real, real2; real.base = mem; real2.base = mem2;
  tc_carve(t, Cin, Cout, k, up, true, real, real2, split != 0);
  int rc = 0;
  // operand descriptors: the tensors themselves, or (split) their twins
  bf16 *xt = nullptr, *dyt = nullptr;
  int x_ld = Cin, dy_ld = Cout;
  if (split) {
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    // pixel counts of the two operands: x lives on (B,H,W); dy on (B,H,W) (plain), (B,H/2,W/2) (down) or (B,2H,2W) (up)
    const long long Px = (long long)B * H * W;
    const long long Pdy = plain ? Px : (down ? (long long)B * (H / 2) * (W / 2) : 4 * Px);
    if (x && mode != 1) {
      if (cudaMalloc(&xt, (size_t)Px * Cin * 4) != cudaSuccess) { tc_err() = "cudaMalloc failed"; return -1; }
      tc_split(reinterpret_cast<const float*>(x), Cin, Cin, xt, 2 * Cin, Cin, Px, sms, stream, cnt);
      x = xt; x_ld = 2 * Cin;
    }
    if (dy && mode != 0) {
      if (cudaMalloc(&dyt, (size_t)Pdy * Cout * 4) != cudaSuccess) { tc_err() = "cudaMalloc failed"; return -1; }
      tc_split(reinterpret_cast<const float*>(dy), Cout, Cout, dyt, 2 * Cout, Cout, Pdy, sms, stream, cnt);
      dy = dyt; dy_ld = 2 * Cout;
    }
  }
  if (mode == 2) {
    cudaMemsetAsync(mem2, 0, dry2.off + 256, stream);
    cudaMemsetAsync(dw, 0, (size_t)Cin * Cout * k * k * sizeof(float), stream);   // 1x1 layers accumulate in place
    if (plain) rc = tc_conv_wgrad(t, x, x_ld, dy, dy_ld, B, H, W, dw, stream, cnt);
    else if (down) rc = tc_down_wgrad(t, x, x_ld, dy, dy_ld, B, H, W, dw, stream, cnt);
    else rc = tc_up_wgrad(t, x, x_ld, dy, dy_ld, B, H, W, dw, stream, cnt);
  } else {
    rc = tc_pack(t, w, stream, cnt);
    if (!rc && mode == 0) {
      if (plain) rc = tc_conv_forward(t, x, x_ld, y_or_dx, Cout, B, H, W, bias, relu, stats, nullptr, 0, nullptr, nullptr, 0, stream, cnt);
      else if (down) rc = tc_down_forward(t, x, x_ld, y_or_dx, Cout, B, H, W, bias, stream, cnt);
      else rc = tc_up_forward(t, x, x_ld, y_or_dx, Cout, B, H, W, bias, stream, cnt);
    } else if (!rc) {
      if (plain) rc = tc_conv_dgrad(t, dy, dy_ld, y_or_dx, Cin, B, H, W, 0, stream, cnt);
      else if (down) rc = tc_down_dgrad(t, dy, dy_ld, y_or_dx, Cin, B, H, W, 0, stream, cnt);
      else rc = tc_up_dgrad(t, dy, dy_ld, y_or_dx, Cin, B, H, W, stream, cnt);
    }
  }
  cudaError_t e = cudaStreamSynchronize(stream);
  cudaFree(mem);
  cudaFree(mem2);
  if (xt) cudaFree(xt);
  if (dyt) cudaFree(dyt);
  if (!rc && e != cudaSuccess) { tc_err() = std::string("tc_test_conv: ") + cudaGetErrorString(e); return -1; }
  return rc;
}

}  // namespace fu
