// Sample preparation and inference post-processing kernels: the callers and data formats either side of the
// U-Net.  All HBM-bound CUDA-core work; every tensor fp32 NCHW / u8 exactly as the
// reference's host code holds it.
//   prep_stats / prep_apply      dataset.py:287-293   reflect pad + per-tile z-score
//   heatmap_targets              dataset.py:295-325   Gaussian heat-map targets, sigma 2.5
//   ens_minmax / ens_combine     util.py:331-370      ensemble average, per-net heat min-max, arg-max labels
//   extract_landmarks            est_lands_csv.py:87-134 + ncc.py:12-38   masked arg-max + template NCC test
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace fu {

// numpy.pad(mode='reflect') index (the edge sample is not repeated); valid while the overshoot is < n.
__device__ __forceinline__ int reflect_idx(int j, int n) {
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return j;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over a block of <= 1024 threads; result valid in every thread.  `sh` holds 33 doubles.
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_d(v);
  __syncthreads();  // sh may still be read from a previous call
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? sh[lane] : 0.0;
    t = warp_sum_d(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// order-preserving float <-> uint32 map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---------------------------------------------------------------------------------------------
// dataset.py:287-293.  tiles (B,h,w) -> out (B,1,h+2p,w+2p): reflect pad, then (p - mean) / std with the
// unbiased standard deviation of the PADDED tile.
struct PrepArgs {
  const float* src;  // (B,h,w)
  float* out;        // (B,Hp,Wp)
  double* sums;      // (B,2): sum, sum of squares over the padded tile
  int B, h, w, pad, Hp, Wp, normalize;
  int vec;           // out 16-byte aligned and Wp % 4 == 0: float4 stores
};

__global__ void __launch_bounds__(256) prep_stats_kernel(PrepArgs a) {
  __shared__ double sh[33];
  const int b = blockIdx.y;
  const float* s = a.src + (size_t)b * a.h * a.w;
  const int n = a.Hp * a.Wp;
  double acc = 0.0, acc2 = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // four independent loads in flight before the fp64 adds
      const int j = i + k * stride;
      const int y = j / a.Wp, x = j - y * a.Wp;
      v[k] = j < n ? s[(size_t)reflect_idx(y - a.pad, a.h) * a.w + reflect_idx(x - a.pad, a.w)] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc += (double)v[k];
      acc2 += (double)v[k] * (double)v[k];
    }
  }
  acc = block_sum_d(acc, sh);
  acc2 = block_sum_d(acc2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&a.sums[2 * b], acc);
    atomicAdd(&a.sums[2 * b + 1], acc2);
  }
}

// One thread = 4 consecutive pixels of a padded row (Wp4 = ceil(Wp/4) groups per row); the tile's mean / std are
// finalised once per block.  16-byte stores when the padded row length is a multiple of 4.
__global__ void __launch_bounds__(256) prep_apply_kernel(PrepArgs a) {
  __shared__ float s_mean, s_std;
  const int b = blockIdx.y;
  const int n = a.Hp * a.Wp;
  if (a.normalize && threadIdx.x == 0) {
    const double s = a.sums[2 * b], ss = a.sums[2 * b + 1];
    const double mean = s / n;
    s_mean = (float)mean;
    s_std = (float)sqrt(fmax(ss - s * mean, 0.0) / (double)(n - 1));
  }
  __syncthreads();
  const int Wp4 = (a.Wp + 3) >> 2;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.Hp * Wp4) return;
  const int y = g / Wp4, x0 = (g - y * Wp4) << 2;
  const float* row = a.src + (size_t)b * a.h * a.w + (size_t)reflect_idx(y - a.pad, a.h) * a.w;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    v[k] = x < a.Wp ? row[reflect_idx(x - a.pad, a.w)] : 0.f;
    if (a.normalize) v[k] = __fdiv_rn(__fsub_rn(v[k], s_mean), s_std);
  }
  float* dst = a.out + (size_t)b * n + (size_t)y * a.Wp + x0;
  if (a.vec) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < a.Wp) dst[k] = v[k];
  }
}

// ---------------------------------------------------------------------------------------------
// dataset.py:295-325.  lands (B,2,L): row 0 = x (column), row 1 = y (row); +-inf marks a landmark outside
// the view (its plane stays zero).  out (B,L,H,W).  The arithmetic follows the reference's fp32 expression
// term by term (no fused multiply-add) so the planes agree to the ulp of expf.
struct HeatArgs {
  const float* lands;
  float* out;
  int B, L, H, W;
  float neg2ss;   // sigma * sigma * -2
  float norm;     // 2 * pi * sigma * sigma
  float far_r2;   // beyond this squared distance exp() is exactly 0 in fp32 (argument < -105)
  int R;          // ceil(sqrt(far_r2)) + 1: half side of the box that holds every non-zero pixel
};

__device__ __forceinline__ float gauss_px(float x, float y, float mx, float my, float neg2ss, float norm, float far_r2) {
  const float dx = __fsub_rn(x, mx), dy = __fsub_rn(y, my);
  const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  if (!(r2 <= far_r2)) return r2 != r2 ? r2 : 0.f;   // underflows to +0 (also below the smallest denormal); NaN stays NaN
  return __fdiv_rn(expf(__fdiv_rn(r2, neg2ss)), norm);
}

// The planes are zero-filled by a memset (write-bound, no instructions per pixel); this kernel then visits only the
// (2R+2)^2 bounding box of each landmark's support disc (R = ceil(sqrt(far_r2)), 37 pixels for sigma 2.5): outside
// it the fp32 value of the reference expression is exactly +0.  grid = (box blocks, B*L); a NaN coordinate
// makes the whole plane NaN in the reference, so such a plane is swept completely.
__global__ void __launch_bounds__(256) heatmap_targets_kernel(HeatArgs a) {
  const int bl = blockIdx.y;
  const int b = bl / a.L, l = bl - b * a.L;
  const float mx = a.lands[((size_t)b * 2 + 0) * a.L + l], my = a.lands[((size_t)b * 2 + 1) * a.L + l];
  if (isinf(mx) || isinf(my)) return;  // dataset.py:316: plane stays zero
  float* dst = a.out + (size_t)bl * a.H * a.W;
  if (mx != mx || my != my) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.H * a.W; i += gridDim.x * blockDim.x) dst[i] = CUDART_NAN_F;
    return;
  }
  const int side = 2 * a.R + 2;
  // box origin; coordinates far outside the image give an empty intersection (clamped before the int conversion)
  const int x0 = (int)floorf(fminf(fmaxf(mx, -1.0e6f), 1.0e6f)) - a.R;
  const int y0 = (int)floorf(fminf(fmaxf(my, -1.0e6f), 1.0e6f)) - a.R;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < side * side; i += gridDim.x * blockDim.x) {
    const int by = i / side, bx = i - by * side;
    const int y = y0 + by, x = x0 + bx;
    if (y < 0 || y >= a.H || x < 0 || x >= a.W) continue;
    const float v = gauss_px((float)x, (float)y, mx, my, a.neg2ss, a.norm, a.far_r2);
    if (v != 0.f) dst[(size_t)y * a.W + x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// util.py:331-370.  N networks' full-size outputs (B,C,H,W) / (B,L,H,W); the centre-crop window (r0,c0,h,w) is
// folded into the indexing.  Heat-maps are min-max normalised per (network, image) before averaging.
constexpr int kMaxNets = 16;
struct EnsArgs {
  const float* seg[kMaxNets];
  const float* heat[kMaxNets];
  int n_nets, B, C, L, H, W, r0, c0, h, w;
  int b0, nb;       // this launch covers images [b0, b0 + nb)
  uint32_t* mn;     // (n_nets*B) ordered-uint minima
  uint32_t* mx;     // (n_nets*B) ordered-uint maxima
  uint8_t* labels;  // (B,h,w)
  float* avg_heat;  // (B,L,h,w)
};

// The host launches min-max + combine per CHUNK of images whose heat-maps (all networks) fit in a fraction of the
// 126 MB L2, so the combine pass re-reads them from L2 instead of HBM (the min / max of a whole (network, image)
// tensor must be known before its first pixel can be normalised: two passes are inherent).
// 256 threads = 4 rows x 64 columns of the crop window; one integer division per row, none per pixel.
__global__ void __launch_bounds__(256) ens_minmax_kernel(EnsArgs a) {
  __shared__ float s_mn[8], s_mx[8];
  const int n = blockIdx.y / a.nb, b = a.b0 + (blockIdx.y - n * a.nb);
  const float* src = a.heat[n] + (size_t)b * a.L * a.H * a.W;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int rows = a.L * a.h;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (int row = blockIdx.x * 8 + ty; row < rows; row += gridDim.x * 8) {
    const int row2 = min(row + 4, rows - 1);  // second row of the pair (a repeat of a valid row at the end)
    const int l = row / a.h, y = row - l * a.h, l2 = row2 / a.h, y2 = row2 - l2 * a.h;
    const float* p = src + ((size_t)l * a.H + a.r0 + y) * a.W + a.c0;
    const float* p2 = src + ((size_t)l2 * a.H + a.r0 + y2) * a.W + a.c0;
    for (int x = tx; x < a.w; x += 256) {  // eight independent loads in flight
      float v[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = x + 64 * k < a.w ? x + 64 * k : x;
        v[k] = p[xx];
        v[4 + k] = p2[xx];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { mn = fminf(mn, v[k]); mx = fmaxf(mx, v[k]); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_mn[wid] = mn; s_mx[wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); }
    atomicMin(&a.mn[n * a.B + b], f2ord(mn));
    atomicMax(&a.mx[n * a.B + b], f2ord(mx));
  }
}

// grid = (pixel blocks, images of the chunk, L + 1): z < L normalises and averages heat-map plane z, z == L does the
// class average + arg-max.  NN > 0: network count known at compile time, so the NN loads of a pixel are issued
// together; NN == 0: any count up to kMaxNets.  The sums run over the networks in list order, as util.py:341-356.
constexpr int kEnsPix = 4;  // pixels per thread in ens_combine (strided by the block size: coalesced, 4*NN loads in flight)
template <int NN>
__global__ void __launch_bounds__(256) ens_combine_kernel(EnsArgs a) {
  __shared__ float s_lo[kMaxNets], s_den[kMaxNets];
  const int b = a.b0 + blockIdx.y;
  const int nn = NN > 0 ? NN : a.n_nets;
  const bool heat_plane = (int)blockIdx.z < a.L;
  if (heat_plane && threadIdx.x < nn) {
    const float lo = ord2f(a.mn[threadIdx.x * a.B + b]), hi = ord2f(a.mx[threadIdx.x * a.B + b]);
    s_lo[threadIdx.x] = lo;
    s_den[threadIdx.x] = __fsub_rn(hi, lo);   // util.py:351
  }
  __syncthreads();
  const int hw = a.h * a.w;
  const size_t plane = (size_t)a.H * a.W;
  const float fn = (float)nn;
  constexpr int U = NN > 0 ? NN : 1;
  int idx[kEnsPix];
  size_t pix[kEnsPix];
#pragma unroll
  for (int k = 0; k < kEnsPix; ++k) {
    idx[k] = (blockIdx.x * kEnsPix + k) * blockDim.x + threadIdx.x;
    const int i = min(idx[k], hw - 1);  // out-of-range lanes read a valid pixel and skip the store
    const int y = i / a.w, x = i - y * a.w;
    pix[k] = (size_t)(a.r0 + y) * a.W + a.c0 + x;
  }
  if (heat_plane) {
    const int l = blockIdx.z;
    const size_t base = ((size_t)b * a.L + l) * plane;
    float acc[kEnsPix];
    if (NN > 0) {
      float v[kEnsPix][U];
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k)
#pragma unroll
        for (int n = 0; n < U; ++n) v[k][n] = a.heat[n][base + pix[k]];
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k) {
        acc[k] = __fdiv_rn(__fsub_rn(v[k][0], s_lo[0]), s_den[0]);
#pragma unroll
        for (int n = 1; n < U; ++n) acc[k] = __fadd_rn(acc[k], __fdiv_rn(__fsub_rn(v[k][n], s_lo[n]), s_den[n]));
      }
    } else {
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k) {
        acc[k] = __fdiv_rn(__fsub_rn(a.heat[0][base + pix[k]], s_lo[0]), s_den[0]);
        for (int n = 1; n < nn; ++n)
          acc[k] = __fadd_rn(acc[k], __fdiv_rn(__fsub_rn(a.heat[n][base + pix[k]], s_lo[n]), s_den[n]));
      }
    }
#pragma unroll
    for (int k = 0; k < kEnsPix; ++k)
      if (idx[k] < hw) a.avg_heat[((size_t)b * a.L + l) * hw + idx[k]] = __fdiv_rn(acc[k], fn);
    return;
  }
  // avg_masks = sum over nets in list order, / num_nets, torch.max(dim=1): the first maximum wins
  float best[kEnsPix];
  int arg[kEnsPix];
#pragma unroll
  for (int k = 0; k < kEnsPix; ++k) { best[k] = 0.f; arg[k] = 0; }
  for (int c = 0; c < a.C; ++c) {
    const size_t base = ((size_t)b * a.C + c) * plane;
    float acc[kEnsPix];
    if (NN > 0) {
      float v[kEnsPix][U];
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k)
#pragma unroll
        for (int n = 0; n < U; ++n) v[k][n] = a.seg[n][base + pix[k]];
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k) {
        acc[k] = v[k][0];
#pragma unroll
        for (int n = 1; n < U; ++n) acc[k] = __fadd_rn(acc[k], v[k][n]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kEnsPix; ++k) {
        acc[k] = a.seg[0][base + pix[k]];
        for (int n = 1; n < nn; ++n) acc[k] = __fadd_rn(acc[k], a.seg[n][base + pix[k]]);
      }
    }
#pragma unroll
    for (int k = 0; k < kEnsPix; ++k) {
      const float m = __fdiv_rn(acc[k], fn);
      if (c == 0 || m > best[k]) { best[k] = m; arg[k] = c; }
    }
  }
#pragma unroll
  for (int k = 0; k < kEnsPix; ++k)
    if (idx[k] < hw) a.labels[(size_t)b * hw + idx[k]] = (uint8_t)arg[k];
}

// ---------------------------------------------------------------------------------------------
// est_lands_csv.py:87-134.  One block per (projection, landmark): arg-max of the heat-map (restricted to the
// pixels whose segmentation label equals the landmark's anatomy label, when given), then the NCC (ncc.py:12-38)
// of a D x D Gaussian template (util.py:36-48) with the window of the reflect-padded heat-map centred on the
// arg-max; the landmark is reported only when NCC >= min_ncc.  (-1,-1) = not found.
constexpr int kMaxLands = 64;
struct LandArgs {
  const float* heats;   // (P,L,h,w)
  const uint8_t* segs;  // (P,h,w) or nullptr
  int32_t label[kMaxLands];  // anatomy label per landmark, < 0: do not mask
  int32_t* out;         // (P,L,2): row, col
  float* ncc_out;       // (P,L) or nullptr; NaN when the masked arg-max found nothing
  int P, L, h, w, D;
  int vec;              // heat-map planes 16-byte aligned, seg planes 4-byte aligned, h*w % 4 == 0
  float neg2ss, norm, min_ncc;
};

__global__ void __launch_bounds__(256) extract_landmarks_kernel(LandArgs a) {
  __shared__ double sh[33];
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_arg;
  __shared__ float s_best;
  const int p = blockIdx.x / a.L, l = blockIdx.x - p * a.L;
  const float* heat = a.heats + (size_t)blockIdx.x * a.h * a.w;
  const int hw = a.h * a.w;
  const int label = a.label[l];
  const bool masked = a.segs != nullptr && label >= 0;
  const uint8_t* seg = masked ? a.segs + (size_t)p * hw : nullptr;
  float best = -CUDART_INF_F;
  int arg = INT_MAX;
  if (a.vec) {  // planes are 16-byte (heat) / 4-byte (seg) aligned and a multiple of 4 pixels: 4 pixels per load
    const float4* h4 = reinterpret_cast<const float4*>(heat);
    const uchar4* s4 = reinterpret_cast<const uchar4*>(seg);
#pragma unroll 4
    for (int q = threadIdx.x; q < (hw >> 2); q += blockDim.x) {
      float4 v = h4[q];
      if (masked) {
        const uchar4 m = s4[q];
        if ((int)m.x != label) v.x = -CUDART_INF_F;
        if ((int)m.y != label) v.y = -CUDART_INF_F;
        if ((int)m.z != label) v.z = -CUDART_INF_F;
        if ((int)m.w != label) v.w = -CUDART_INF_F;
      }
      const int i = q << 2;   // ascending index order within the thread: strict > keeps the first maximum
      if (v.x > best) { best = v.x; arg = i; }
      if (v.y > best) { best = v.y; arg = i + 1; }
      if (v.z > best) { best = v.z; arg = i + 2; }
      if (v.w > best) { best = v.w; arg = i + 3; }
    }
  } else {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      float v = heat[i];
      if (masked && (int)seg[i] != label) v = -CUDART_INF_F;
      if (v > best) { best = v; arg = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_v[wid] = best; s_i[wid] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (s_v[k] > best || (s_v[k] == best && s_i[k] < arg)) { best = s_v[k]; arg = s_i[k]; }
    s_arg = arg == INT_MAX ? 0 : arg;  // torch.argmax of an all -inf plane is 0
    s_best = best;
  }
  __syncthreads();
  arg = s_arg;
  best = s_best;
  int32_t* out = a.out + (size_t)blockIdx.x * 2;
  if (masked && best == -CUDART_INF_F) {  // est_lands_csv.py:108-109
    if (threadIdx.x == 0) {
      out[0] = -1; out[1] = -1;
      if (a.ncc_out) a.ncc_out[blockIdx.x] = CUDART_NAN_F;
    }
    return;
  }
  const int r = arg / a.w, c = arg - r * a.w;
  const int D = a.D, half = D / 2, N = D * D;
  // template value (util.py:36-48, fp32) and window value (est_lands_csv.py:94,114-117) of element k
  auto tmpl = [&](int k) -> float {
    const int i = k / D, j = k - i * D;
    const float dx = (float)(j - half), dy = (float)(i - half);
    return __fdiv_rn(expf(__fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), a.neg2ss)), a.norm);
  };
  auto roi = [&](int k) -> float {
    const int i = k / D, j = k - i * D;
    return heat[(size_t)reflect_idx(r - half + i, a.h) * a.w + reflect_idx(c - half + j, a.w)];
  };
  double st = 0.0, sr = 0.0;
  for (int k = threadIdx.x; k < N; k += blockDim.x) { st += (double)tmpl(k); sr += (double)roi(k); }
  const double mt = block_sum_d(st, sh) / N;
  const double mr = block_sum_d(sr, sh) / N;
  double tt = 0.0, rr = 0.0, tr = 0.0;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const double dt = (double)tmpl(k) - mt, dr = (double)roi(k) - mr;
    tt += dt * dt; rr += dr * dr; tr += dt * dr;
  }
  tt = block_sum_d(tt, sh);
  rr = block_sum_d(rr, sh);
  tr = block_sum_d(tr, sh);
  if (threadIdx.x == 0) {
    const double sd_t = sqrt(tt / (N - 1)), sd_r = sqrt(rr / (N - 1));
    const float ncc = (float)(tr / ((double)N * (sd_t * sd_r) + 1.0e-8));
    const bool reject = ncc < a.min_ncc;  // est_lands_csv.py:119-120 (a NaN score is kept, as there)
    out[0] = reject ? -1 : r;
    out[1] = reject ? -1 : c;
    if (a.ncc_out) a.ncc_out[blockIdx.x] = ncc;
  }
}

}  // namespace fu
