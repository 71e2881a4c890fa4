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
};

__global__ void __launch_bounds__(256) prep_stats_kernel(PrepArgs a) {
  __shared__ double sh[33];
  const int b = blockIdx.y;
  const float* s = a.src + (size_t)b * a.h * a.w;
  const int n = a.Hp * a.Wp;
  double acc = 0.0, acc2 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / a.Wp, x = i - y * a.Wp;
    const double v = (double)s[(size_t)reflect_idx(y - a.pad, a.h) * a.w + reflect_idx(x - a.pad, a.w)];
    acc += v;
    acc2 += v * v;
  }
  acc = block_sum_d(acc, sh);
  acc2 = block_sum_d(acc2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&a.sums[2 * b], acc);
    atomicAdd(&a.sums[2 * b + 1], acc2);
  }
}

__global__ void __launch_bounds__(256) prep_apply_kernel(PrepArgs a) {
  const int b = blockIdx.y;
  const int n = a.Hp * a.Wp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int y = i / a.Wp, x = i - y * a.Wp;
  float v = a.src[(size_t)b * a.h * a.w + (size_t)reflect_idx(y - a.pad, a.h) * a.w + reflect_idx(x - a.pad, a.w)];
  if (a.normalize) {
    const double s = a.sums[2 * b], ss = a.sums[2 * b + 1];
    const double mean = s / n;
    const double var = fmax(ss - s * mean, 0.0) / (double)(n - 1);
    v = __fdiv_rn(__fsub_rn(v, (float)mean), (float)sqrt(var));
  }
  a.out[(size_t)b * n + i] = v;
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
};

__global__ void __launch_bounds__(256) heatmap_targets_kernel(HeatArgs a) {
  const int bl = blockIdx.y;
  const int b = bl / a.L, l = bl - b * a.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.H * a.W) return;
  const float mx = a.lands[((size_t)b * 2 + 0) * a.L + l], my = a.lands[((size_t)b * 2 + 1) * a.L + l];
  float v = 0.f;
  if (!isinf(mx) && !isinf(my)) {
    const int y = i / a.W, x = i - y * a.W;
    const float dx = __fsub_rn((float)x, mx), dy = __fsub_rn((float)y, my);
    const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    v = __fdiv_rn(expf(__fdiv_rn(r2, a.neg2ss)), a.norm);
  }
  a.out[(size_t)bl * a.H * a.W + i] = v;
}

// ---------------------------------------------------------------------------------------------
// util.py:331-370.  N networks' full-size outputs (B,C,H,W) / (B,L,H,W); the centre-crop window (r0,c0,h,w) is
// folded into the indexing.  Heat-maps are min-max normalised per (network, image) before averaging.
constexpr int kMaxNets = 16;
struct EnsArgs {
  const float* seg[kMaxNets];
  const float* heat[kMaxNets];
  int n_nets, B, C, L, H, W, r0, c0, h, w;
  uint32_t* mn;     // (n_nets*B) ordered-uint minima
  uint32_t* mx;     // (n_nets*B) ordered-uint maxima
  uint8_t* labels;  // (B,h,w)
  float* avg_heat;  // (B,L,h,w)
};

__global__ void __launch_bounds__(256) ens_minmax_kernel(EnsArgs a) {
  __shared__ float s_mn[8], s_mx[8];
  const int nb = blockIdx.y;
  const int n = nb / a.B, b = nb - n * a.B;
  const float* src = a.heat[n] + (size_t)b * a.L * a.H * a.W;
  const int hw = a.h * a.w, tot = a.L * hw;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
    const int l = i / hw, r = i - l * hw;
    const int y = r / a.w, x = r - y * a.w;
    const float v = src[((size_t)l * a.H + a.r0 + y) * a.W + a.c0 + x];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
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
    atomicMin(&a.mn[nb], f2ord(mn));
    atomicMax(&a.mx[nb], f2ord(mx));
  }
}

__global__ void __launch_bounds__(256) ens_combine_kernel(EnsArgs a) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int hw = a.h * a.w;
  if (i >= hw) return;
  const int y = i / a.w, x = i - y * a.w;
  const size_t pix = (size_t)(a.r0 + y) * a.W + a.c0 + x;
  const float fn = (float)a.n_nets;
  // avg_masks = sum over nets in list order, / num_nets, torch.max(dim=1): the first maximum wins
  float best = 0.f;
  int arg = 0;
  for (int c = 0; c < a.C; ++c) {
    const size_t off = ((size_t)b * a.C + c) * a.H * a.W + pix;
    float acc = a.seg[0][off];
    for (int n = 1; n < a.n_nets; ++n) acc = __fadd_rn(acc, a.seg[n][off]);
    acc = __fdiv_rn(acc, fn);
    if (c == 0 || acc > best) { best = acc; arg = c; }
  }
  a.labels[(size_t)b * hw + i] = (uint8_t)arg;
  if (a.avg_heat) {
    for (int l = 0; l < a.L; ++l) {
      const size_t off = ((size_t)b * a.L + l) * a.H * a.W + pix;
      float acc = 0.f;
      for (int n = 0; n < a.n_nets; ++n) {
        const float lo = ord2f(a.mn[n * a.B + b]), hi = ord2f(a.mx[n * a.B + b]);
        const float v = __fdiv_rn(__fsub_rn(a.heat[n][off], lo), __fsub_rn(hi, lo));
        acc = n == 0 ? v : __fadd_rn(acc, v);
      }
      a.avg_heat[((size_t)b * a.L + l) * hw + i] = __fdiv_rn(acc, fn);
    }
  }
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
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    float v = heat[i];
    if (masked && (int)seg[i] != label) v = -CUDART_INF_F;
    if (v > best) { best = v; arg = i; }
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
