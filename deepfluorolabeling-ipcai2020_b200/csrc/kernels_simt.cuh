// CUDA-core kernels of the engine (sm_100a).  They are (a) the whole parity-mode
// (fp32) path, (b) in throughput mode every op that is HBM-bound or too thin for
// tensor cores (BN statistics/apply, pooling, heads, C_in=1 first conv, ...), and
// (c) the in-library cross-check for the tcgen05 kernels in kernels_tc.cuh.
// Layout: NHWC, channel-contiguous, arbitrary pixel stride `ld` (so a tensor can
// be a channel slice of a wider concat buffer without a copy, unet.py:257).
#pragma once
#include "common.cuh"

namespace fu {

// --------------------------------------------------------------------------
// Implicit-GEMM convolution, generic in (k, stride, pad) and in its epilogue.
//   Y[m, col] = sum_{tap, ci} X[pix(m) @ tap, ci] * Wp[tap][ci][col]
// covers conv3x3/pad1, conv1x1, conv2x2/s2, their data gradients (with
// transposed/flipped packings of the weights) and, with the pixel-shuffle
// store, ConvTranspose2d(k=2,s=2) and the data gradient of conv2x2/s2.
// --------------------------------------------------------------------------
struct ConvArgs {
  const void* x; int x_ld;
  void* y; int y_ld;
  const float* w;      // packed [taps][Cin][Npad], zero padded columns
  const float* bias;   // nullable; indexed by col % bias_mod
  int bias_mod;
  int B, Hi, Wi, Cin;
  int Ho, Wo, N, Npad;
  int KH, KW, stride, pad;
  int relu, accumulate;
  int shuffle;         // 1: col = (a*2+b)*Cout + co -> y[n, 2oh+a, 2ow+b, co]  (N = 4*Cout)
  int nchw_out;        // 1: y is fp32 NCHW (B,N,Ho,Wo)
  int vec_out;         // 4 consecutive columns may be stored as one vector
  const void* t; int t_ld; const float* bn_a; const float* bn_b;  // + bn_a[col]*t[m,col] + bn_b[col]
  double* stat;        // nullable; [2*N]: per-column sum / sum of squares of the stored value
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) igemm_simt_kernel(const ConvArgs p) {
  pdl_wait(); pdl_trigger();
  constexpr int BM = 128, BN = 64, BK = 16, AP = BM + 4;
  __shared__ __align__(16) float As[BK][AP];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const long long M = (long long)p.B * p.Ho * p.Wo;
  const long long m_base = (long long)blockIdx.x * BM;
  const int n_base = blockIdx.y * BN;
  const T* xp = reinterpret_cast<const T*>(p.x);

  const int a_k = (tid & 3) * 4;
  const int a_m = tid >> 2;
  int an[2], aoh[2], aow[2];
  bool amv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    long long m = m_base + a_m + i * 64;
    amv[i] = m < M;
    long long mm = amv[i] ? m : 0;
    aow[i] = (int)(mm % p.Wo);
    long long t = mm / p.Wo;
    aoh[i] = (int)(t % p.Ho);
    an[i] = (int)(t / p.Ho);
  }
  const int b_k = tid >> 4;
  const int b_n = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int kchunks = (p.Cin + BK - 1) / BK;
  const int total = p.KH * p.KW * kchunks;
  float4 ra[2], rb;

  auto gload = [&](int it) {
    const int tap = it / kchunks;
    const int c0 = (it - tap * kchunks) * BK;
    const int kh = tap / p.KW, kw = tap - kh * p.KW;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ih = aoh[i] * p.stride + kh - p.pad, iw = aow[i] * p.stride + kw - p.pad;
      if (amv[i] && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi) {
        const int c = c0 + a_k;
        const T* src = xp + ((long long)(an[i] * p.Hi + ih) * p.Wi + iw) * p.x_ld + c;
        if (VEC) {
          if (c < p.Cin) v = ld4(src);
        } else {
          if (c < p.Cin) v.x = ld1(src);
          if (c + 1 < p.Cin) v.y = ld1(src + 1);
          if (c + 2 < p.Cin) v.z = ld1(src + 2);
          if (c + 3 < p.Cin) v.w = ld1(src + 3);
        }
      }
      ra[i] = v;
    }
    const int c = c0 + b_k;
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < p.Cin && n_base + b_n < p.Npad)
      rb = *reinterpret_cast<const float4*>(p.w + ((long long)tap * p.Cin + c) * p.Npad + n_base + b_n);
  };

  gload(0);
  for (int it = 0; it < total; ++it) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[a_k + 0][a_m + i * 64] = ra[i].x;
      As[a_k + 1][a_m + i * 64] = ra[i].y;
      As[a_k + 2][a_m + i * 64] = ra[i].z;
      As[a_k + 3][a_m + i * 64] = ra[i].w;
    }
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = rb;
    __syncthreads();
    if (it + 1 < total) gload(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int col0 = n_base + tx * 4;
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  float bias[4] = {0.f, 0.f, 0.f, 0.f}, bna[4] = {0.f, 0.f, 0.f, 0.f}, bnb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = col0 + j;
    if (col < p.N) {
      if (p.bias) bias[j] = p.bias[col % p.bias_mod];
      if (p.t) { bna[j] = p.bn_a[col]; bnb[j] = p.bn_b[col]; }
    }
  }
  T* yp = reinterpret_cast<T*>(p.y);
  const T* tp = reinterpret_cast<const T*>(p.t);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m_base + ty * 8 + i;
    if (m >= M || col0 >= p.N) continue;
    const int ow = (int)(m % p.Wo);
    const long long tt = m / p.Wo;
    const int oh = (int)(tt % p.Ho);
    const int n = (int)(tt / p.Ho);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (p.t) {
      const T* ts = tp + m * p.t_ld + col0;
      if (p.vec_out) {
        float4 tv = ld4(ts);
        v[0] += bna[0] * tv.x + bnb[0]; v[1] += bna[1] * tv.y + bnb[1];
        v[2] += bna[2] * tv.z + bnb[2]; v[3] += bna[3] * tv.w + bnb[3];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col0 + j < p.N) v[j] += bna[j] * ld1(ts + j) + bnb[j];
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.nchw_out) {
      float* yf = reinterpret_cast<float*>(p.y);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < p.N) yf[(((long long)n * p.N + col0 + j) * p.Ho + oh) * p.Wo + ow] = v[j];
      continue;
    }
    T* dst;
    if (p.shuffle) {
      const int cout = p.N >> 2;
      const int ab = col0 / cout, co = col0 - ab * cout;   // 4 consecutive columns share ab (cout % 4 == 0)
      const int a = ab >> 1, b = ab & 1;
      dst = yp + ((long long)(n * (2 * p.Ho) + 2 * oh + a) * (2 * p.Wo) + 2 * ow + b) * p.y_ld + co;
    } else {
      dst = yp + m * p.y_ld + col0;
    }
    if (p.vec_out) {
      if (p.accumulate) {
        float4 o = ld4(dst);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
      st4(dst, make_float4(v[0], v[1], v[2], v[3]));
#pragma unroll
      for (int j = 0; j < 4; ++j) { float r = rnd(v[j], dst); cs[j] += r; cq[j] += r * r; }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (col0 + j < p.N) {
          if (p.accumulate) v[j] += ld1(dst + j);
          st1(dst + j, v[j]);
          float r = rnd(v[j], dst); cs[j] += r; cq[j] += r * r;
        }
      }
    }
  }
  if (p.stat) {
    float* red = &As[0][0];   // 2 x [16][64]
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[ty * 64 + tx * 4 + j] = cs[j];
      red[1024 + ty * 64 + tx * 4 + j] = cq[j];
    }
    __syncthreads();
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[which * 1024 + r * 64 + c];
      const int col = n_base + c;
      if (col < p.N) atomicAdd(&p.stat[which * p.N + col], (double)s);
    }
  }
}

// --------------------------------------------------------------------------
// First-layer convolutions (C_in = in_channels, 1 in every reference script): K is 9 or 1, so
// tensor cores are pointless and the op is bound by writing the C_out-channel output once.
// One thread = one pixel x 8 output channels; weights live in shared memory.
// y = [relu](conv_k(x) + bias) [+ bn_a*t + bn_b], optional per-channel sum / sum-of-squares.
// --------------------------------------------------------------------------
struct SmallCinArgs {
  const void* x; int x_ld; int Cin;
  void* y; int y_ld; int Cout;
  const float* w;        // torch layout (Cout, Cin, k, k)
  const float* bias;
  int B, H, W, k, pad, relu;
  const void* t; int t_ld; const float* bn_a; const float* bn_b;
  double* stat;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_small_cin_kernel(const SmallCinArgs p) {
  pdl_wait(); pdl_trigger();
  extern __shared__ float sm_w[];            // [taps*Cin][Cout] + [Cout] bias + 2*[Cout] stats
  const int taps = p.k * p.k, KK = taps * p.Cin;
  float* sm_b = sm_w + KK * p.Cout;
  float* sm_s = sm_b + p.Cout;
  for (int i = threadIdx.x; i < KK * p.Cout; i += blockDim.x) {
    const int co = i % p.Cout, kk = i / p.Cout;       // kk = tap*Cin + ci
    const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
    sm_w[i] = p.w[((long long)co * p.Cin + ci) * taps + tap];
  }
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
    sm_b[i] = p.bias ? p.bias[i] : 0.f;
    sm_s[i] = 0.f; sm_s[p.Cout + i] = 0.f;
  }
  __syncthreads();
  const int groups = p.Cout >> 3;
  const long long P = (long long)p.B * p.H * p.W;
  const long long total = P * groups;
  const T* xp = reinterpret_cast<const T*>(p.x);
  const T* tp = reinterpret_cast<const T*>(p.t);
  T* yp = reinterpret_cast<T*>(p.y);
  float cs[8], cq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;
    const int w = (int)(pix % p.W);
    const long long r = pix / p.W;
    const int h = (int)(r % p.H);
    const int n = (int)(r / p.H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sm_b[g * 8 + j];
    for (int tap = 0; tap < taps; ++tap) {
      const int ih = h + tap / p.k - p.pad, iw = w + tap % p.k - p.pad;
      if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
      const T* src = xp + ((long long)(n * p.H + ih) * p.W + iw) * p.x_ld;
      for (int ci = 0; ci < p.Cin; ++ci) {
        const float xv = ld1(src + ci);
        const float4 w0 = *reinterpret_cast<const float4*>(sm_w + (tap * p.Cin + ci) * p.Cout + g * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sm_w + (tap * p.Cin + ci) * p.Cout + g * 8 + 4);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
        acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
    if (p.t) {
      const float4 t0 = ld4(tp + pix * p.t_ld + g * 8), t1 = ld4(tp + pix * p.t_ld + g * 8 + 4);
      const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += p.bn_a[g * 8 + j] * tv[j] + p.bn_b[g * 8 + j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    T* dst = yp + pix * p.y_ld + g * 8;
    st4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(dst + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float v = rnd(acc[j], dst); cs[j] += v; cq[j] += v * v; }
  }
  if (p.stat) {
    // lanes with equal (lane % groups) hold the same channels (groups is a power of two <= 32)
    for (int o = groups; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
        cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], o);
      }
    }
    // deterministic block reduction (fixed order over the 8 warps): float atomics here would perturb the
    // BN statistics by ~1e-7 run to run, which bf16 rounding downstream amplifies
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    float* part = sm_s + 2 * p.Cout;            // [8 warps][2*Cout]
    if (lane < groups) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        part[wrp * 2 * p.Cout + lane * 8 + j] = cs[j];
        part[wrp * 2 * p.Cout + p.Cout + lane * 8 + j] = cq[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) {
      float v = 0.f;
      for (int w2 = 0; w2 < 8; ++w2) v += part[w2 * 2 * p.Cout + i];
      atomicAdd(&p.stat[i], (double)v);
    }
  }
}

// dW(Cout,Cin,k,k) += sum_pixels x[p@tap, ci] * dy[p, co]  for tiny Cin (first layer): reads dy once.
struct SmallCinWgradArgs {
  const void* x; int x_ld; int Cin;
  const void* dy; int dy_ld; int Cout;
  int B, H, W, k, pad;
  float* dw;
};

template <typename T, int KK>   // KK = taps*Cin accumulators per thread (x4 channels)
__global__ void __launch_bounds__(256) wgrad_small_cin_kernel(const SmallCinWgradArgs p) {
  pdl_wait(); pdl_trigger();
  __shared__ float red[256 * 4];
  const int cg = p.Cout >> 2;                 // channel groups of 4
  const int g = threadIdx.x % cg, lane_p = threadIdx.x / cg, rows = 256 / cg;
  const int taps = p.k * p.k;
  const long long P = (long long)p.B * p.H * p.W;
  const T* xp = reinterpret_cast<const T*>(p.x);
  const T* dyp = reinterpret_cast<const T*>(p.dy);
  float acc[KK][4];
#pragma unroll
  for (int i = 0; i < KK; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  for (long long pix = (long long)blockIdx.x * rows + lane_p; pix < P; pix += (long long)gridDim.x * rows) {
    const float4 d = ld4(dyp + pix * p.dy_ld + g * 4);
    const int w = (int)(pix % p.W);
    const long long r = pix / p.W;
    const int h = (int)(r % p.H);
    const int n = (int)(r / p.H);
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
      const int ih = h + tap / p.k - p.pad, iw = w + tap % p.k - p.pad;
      float xv = 0.f;
      if (tap < taps && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
        xv = ld1(xp + ((long long)(n * p.H + ih) * p.W + iw) * p.x_ld + ci);
      acc[kk][0] = fmaf(xv, d.x, acc[kk][0]); acc[kk][1] = fmaf(xv, d.y, acc[kk][1]);
      acc[kk][2] = fmaf(xv, d.z, acc[kk][2]); acc[kk][3] = fmaf(xv, d.w, acc[kk][3]);
    }
  }
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    __syncthreads();
    red[threadIdx.x * 4 + 0] = acc[kk][0]; red[threadIdx.x * 4 + 1] = acc[kk][1];
    red[threadIdx.x * 4 + 2] = acc[kk][2]; red[threadIdx.x * 4 + 3] = acc[kk][3];
    __syncthreads();
    const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
    if (threadIdx.x < p.Cout && tap < taps) {
      const int co = threadIdx.x, gg = co >> 2, j = co & 3;
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += red[(r * cg + gg) * 4 + j];
      atomicAdd(p.dw + ((long long)co * p.Cin + ci) * taps + tap, s);
    }
  }
}

// --------------------------------------------------------------------------
// C_in = 1 fast path (every reference script: in_channels = 1, unet.py:41, train.py:313).  The generic
// first-layer kernels above issue one scalar global load and two shared-memory weight loads per 8 FMAs and
// run at 5-10 % of the HBM rate of their 32-channel operand.  Here a thread owns 8 output channels for the
// whole kernel and keeps their K*K weights in registers, walks 4-row pixel columns whose input window sits in
// registers, and lanes map to (pixel, channel group) with the group fastest so a warp's 16-byte vectors form
// one contiguous 512-byte segment of the NHWC tensor.
// --------------------------------------------------------------------------
template <typename T> struct Ld8;
template <> struct Ld8<float> {
  struct Raw { float4 a, b; };          // 8 elements as loaded (kept packed while several loads are in flight)
  static __device__ __forceinline__ Raw ldraw(const float* p) {
    Raw r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
  }
  static __device__ __forceinline__ Raw zero() { Raw r; r.a = r.b = make_float4(0.f, 0.f, 0.f, 0.f); return r; }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Ld8<bf16> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw ldraw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void ld(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(bf16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// rows per thread and loop trip
__host__ __device__ constexpr int cin1_tile_h(int K) { return 4; }
// dynamic shared memory of the two kernels below: block-reduction scratch + (3x3 only) the staged input halo
inline size_t cin1_smem_bytes(int K, int Cout, bool wgrad) {
  const int TW = 256 / (Cout >> 3), PAD = K / 2;
  const size_t halo = K == 3 ? (size_t)(((TW + 2 * PAD) * (cin1_tile_h(K) + 2 * PAD) + 7) & ~7) : 0;
  return sizeof(float) * (halo + (wgrad ? (size_t)8 * K * K * Cout : (size_t)8 * 2 * Cout));
}
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldg1(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldg(reinterpret_cast<const unsigned short*>(p))));
}

// The (TH+2*PAD) x K input window of a column of TH pixels, zero padded.  1x1: straight from global memory
// (the single-channel input is tiny and L1/L2 resident), no barrier in the pixel loop.  3x3: the block's halo
// tile is staged in shared memory first (18 scalar global loads per thread measured slower than the two
// barriers per tile: 70 vs 53 us forward, 116 vs 80 us weight gradient at 32x192x192).
template <typename T, int K>
__device__ __forceinline__ void cin1_window(const T* xp, int x_ld, float* sx, int n, int h0, int w0, int slot, int TW,
                                            int H, int W, float (&xw)[cin1_tile_h(K) + 2 * (K / 2)][K]) {
  constexpr int PAD = K / 2, XH = cin1_tile_h(K) + 2 * PAD;
  if constexpr (K == 3) {
    const int XW = TW + 2 * PAD;
    __syncthreads();                                   // the previous tile's readers are done
    for (int i = threadIdx.x; i < XW * XH; i += blockDim.x) {
      const int r = i / XW, c = i - r * XW;
      const int ih = h0 + r - PAD, iw = w0 + c - PAD;
      sx[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? ldg1(xp + ((long long)(n * H + ih) * W + iw) * x_ld) : 0.f;
    }
    __syncthreads();                                   // (the window is read from sx at the point of use)
  } else {
    const int wc = w0 + slot;
#pragma unroll
    for (int r = 0; r < XH; ++r) {
      const int ih = h0 + r - PAD;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int iw = wc + c - PAD;
        xw[r][c] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? ldg1(xp + ((long long)(n * H + ih) * W + iw) * x_ld) : 0.f;
      }
    }
  }
}

// (3x3, measured and not kept: fetching the NEXT tile's halo into registers before working on the current one -- the
//  global-load latency between the two barriers is not what bounds these kernels: forward 53.7 -> 55.7 us, weight gradient
//  75.7 -> 80.3 us at 32 x 192 x 192.)
// y = [relu](conv_K(x) + bias) [+ bn_a*t + bn_b], optional per-channel sum / sum of squares (SmallCinArgs, Cin == 1)
template <typename T, int K>
__global__ void __launch_bounds__(256, K == 1 ? 3 : 2) conv_cin1_kernel(const SmallCinArgs p) {
  pdl_wait(); pdl_trigger();
  constexpr int KK = K * K, TH = cin1_tile_h(K);
  extern __shared__ __align__(16) float cin1_sm[];
  const int groups = p.Cout >> 3, TW = 256 / groups;
  float* sx = cin1_sm;                                    // 3x3: [TH+2][TW+2] input halo
  float* part = cin1_sm + (K == 3 ? (((TW + 2) * (TH + 2) + 7) & ~7) : 0);   // [8 warps][2*Cout]
  const int g = threadIdx.x % groups, slot = threadIdx.x / groups;
  float w[KK][8], bias[8], ba[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = g * 8 + j;
#pragma unroll
    for (int t = 0; t < KK; ++t) w[t][j] = p.w[(long long)co * KK + t];
    bias[j] = (p.bias ? p.bias[co] : 0.f) + (p.t ? p.bn_b[co] : 0.f);
    ba[j] = p.t ? p.bn_a[co] : 0.f;
  }
  const T* xp = reinterpret_cast<const T*>(p.x);
  const T* tp = reinterpret_cast<const T*>(p.t);
  T* yp = reinterpret_cast<T*>(p.y);
  float cs[8], cq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
  const int tiles_w = (p.W + TW - 1) / TW, tiles_h = (p.H + TH - 1) / TH;
  const int total = tiles_w * tiles_h * p.B;
  auto tile_origin = [&](int tile, int& n, int& h0, int& w0) {
    n = tile / (tiles_w * tiles_h);
    const int rem = tile - n * (tiles_w * tiles_h);
    h0 = (rem / tiles_w) * TH; w0 = (rem % tiles_w) * TW;
  };
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int n, h0, w0;
    tile_origin(tile, n, h0, w0);
    const int wc = w0 + slot;
    float xw[TH + 2 * (K / 2)][K];
    cin1_window<T, K>(xp, p.x_ld, sx, n, h0, w0, slot, TW, p.H, p.W, xw);
    // (the second epilogue operand of all rows is requested before any row is worked on)
    // (1x1 only: the 3x3 variant holds 72 weights in registers and is never given a second operand by the engine)
    typename Ld8<T>::Raw traw[K == 1 ? TH : 1];
    if (K == 1 && p.t) {
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        const int h = h0 + r;
        traw[r] = (h < p.H && wc < p.W) ? Ld8<T>::ldraw(tp + (((long long)n * p.H + h) * p.W + wc) * p.t_ld + g * 8) : Ld8<T>::zero();
      }
    }
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      const int h = h0 + r;
      if (h < p.H && wc < p.W) {
        const long long pix = ((long long)n * p.H + h) * p.W + wc;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bias[j];
#pragma unroll
        for (int t = 0; t < KK; ++t) {
          const float xv = K == 3 ? sx[(r + t / K) * (TW + 2) + slot + t % K] : xw[r + t / K][t % K];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, w[t][j], acc[j]);
        }
        if (p.t) {
          float tv[8];
          if constexpr (K == 1) Ld8<T>::unpack(traw[r], tv);
          else Ld8<T>::ld(tp + pix * p.t_ld + g * 8, tv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(ba[j], tv[j], acc[j]);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
        }
        T* dst = yp + pix * p.y_ld + g * 8;
        Ld8<T>::st(dst, acc);
        if (p.stat) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float v = rnd(acc[j], dst); cs[j] += v; cq[j] = fmaf(v, v, cq[j]); }
        }
      }
    }
  }
  if (p.stat) {
    // lanes with equal (lane % groups) hold the same channels (groups is a power of two <= 32); the block
    // reduction runs in a fixed order so the BN statistics are reproducible run to run
    for (int o = groups; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
        cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], o);
      }
    }
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    __syncthreads();
    if (lane < groups) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        part[wrp * 2 * p.Cout + lane * 8 + j] = cs[j];
        part[wrp * 2 * p.Cout + p.Cout + lane * 8 + j] = cq[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) {
      float v = 0.f;
      for (int w2 = 0; w2 < 8; ++w2) v += part[w2 * 2 * p.Cout + i];
      atomicAdd(&p.stat[i], (double)v);
    }
  }
}

// dW(Cout,1,K,K) += sum_pixels x[p@tap] * dy[p, co]   (SmallCinWgradArgs, Cin == 1): reads dy once
template <typename T, int K>
__global__ void __launch_bounds__(256, K == 1 ? 3 : 2) wgrad_cin1_kernel(const SmallCinWgradArgs p) {
  pdl_wait(); pdl_trigger();
  constexpr int KK = K * K, TH = cin1_tile_h(K);
  extern __shared__ __align__(16) float cin1_sm[];
  const int groups = p.Cout >> 3, TW = 256 / groups;
  float* sx = cin1_sm;                                    // 3x3: [TH+2][TW+2] input halo
  float* part = cin1_sm + (K == 3 ? (((TW + 2) * (TH + 2) + 7) & ~7) : 0);   // [8 warps][KK][Cout]
  const int g = threadIdx.x % groups, slot = threadIdx.x / groups;
  const T* xp = reinterpret_cast<const T*>(p.x);
  const T* dyp = reinterpret_cast<const T*>(p.dy);
  float acc[KK][8];
#pragma unroll
  for (int t = 0; t < KK; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int tiles_w = (p.W + TW - 1) / TW, tiles_h = (p.H + TH - 1) / TH;
  const int total = tiles_w * tiles_h * p.B;
  auto tile_origin = [&](int tile, int& n, int& h0, int& w0) {
    n = tile / (tiles_w * tiles_h);
    const int rem = tile - n * (tiles_w * tiles_h);
    h0 = (rem / tiles_w) * TH; w0 = (rem % tiles_w) * TW;
  };
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int n, h0, w0;
    tile_origin(tile, n, h0, w0);
    const int wc = w0 + slot;
    float d[TH][8];
#pragma unroll
    for (int r = 0; r < TH; ++r) {       // all of the tile's dy loads in flight before the barrier / the FMAs
      const int h = h0 + r;
      if (h < p.H && wc < p.W) {
        Ld8<T>::ld(dyp + (((long long)n * p.H + h) * p.W + wc) * p.dy_ld + g * 8, d[r]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[r][j] = 0.f;
      }
    }
    float xw[TH + 2 * (K / 2)][K];
    cin1_window<T, K>(xp, p.x_ld, sx, n, h0, w0, slot, TW, p.H, p.W, xw);
#pragma unroll
    for (int r = 0; r < TH; ++r) {
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        const float xv = K == 3 ? sx[(r + t / K) * (TW + 2) + slot + t % K] : xw[r + t / K][t % K];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xv, d[r][j], acc[t][j]);
      }
    }
  }
  for (int o = groups; o < 32; o <<= 1) {
#pragma unroll
    for (int t = 0; t < KK; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], o);
  }
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  __syncthreads();
  if (lane < groups) {
#pragma unroll
    for (int t = 0; t < KK; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) part[(wrp * KK + t) * p.Cout + lane * 8 + j] = acc[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KK * p.Cout; i += blockDim.x) {
    const int t = i / p.Cout, co = i - t * p.Cout;
    float v = 0.f;
    for (int w2 = 0; w2 < 8; ++w2) v += part[(w2 * KK + t) * p.Cout + co];
    atomicAdd(p.dw + (long long)co * KK + t, v);
  }
}

// --------------------------------------------------------------------------
// Fused heads (unet.py:176-191): per pixel  logits = Wseg feat ; seg = softmax(logits) ;
// heat = W2 (W1 [feat ; logits]) = W21 [feat ; logits] with W21 = W2 W1 folded once per block
// (no non-linearity sits between the two 1x1 convs, unet.py:148-153).  One thread = TWO pixels so that every
// 16-byte shared-memory weight load feeds 8 FMAs (the first version issued one LDS per FMA and was
// LDS-bound at 0.25 ms; the op's HBM floor is ~30 us).  seg / heat / (optional) logits are written as fp32
// NCHW boundary tensors; the logits also go, in the storage type, into the head's concat buffer.
// --------------------------------------------------------------------------
template <int CF, int NC, int NF, int NL>
struct HeadsDims {
  static constexpr int NCAT = CF + NC;
  static constexpr int NCATP = (NCAT + 3) / 4 * 4;
  static constexpr int NCP = (NC + 3) / 4 * 4;
  static constexpr int NLp = NL > 0 ? NL : 1;
};

// W21[l][j] = sum_m W2[l][m] W1[m][j], rows padded with zeros to NCATP
template <int CF, int NC, int NF, int NL>
__device__ __forceinline__ void heads_fold_w21(const float* w1, const float* w2, float* s_w21) {
  typedef HeadsDims<CF, NC, NF, NL> D;
  if (NL > 0) {
    for (int i = threadIdx.x; i < NL * D::NCATP; i += blockDim.x) {
      const int l = i / D::NCATP, j = i - l * D::NCATP;
      float a = 0.f;
      if (j < D::NCAT)
        for (int m = 0; m < NF; ++m) a = fmaf(w2[l * NF + m], w1[m * D::NCAT + j], a);
      s_w21[i] = a;
    }
  }
}

// logits of two pixels from their feature vectors (4 independent FMA chains per pixel)
template <int CF, int NC>
__device__ __forceinline__ void heads_logits2(const float* q_wseg, const float (&f)[2][CF], float (&lg)[2][(NC + 3) / 4 * 4]) {
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int c = 0; c < CF; c += 4) {
      const float4 w = *reinterpret_cast<const float4*>(q_wseg + k * CF + c);
      a0.x = fmaf(w.x, f[0][c], a0.x); a0.y = fmaf(w.y, f[0][c + 1], a0.y);
      a0.z = fmaf(w.z, f[0][c + 2], a0.z); a0.w = fmaf(w.w, f[0][c + 3], a0.w);
      a1.x = fmaf(w.x, f[1][c], a1.x); a1.y = fmaf(w.y, f[1][c + 1], a1.y);
      a1.z = fmaf(w.z, f[1][c + 2], a1.z); a1.w = fmaf(w.w, f[1][c + 3], a1.w);
    }
    lg[0][k] = (a0.x + a0.y) + (a0.z + a0.w);
    lg[1][k] = (a1.x + a1.y) + (a1.z + a1.w);
  }
#pragma unroll
  for (int k = NC; k < (NC + 3) / 4 * 4; ++k) { lg[0][k] = 0.f; lg[1][k] = 0.f; }
}

template <typename T, int CF, int NC, int NF, int NL>
__global__ void __launch_bounds__(128) heads_fwd_fused_kernel(const T* feat, int ld, const float* wseg, const float* w1,
                                                              const float* w2, T* logits_nhwc, float* seg,
                                                              float* logits_out, float* heat, int B, long long HW,
                                                              int do_softmax) {
  pdl_wait(); pdl_trigger();
  typedef HeadsDims<CF, NC, NF, NL> D;
  static_assert(CF % 4 == 0, "feature channels must be a multiple of 4");
  __shared__ __align__(16) float s_wseg[NC * CF];
  __shared__ __align__(16) float s_w21[NL > 0 ? NL * D::NCATP : 4];
  for (int i = threadIdx.x; i < NC * CF; i += blockDim.x) s_wseg[i] = wseg[i];
  heads_fold_w21<CF, NC, NF, NL>(w1, w2, s_w21);
  __syncthreads();
  const long long P = (long long)B * HW;
  const long long nchunks = (P + 255) / 256;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    // z == 0, but opaque to the compiler: keeps it from hoisting the (loop-invariant) shared-memory weights
    // out of the chunk loop into registers (255 registers + spills otherwise, measured)
    const int z = (int)((unsigned long long)chunk >> 44);
    const float* q_wseg = s_wseg + z;
    const float* q_w21 = s_w21 + z;
    long long pix[2], n[2], hw[2];
    bool ok[2];
    float f[2][CF];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pix[u] = chunk * 256 + u * 128 + threadIdx.x;
      ok[u] = pix[u] < P;
      n[u] = ok[u] ? pix[u] / HW : 0;
      hw[u] = ok[u] ? pix[u] - n[u] * HW : 0;
#pragma unroll
      for (int c = 0; c < CF; c += 4) {
        const float4 v = ok[u] ? ld4(feat + pix[u] * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[u][c] = v.x; f[u][c + 1] = v.y; f[u][c + 2] = v.z; f[u][c + 3] = v.w;
      }
    }
    float lg[2][D::NCP];
    heads_logits2<CF, NC>(q_wseg, f, lg);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      if (logits_nhwc) {
#pragma unroll
        for (int k = 0; k < NC; ++k) st1(logits_nhwc + pix[u] * ld + k, lg[u][k]);
      }
      if (logits_out) {
#pragma unroll
        for (int k = 0; k < NC; ++k) logits_out[(n[u] * NC + k) * HW + hw[u]] = lg[u][k];
      }
      if (do_softmax) {
        float mx = lg[u][0];
#pragma unroll
        for (int k = 1; k < NC; ++k) mx = fmaxf(mx, lg[u][k]);
        float pr[NC], ssum = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) { pr[k] = expf(lg[u][k] - mx); ssum += pr[k]; }
        const float inv = 1.f / ssum;
#pragma unroll
        for (int k = 0; k < NC; ++k) seg[(n[u] * NC + k) * HW + hw[u]] = pr[k] * inv;
      } else {
#pragma unroll
        for (int k = 0; k < NC; ++k) seg[(n[u] * NC + k) * HW + hw[u]] = lg[u][k];
      }
    }
    if (NL > 0) {
#pragma unroll 2
      for (int l = 0; l < NL; ++l) {
        const float* wr = q_w21 + l * D::NCATP;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
        for (int c = 0; c < CF; c += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wr + c);
          a0.x = fmaf(w.x, f[0][c], a0.x); a0.y = fmaf(w.y, f[0][c + 1], a0.y);
          a0.z = fmaf(w.z, f[0][c + 2], a0.z); a0.w = fmaf(w.w, f[0][c + 3], a0.w);
          a1.x = fmaf(w.x, f[1][c], a1.x); a1.y = fmaf(w.y, f[1][c + 1], a1.y);
          a1.z = fmaf(w.z, f[1][c + 2], a1.z); a1.w = fmaf(w.w, f[1][c + 3], a1.w);
        }
#pragma unroll
        for (int k = 0; k < D::NCP; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wr + CF + k);   // padded columns hold zeros
          a0.x = fmaf(w.x, lg[0][k], a0.x); a0.y = fmaf(w.y, lg[0][k + 1], a0.y);
          a0.z = fmaf(w.z, lg[0][k + 2], a0.z); a0.w = fmaf(w.w, lg[0][k + 3], a0.w);
          a1.x = fmaf(w.x, lg[1][k], a1.x); a1.y = fmaf(w.y, lg[1][k + 1], a1.y);
          a1.z = fmaf(w.z, lg[1][k + 2], a1.z); a1.w = fmaf(w.w, lg[1][k + 3], a1.w);
        }
        if (ok[0]) heat[(n[0] * NL + l) * HW + hw[0]] = (a0.x + a0.y) + (a0.z + a0.w);
        if (ok[1]) heat[(n[1] * NL + l) * HW + hw[1]] = (a1.x + a1.y) + (a1.z + a1.w);
      }
    }
  }
}

// Fused heads backward.  Per pixel (one thread = two pixels): recompute logits/softmax from the features,
//   dcat = W21^T dheat ; dlg = dcat[CF:] + softmax'(dseg) ; dfeat = dcat[:CF] + Wseg^T dlg.
// Weight gradients: with cat = [feat ; logits] and mid = W1 cat,
//   dW2 = sum dheat (x) mid = G1 W1^T ,  dW1 = sum dmid (x) cat = W2^T G1 ,  G1 = sum_pixels dheat (x) cat  (NL x (CF+NC))
//   dWseg = Gseg = sum_pixels dlg (x) feat  (NC x CF)
// so only G1 and Gseg (770 numbers for the paper heads) are reduced over pixels: per 256-pixel chunk the
// per-pixel vectors go through shared memory and register-tiled outer products accumulate them; blocks are
// persistent and add their partial G's to global memory once.  heads_bwd_finalize_kernel forms dW1, dW2.
// Dynamic shared memory: heads_bwd_smem_bytes<...>().
template <int CF, int NC, int NF, int NL>
constexpr size_t heads_bwd_smem_bytes() {
  typedef HeadsDims<CF, NC, NF, NL> D;
  return sizeof(float) * ((size_t)(D::NLp + D::NCAT + NC) * 256 + NC * CF + (NL > 0 ? NL * D::NCATP : 4));
}
// bf16 storage: the two outer-product reductions run on the tensor cores (warp-level mma.sync over bf16 copies of the
// per-pixel vectors), see heads_bwd_fused_kernel.  Shared memory: fp32 dheat rows + bf16 [rows][256 + 8] tiles.
template <int CF, int NC, int NL>
struct HeadsMma {
  static constexpr int MR = NL + NC;                         // gradient rows: dheat (NL), dlg (NC)
  static constexpr int MT = (MR + 15) / 16;                  // 16-row tiles
  static constexpr int NB = CF + (NL > 0 ? NC : 0);          // columns: feat (CF) [, logits (NC)]
  static constexpr int NT = (NB + 7) / 8;                    // 8-column tiles
  static constexpr int PITCH = 256 + 8;                      // bf16 elements per row: 528 B, conflict-free ldmatrix rows
  static constexpr int NLp = NL > 0 ? NL : 1;
};
template <int CF, int NC, int NF, int NL>
constexpr size_t heads_bwd_smem_bytes_mma() {
  typedef HeadsDims<CF, NC, NF, NL> D;
  typedef HeadsMma<CF, NC, NL> M;
  return sizeof(float) * ((size_t)M::NLp * 256 + NC * CF + (NL > 0 ? NL * D::NCATP : 4)) +
         2 * (size_t)(M::MT * 16 + M::NT * 8) * M::PITCH;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// --------------------------------------------------------------------------
// Fused heads forward on the tensor cores (bf16 storage; the paper heads: 32 features, 7 classes, 39 -> 21 -> 14).
// The CUDA-core kernel above spends 784 FMAs per pixel and ran at a quarter of its issue rate (109 us for 1.18 M pixels
// against ~30 us of HBM time).  Here a warp owns 16-pixel tiles and the two small matrix products are warp-level
// mma.sync.m16n8k16 over the features exactly as they are stored:
//   * A (features): the contraction index may be permuted freely as long as A and B agree, so thread (g, t) of the warp
//     takes channels 8t .. 8t+7 of pixels g and g + 8 -- ONE 16-byte load per pixel row feeds both K steps, and a warp
//     instruction reads 8 x 64 contiguous bytes (every sector fully used);
//   * B (weights): kept in registers for the whole persistent block as split bf16 pairs W = hi + lo (two MMAs per
//     product), so the fp32 weights lose nothing to the bf16 operand format (~2^-17 relative);
//   * the logits accumulator fragment of a tile IS the A fragment of the third K step of the landmark product
//     (heat = W21 [feat ; logits]); the logits go in as hi + lo as well (three MMAs: hi*hi, hi*lo, lo*hi);
//   * softmax over the 7 classes of a pixel = two xor-shuffles among the 4 threads that hold its row;
//   * the fp32 NCHW outputs are written straight from the accumulator fragments: a warp store covers 4 planes x
//     8 consecutive pixels = four fully written 32-byte sectors.
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t heads_pack_hi(float a, float b, float& ra, float& rb) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
  ra = a - __bfloat162float(ha); rb = b - __bfloat162float(hb);
  return (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
}
__device__ __forceinline__ uint32_t heads_pack(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// The training loss inside the heads kernels (Dice / NCC partial sums from the head kernel's epilogue, the loss
// gradient formed in the backward head kernel's prologue).  Targets are (B, C, Ht, Wt) fp32 planes addressed through the centre-crop
// window (origin r0, c0 of the full H x W output; util.py:92-114 as called at train.py:414-417).
struct HeadsLoss {
  const float* mask; long long mask_sb, mask_sc; int mask_sr;            // segmentation targets: batch, channel, row strides
  const float* heat_t; long long heat_t_sb, heat_t_sc; int heat_t_sr;    // heat-map targets
  int Ht, Wt, r0, c0;
  FastDiv fd_w;               // division by the full output width W
  int W;
  double* sums;               // forward: [B][NC*3 + NL*5] (kernels_loss.cuh), zeroed by the caller
  const float* coef;          // backward: [B][NC*2 + NL*3] per-plane coefficients (loss_coef_kernel)
  const float* heat;          // backward: the heat-map predictions the forward stored, (B, NL, H, W)
};

template <int CF, int NC, int NF, int NL, bool LOSS = false>
__global__ void __launch_bounds__(128, LOSS ? 3 : 6) heads_fwd_mma_kernel(const bf16* __restrict__ feat, int ld, const float* wseg, const float* w1,
                                                            const float* w2, bf16* logits_nhwc, float* seg, float* logits_out,
                                                            float* heat, int P, int HW, FastDiv fd_hw, int do_softmax,
                                                            const HeadsLoss lossp) {
  pdl_wait(); pdl_trigger();
  typedef HeadsDims<CF, NC, NF, NL> D;
  static_assert(CF == 32 && NC <= 8 && NL <= 16, "fragment layout below: 32 features, <= 8 classes, <= 16 landmarks");
  constexpr int NT = NL > 8 ? 2 : (NL > 0 ? 1 : 0);          // 8-column tiles of the landmark product
  __shared__ __align__(16) float s_wseg[NC * CF];
  __shared__ __align__(16) float s_w21[NL > 0 ? NL * D::NCATP : 4];
  for (int i = threadIdx.x; i < NC * CF; i += blockDim.x) s_wseg[i] = wseg[i];
  heads_fold_w21<CF, NC, NF, NL>(w1, w2, s_w21);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B fragments (hi | lo).  K step s of the feature part: logical k = 2t, 2t+1 | 2t+8, 2t+9  <->  channels 8t+4s .. 8t+4s+3
  uint32_t bs_hi[2][2], bs_lo[2][2];                          // logits: [K step][b0 | b1], column n = g (class)
  uint32_t bh_hi[NT > 0 ? NT : 1][3][2], bh_lo[NT > 0 ? NT : 1][3][2];   // landmarks: [N tile][K step (2: the logits)][b0 | b1]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 8 * t + 4 * ks + 2 * h;
      const float a = g < NC ? s_wseg[g * CF + c] : 0.f, b = g < NC ? s_wseg[g * CF + c + 1] : 0.f;
      float ra, rb;
      bs_hi[ks][h] = heads_pack_hi(a, b, ra, rb);
      bs_lo[ks][h] = heads_pack(ra, rb);
    }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int l = 8 * j + g;
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float a = 0.f, b = 0.f;
        if (l < NL) {
          if (ks < 2) { const int c = 8 * t + 4 * ks + 2 * h; a = s_w21[l * D::NCATP + c]; b = s_w21[l * D::NCATP + c + 1]; }
          else if (h == 0) {                                  // logical k = class 2t, 2t+1 (padded columns of W21 hold zeros)
            a = (2 * t) < NC ? s_w21[l * D::NCATP + CF + 2 * t] : 0.f;
            b = (2 * t + 1) < NC ? s_w21[l * D::NCATP + CF + 2 * t + 1] : 0.f;
          }
        }
        float ra, rb;
        bh_hi[j][ks][h] = heads_pack_hi(a, b, ra, rb);
        bh_lo[j][ks][h] = heads_pack(ra, rb);
      }
  }
  const int warp_g = blockIdx.x * 4 + (threadIdx.x >> 5), nwarps = gridDim.x * 4;
  const int ntiles = (P + 15) >> 4;
  // (the feature rows of the warp's next tile are requested before the current tile is worked on)
  auto load_rows = [&](int tile, uint4& q0, uint4& q1) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0;
    if (tile < ntiles && r0 < P) q0 = *reinterpret_cast<const uint4*>(feat + (long long)r0 * ld + 8 * t);
    if (tile < ntiles && r1 < P) q1 = *reinterpret_cast<const uint4*>(feat + (long long)r1 * ld + 8 * t);
  };
  // LOSS: a warp walks a CONTIGUOUS range of tiles, so its loss sums stay in registers until the image index changes (the
  // caller guarantees H * W % 16 == 0: a tile never straddles two images); otherwise tiles are dealt round-robin.
  const int tpw = LOSS ? (ntiles + nwarps - 1) / nwarps : 1;
  const int t_first = LOSS ? warp_g * tpw : warp_g, t_step = LOSS ? 1 : nwarps;
  const int t_end = LOSS ? min(ntiles, t_first + tpw) : ntiles;
  // per-thread loss sums of the current image: Dice (t p, t t, p p) for classes 2t, 2t+1; NCC (x, xx, y, yy, xy) for
  // landmarks 8j + 2t, 8j + 2t + 1.  fp32 per tile (two pixels), fp64 across tiles (the NCC variances are differences of sums)
  double ld_acc[LOSS ? 2 : 1][3], ln_acc[LOSS && NT > 0 ? 2 * NT : 1][5];
  int cur_n = -1;
  if (LOSS) {
#pragma unroll
    for (int i = 0; i < 2; ++i) ld_acc[LOSS ? i : 0][0] = ld_acc[LOSS ? i : 0][1] = ld_acc[LOSS ? i : 0][2] = 0.0;
#pragma unroll
    for (int i = 0; i < (NT > 0 ? 2 * NT : 1); ++i)
#pragma unroll
      for (int k = 0; k < 5; ++k) ln_acc[LOSS && NT > 0 ? i : 0][k] = 0.0;
  }
  auto flush_loss = [&]() {
    if (!LOSS || cur_n < 0) return;
    constexpr int per = NC * 3 + NL * 5;
    double* o = lossp.sums + (long long)cur_n * per;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = ld_acc[LOSS ? i : 0][k];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0 && 2 * t + i < NC) atomicAdd(o + (2 * t + i) * 3 + k, v);
        ld_acc[LOSS ? i : 0][k] = 0.0;
      }
#pragma unroll
    for (int i = 0; i < (NT > 0 ? 2 * NT : 0); ++i)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        double v = ln_acc[LOSS && NT > 0 ? i : 0][k];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int l = 8 * (i >> 1) + 2 * t + (i & 1);
        if (g == 0 && l < NL) atomicAdd(o + NC * 3 + l * 5 + k, v);
        ln_acc[LOSS && NT > 0 ? i : 0][k] = 0.0;
      }
  };
  // (measured and not kept: two tiles ahead -- 63.8 -> 65.8 us)
  uint4 nq0, nq1;
  load_rows(t_first, nq0, nq1);
  for (int tile = t_first; tile < t_end; tile += t_step) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool ok0 = r0 < P, ok1 = r1 < P;
    const uint4 q0 = nq0, q1 = nq1;
    load_rows(tile + t_step < t_end ? tile + t_step : ntiles, nq0, nq1);
    const uint32_t a_k0[4] = {q0.x, q1.x, q0.y, q1.y}, a_k1[4] = {q0.z, q1.z, q0.w, q1.w};
    float lg[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(lg, a_k0, bs_hi[0]); mma_bf16_16816(lg, a_k1, bs_hi[1]);
    mma_bf16_16816(lg, a_k0, bs_lo[0]); mma_bf16_16816(lg, a_k1, bs_lo[1]);
    // lg[0], lg[1]: pixel r0, classes 2t, 2t+1;  lg[2], lg[3]: pixel r1
    const int n0 = fd_hw.div(ok0 ? r0 : 0), n1 = fd_hw.div(ok1 ? r1 : 0);
    const int hw0 = (ok0 ? r0 : 0) - n0 * HW, hw1 = (ok1 ? r1 : 0) - n1 * HW;
    const bool c0ok = 2 * t < NC, c1ok = 2 * t + 1 < NC;
    // LOSS: window coordinates of the two pixels (inside the crop window or not) and the image the tile belongs to
    bool in0 = false, in1 = false;
    long long toff0 = 0, toff1 = 0, hoff0 = 0, hoff1 = 0;      // offsets of the pixels inside a target plane (mask | heat_t)
    if (LOSS) {
      if (n0 != cur_n) { flush_loss(); cur_n = n0; }
      int row, col;
      lossp.fd_w.divmod(hw0, row, col);
      row -= lossp.r0; col -= lossp.c0;
      in0 = ok0 && row >= 0 && row < lossp.Ht && col >= 0 && col < lossp.Wt;
      toff0 = (long long)n0 * lossp.mask_sb + (long long)row * lossp.mask_sr + col;
      hoff0 = (long long)n0 * lossp.heat_t_sb + (long long)row * lossp.heat_t_sr + col;
      lossp.fd_w.divmod(hw1, row, col);
      row -= lossp.r0; col -= lossp.c0;
      in1 = ok1 && row >= 0 && row < lossp.Ht && col >= 0 && col < lossp.Wt;
      toff1 = (long long)n1 * lossp.mask_sb + (long long)row * lossp.mask_sr + col;
      hoff1 = (long long)n1 * lossp.heat_t_sb + (long long)row * lossp.heat_t_sr + col;
    }
    if (logits_nhwc) {
      if (c1ok) {
        if (ok0) *reinterpret_cast<uint32_t*>(logits_nhwc + (long long)r0 * ld + 2 * t) = heads_pack(lg[0], lg[1]);
        if (ok1) *reinterpret_cast<uint32_t*>(logits_nhwc + (long long)r1 * ld + 2 * t) = heads_pack(lg[2], lg[3]);
      } else if (c0ok) {
        if (ok0) logits_nhwc[(long long)r0 * ld + 2 * t] = __float2bfloat16_rn(lg[0]);
        if (ok1) logits_nhwc[(long long)r1 * ld + 2 * t] = __float2bfloat16_rn(lg[2]);
      }
    }
    float* seg0 = seg + ((long long)n0 * NC + 2 * t) * HW + hw0;
    float* seg1 = seg + ((long long)n1 * NC + 2 * t) * HW + hw1;
    float pv[4];                  // the segmentation output of the two pixels (classes 2t, 2t+1)
    if (logits_out) {
      float* lo0 = logits_out + ((long long)n0 * NC + 2 * t) * HW + hw0;
      float* lo1 = logits_out + ((long long)n1 * NC + 2 * t) * HW + hw1;
      if (ok0 && c0ok) lo0[0] = lg[0];
      if (ok0 && c1ok) lo0[HW] = lg[1];
      if (ok1 && c0ok) lo1[0] = lg[2];
      if (ok1 && c1ok) lo1[HW] = lg[3];
    }
    if (do_softmax) {
      float m0 = fmaxf(c0ok ? lg[0] : -INFINITY, c1ok ? lg[1] : -INFINITY);
      float m1 = fmaxf(c0ok ? lg[2] : -INFINITY, c1ok ? lg[3] : -INFINITY);
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float e00 = c0ok ? expf(lg[0] - m0) : 0.f, e01 = c1ok ? expf(lg[1] - m0) : 0.f;
      const float e10 = c0ok ? expf(lg[2] - m1) : 0.f, e11 = c1ok ? expf(lg[3] - m1) : 0.f;
      float s0 = e00 + e01, s1 = e10 + e11;
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float i0 = 1.f / s0, i1 = 1.f / s1;
      pv[0] = e00 * i0; pv[1] = e01 * i0; pv[2] = e10 * i1; pv[3] = e11 * i1;
    } else {
      pv[0] = lg[0]; pv[1] = lg[1]; pv[2] = lg[2]; pv[3] = lg[3];
    }
    if (seg) {
      if (ok0 && c0ok) seg0[0] = pv[0];
      if (ok0 && c1ok) seg0[HW] = pv[1];
      if (ok1 && c0ok) seg1[0] = pv[2];
      if (ok1 && c1ok) seg1[HW] = pv[3];
    }
    if (LOSS) {
      // Dice sums (dice.py:20-55) of the two pixels, classes 2t, 2t+1
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cok = i ? c1ok : c0ok;
        const float t0 = (in0 && cok) ? lossp.mask[toff0 + (long long)(2 * t + i) * lossp.mask_sc] : 0.f;
        const float t1 = (in1 && cok) ? lossp.mask[toff1 + (long long)(2 * t + i) * lossp.mask_sc] : 0.f;
        const float p0 = (in0 && cok) ? pv[i] : 0.f, p1 = (in1 && cok) ? pv[2 + i] : 0.f;
        ld_acc[LOSS ? i : 0][0] += (double)fmaf(t0, p0, t1 * p1);
        ld_acc[LOSS ? i : 0][1] += (double)fmaf(t0, t0, t1 * t1);
        ld_acc[LOSS ? i : 0][2] += (double)fmaf(p0, p0, p1 * p1);
      }
    }
    if (NT > 0) {
      // third K step: A = [logits | 0], hi + lo
      float r00, r01, r10, r11;
      const uint32_t al_hi[4] = {heads_pack_hi(c0ok ? lg[0] : 0.f, c1ok ? lg[1] : 0.f, r00, r01),
                                 heads_pack_hi(c0ok ? lg[2] : 0.f, c1ok ? lg[3] : 0.f, r10, r11), 0u, 0u};
      const uint32_t al_lo[4] = {heads_pack(r00, r01), heads_pack(r10, r11), 0u, 0u};
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        float h4[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(h4, a_k0, bh_hi[j][0]); mma_bf16_16816(h4, a_k1, bh_hi[j][1]);
        mma_bf16_16816(h4, a_k0, bh_lo[j][0]); mma_bf16_16816(h4, a_k1, bh_lo[j][1]);
        mma_bf16_16816(h4, al_hi, bh_hi[j][2]); mma_bf16_16816(h4, al_hi, bh_lo[j][2]); mma_bf16_16816(h4, al_lo, bh_hi[j][2]);
        const int l = 8 * j + 2 * t;
        float* h0 = heat + ((long long)n0 * NL + l) * HW + hw0;
        float* h1 = heat + ((long long)n1 * NL + l) * HW + hw1;
        if (ok0 && l < NL) h0[0] = h4[0];
        if (ok0 && l + 1 < NL) h0[HW] = h4[1];
        if (ok1 && l < NL) h1[0] = h4[2];
        if (ok1 && l + 1 < NL) h1[HW] = h4[3];
        if (LOSS) {
          // NCC moments (ncc.py:12-38) of the two pixels, landmarks l, l + 1
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const bool lok = l + i < NL;
            const float x0 = (in0 && lok) ? h4[i] : 0.f, x1 = (in1 && lok) ? h4[2 + i] : 0.f;
            const float y0 = (in0 && lok) ? lossp.heat_t[hoff0 + (long long)(l + i) * lossp.heat_t_sc] : 0.f;
            const float y1 = (in1 && lok) ? lossp.heat_t[hoff1 + (long long)(l + i) * lossp.heat_t_sc] : 0.f;
            double* a = ln_acc[LOSS && NT > 0 ? 2 * j + i : 0];
            a[0] += (double)(x0 + x1); a[1] += (double)fmaf(x0, x0, x1 * x1);
            a[2] += (double)(y0 + y1); a[3] += (double)fmaf(y0, y0, y1 * y1);
            a[4] += (double)fmaf(x0, y0, x1 * y1);
          }
        }
      }
    }
  }
  flush_loss();
}

template <typename T, int CF, int NC, int NF, int NL>
__global__ void __launch_bounds__(128, NL > 0 ? 3 : 2) heads_bwd_fused_kernel(const T* feat, int ld, const float* wseg, const float* w1,
                                                              const float* w2, const float* d_seg, const float* d_heat,
                                                              T* d_feat, int d_ld, float* g_acc /*[NL*(CF+NC) + NC*CF]*/,
                                                              int B, long long HW, int do_softmax) {
  pdl_wait(); pdl_trigger();
  typedef HeadsDims<CF, NC, NF, NL> D;
  constexpr int NCAT = D::NCAT, NCATP = D::NCATP, NLp = D::NLp;
  constexpr int ROWS = NLp + NCAT + NC;          // dheat rows, cat rows, dlg rows
  constexpr int TA = 2, TB = 3;                  // G1 register tile
  constexpr int NTA = (NLp + TA - 1) / TA, NTB = (NCAT + TB - 1) / TB;
  static_assert(NTA * NTB + CF <= 128, "outer-product tiles must fit one block");
  extern __shared__ __align__(16) float heads_smem[];
  // bf16 storage: outer products on the tensor cores.  The per-pixel vectors are ALSO written as bf16 rows
  // A_s = [dheat ; dlg] (MT*16 rows) and B_s = [feat ; logits] (NT*8 rows), 256 pixels (K) per row; warp w owns the
  // 8-column tiles w, w + 4 of  G = A_s B_s^T  and runs 16 K steps of mma.sync.m16n8k16 per chunk with fp32 accumulators
  // kept in registers for the whole (persistent) block.  The fp32 version below read 5 float4 per 24 FMAs from shared
  // memory and was bound by its bandwidth (~80 of the kernel's 240 us at B = 32 @192x192).
  constexpr bool kMma = sizeof(T) == 2;
  typedef HeadsMma<CF, NC, NL> MM;
  float* s_v = heads_smem;                       // [ROWS][256] per-pixel vectors, one column per pixel (kMma: dheat rows only)
  float* s_wseg = s_v + (kMma ? NLp : ROWS) * 256;   // [NC][CF]
  float* s_w21 = s_wseg + NC * CF;               // [NL][NCATP]
  bf16* A_s = reinterpret_cast<bf16*>(s_w21 + (NL > 0 ? NL * NCATP : 4));   // [MT*16][PITCH]
  bf16* B_s = A_s + MM::MT * 16 * MM::PITCH;                                 // [NT*8][PITCH]
  for (int i = threadIdx.x; i < NC * CF; i += blockDim.x) s_wseg[i] = wseg[i];
  heads_fold_w21<CF, NC, NF, NL>(w1, w2, s_w21);
  if (kMma) {       // padding rows are never written: zero everything once
    uint32_t* z = reinterpret_cast<uint32_t*>(A_s);
    for (int i = threadIdx.x; i < (MM::MT * 16 + MM::NT * 8) * MM::PITCH / 2; i += blockDim.x) z[i] = 0u;
  }
  float macc[MM::MT][2][4];
#pragma unroll
  for (int i = 0; i < MM::MT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { macc[i][j][0] = macc[i][j][1] = macc[i][j][2] = macc[i][j][3] = 0.f; }
  const int tid = threadIdx.x;
  // outer-product ownership
  const bool own_g1 = NL > 0 && tid < NTA * NTB;
  const int ta = own_g1 ? (tid / NTB) * TA : 0, tb = own_g1 ? (tid % NTB) * TB : 0;
  const bool own_gs = tid >= 128 - CF;            // last CF threads: one feature column each, all NC rows
  const int gs_c = tid - (128 - CF);
  float g1[TA][TB], gs[NC];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) g1[i][j] = 0.f;
#pragma unroll
  for (int k = 0; k < NC; ++k) gs[k] = 0.f;
  const long long P = (long long)B * HW;
  const long long nchunks = (P + 255) / 256;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    __syncthreads();                              // weights ready / previous chunk's outer products done with s_v
    const int z = (int)((unsigned long long)chunk >> 44);   // == 0, opaque (see heads_fwd_fused_kernel)
    const float* q_wseg = s_wseg + z;
    const float* q_w21 = s_w21 + z;
    long long pix[2], n[2], hw[2];
    bool ok[2];
    float f[2][CF];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pix[u] = chunk * 256 + u * 128 + tid;
      ok[u] = pix[u] < P;
      n[u] = ok[u] ? pix[u] / HW : 0;
      hw[u] = ok[u] ? pix[u] - n[u] * HW : 0;
#pragma unroll
      for (int c = 0; c < CF; c += 4) {
        const float4 v = ok[u] ? ld4(feat + pix[u] * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[u][c] = v.x; f[u][c + 1] = v.y; f[u][c + 2] = v.z; f[u][c + 3] = v.w;
      }
    }
    // upstream gradients (fp32 NCHW: consecutive threads read consecutive addresses)
    // (dheat goes straight to its shared-memory rows and is read back per landmark below: 28 registers less)
    float dlg[2][D::NCP];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float* col = s_v + u * 128 + tid;
#pragma unroll
      for (int l = 0; l < NLp; ++l) {
        const float dv = (NL > 0 && d_heat && ok[u]) ? d_heat[(n[u] * NL + l) * HW + hw[u]] : 0.f;
        col[l * 256] = dv;
        if (kMma && NL > 0) A_s[l * MM::PITCH + u * 128 + tid] = __float2bfloat16_rn(dv);
      }
#pragma unroll
      for (int k = 0; k < D::NCP; ++k) dlg[u][k] = (k < NC && d_seg && ok[u]) ? d_seg[(n[u] * NC + k) * HW + hw[u]] : 0.f;
    }
    float lg[2][D::NCP];
    heads_logits2<CF, NC>(q_wseg, f, lg);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float* col = s_v + u * 128 + tid;           // row r of this pixel at col[r*256]
      if (kMma) {
#pragma unroll
        for (int c = 0; c < CF; ++c) B_s[c * MM::PITCH + u * 128 + tid] = __float2bfloat16_rn(f[u][c]);
        if (NL > 0) {
#pragma unroll
          for (int k = 0; k < NC; ++k) B_s[(CF + k) * MM::PITCH + u * 128 + tid] = __float2bfloat16_rn(lg[u][k]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CF; ++c) col[(NLp + c) * 256] = f[u][c];
#pragma unroll
        for (int k = 0; k < NC; ++k) col[(NLp + CF + k) * 256] = lg[u][k];
      }
      if (do_softmax && d_seg) {
        float mx = lg[u][0];
#pragma unroll
        for (int k = 1; k < NC; ++k) mx = fmaxf(mx, lg[u][k]);
        float pr[NC], ssum = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) { pr[k] = expf(lg[u][k] - mx); ssum += pr[k]; }
        const float inv = 1.f / ssum;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) { pr[k] *= inv; dot = fmaf(pr[k], dlg[u][k], dot); }
#pragma unroll
        for (int k = 0; k < NC; ++k) dlg[u][k] = pr[k] * (dlg[u][k] - dot);
      }
    }
    // dcat = W21^T dheat: f is dead from here on, its registers become dfeat
    float (&df)[2][CF] = f;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int c = 0; c < CF; ++c) df[u][c] = 0.f;
    if (NL > 0) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const float* wr = q_w21 + l * NCATP;
        const float dh[2] = {s_v[l * 256 + tid], s_v[l * 256 + 128 + tid]};
#pragma unroll
        for (int c = 0; c < CF; c += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wr + c);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            df[u][c] = fmaf(w.x, dh[u], df[u][c]); df[u][c + 1] = fmaf(w.y, dh[u], df[u][c + 1]);
            df[u][c + 2] = fmaf(w.z, dh[u], df[u][c + 2]); df[u][c + 3] = fmaf(w.w, dh[u], df[u][c + 3]);
          }
        }
#pragma unroll
        for (int k = 0; k < D::NCP; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wr + CF + k);   // padded columns hold zeros
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            dlg[u][k] = fmaf(w.x, dh[u], dlg[u][k]); dlg[u][k + 1] = fmaf(w.y, dh[u], dlg[u][k + 1]);
            dlg[u][k + 2] = fmaf(w.z, dh[u], dlg[u][k + 2]); dlg[u][k + 3] = fmaf(w.w, dh[u], dlg[u][k + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (kMma) {
        A_s[(NL + k) * MM::PITCH + tid] = __float2bfloat16_rn(dlg[0][k]);
        A_s[(NL + k) * MM::PITCH + 128 + tid] = __float2bfloat16_rn(dlg[1][k]);
      } else {
        s_v[(NLp + NCAT + k) * 256 + tid] = dlg[0][k];
        s_v[(NLp + NCAT + k) * 256 + 128 + tid] = dlg[1][k];
      }
#pragma unroll
      for (int c = 0; c < CF; c += 4) {
        const float4 w = *reinterpret_cast<const float4*>(q_wseg + k * CF + c);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          df[u][c] = fmaf(w.x, dlg[u][k], df[u][c]); df[u][c + 1] = fmaf(w.y, dlg[u][k], df[u][c + 1]);
          df[u][c + 2] = fmaf(w.z, dlg[u][k], df[u][c + 2]); df[u][c + 3] = fmaf(w.w, dlg[u][k], df[u][c + 3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int c = 0; c < CF; c += 4)
        st4(d_feat + pix[u] * d_ld + c, make_float4(df[u][c], df[u][c + 1], df[u][c + 2], df[u][c + 3]));
    }
    __syncthreads();
    if (kMma) {
      // G += A_s B_s^T over the chunk's 256 pixels: this warp's 8-column tiles wq, wq + 4
      const int lane = tid & 31, wq = tid >> 5;
      const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(A_s) +
                              2u * (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * MM::PITCH + 8 * (lane >> 4));
      const uint32_t b_base = (uint32_t)__cvta_generic_to_shared(B_s) +
                              2u * (uint32_t)((lane & 7) * MM::PITCH + 8 * ((lane >> 3) & 1));
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        uint32_t af[MM::MT][4];
#pragma unroll
        for (int mt = 0; mt < MM::MT; ++mt) ldsm_x4(a_base + 2u * (uint32_t)(mt * 16 * MM::PITCH + ks * 16), af[mt]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nt = wq + 4 * j;
          if (nt < MM::NT) {
            uint32_t bfr[2];
            ldsm_x2(b_base + 2u * (uint32_t)(nt * 8 * MM::PITCH + ks * 16), bfr);
#pragma unroll
            for (int mt = 0; mt < MM::MT; ++mt) mma_bf16_16816(macc[mt][j], af[mt], bfr);
          }
        }
      }
      continue;          // (the next chunk's leading __syncthreads orders these reads before its writes)
    }
    // outer products over the chunk's 256 pixels (float4 along the pixel axis); out-of-range pixels hold zeros
    if (own_g1) {
#pragma unroll 2
      for (int p4 = 0; p4 < 64; ++p4) {
        const int px = ((p4 + tid) & 63) * 4;     // staggered start: no bank conflicts between tiles
        float4 a[TA], b[TB];
#pragma unroll
        for (int i = 0; i < TA; ++i) a[i] = (ta + i < NL) ? *reinterpret_cast<const float4*>(s_v + (ta + i) * 256 + px) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < TB; ++j) b[j] = (tb + j < NCAT) ? *reinterpret_cast<const float4*>(s_v + (NLp + tb + j) * 256 + px) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < TA; ++i)
#pragma unroll
          for (int j = 0; j < TB; ++j)
            g1[i][j] += a[i].x * b[j].x + a[i].y * b[j].y + a[i].z * b[j].z + a[i].w * b[j].w;
      }
    }
    if (own_gs) {
#pragma unroll 1
      for (int p4 = 0; p4 < 64; ++p4) {
        const int px = ((p4 + tid) & 63) * 4;
        const float4 fv = *reinterpret_cast<const float4*>(s_v + (NLp + gs_c) * 256 + px);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const float4 dv = *reinterpret_cast<const float4*>(s_v + (NLp + NCAT + k) * 256 + px);
          gs[k] += dv.x * fv.x + dv.y * fv.y + dv.z * fv.z + dv.w * fv.w;
        }
      }
    }
  }
  if (kMma) {
    // fragment (mt, j): rows mt*16 + lane/4 (+8), columns (wq + 4j)*8 + 2*(lane%4) (+1)
    const int lane = tid & 31, wq = tid >> 5;
#pragma unroll
    for (int mt = 0; mt < MM::MT; ++mt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int nt = wq + 4 * j;
        if (nt >= MM::NT) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = mt * 16 + (lane >> 2) + 8 * (e >> 1), colx = nt * 8 + 2 * (lane & 3) + (e & 1);
          if (row < NL) { if (colx < NCAT) atomicAdd(g_acc + row * NCAT + colx, macc[mt][j][e]); }
          else if (row < NL + NC && colx < CF) atomicAdd(g_acc + NLp * NCAT * (NL > 0 ? 1 : 0) + (row - NL) * CF + colx, macc[mt][j][e]);
        }
      }
    return;
  }
  if (own_g1) {
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
      for (int j = 0; j < TB; ++j)
        if (ta + i < NL && tb + j < NCAT) atomicAdd(g_acc + (ta + i) * NCAT + tb + j, g1[i][j]);
  }
  if (own_gs) {
#pragma unroll
    for (int k = 0; k < NC; ++k) atomicAdd(g_acc + NLp * NCAT * (NL > 0 ? 1 : 0) + k * CF + gs_c, gs[k]);
  }
}

// --------------------------------------------------------------------------
// Fused heads backward on the tensor cores (bf16 storage), the counterpart of heads_fwd_mma_kernel: a warp owns 16-pixel
// tiles and every product is a warp-level mma.sync.m16n8k16 on fragments that never leave registers:
//   logits  = feat Wseg^T                       (recomputed; split-bf16 weights)
//   dcat    = dheat W21                         (A = dheat hi + lo, loaded from the fp32 NCHW planes in fragment order)
//   dlg     = dcat[CF:] + softmax'(dseg)        (accumulator-fragment layout: a pixel's classes sit in 4 neighbouring lanes)
//   dfeat   = dcat[:CF] + dlg Wseg              (the dlg fragment is the A operand; output columns are permuted so a thread
//                                                ends up with 8 contiguous channels of a pixel = one 16-byte store)
//   G1  += dheat^T [feat ; logits],  Gseg += dlg^T feat   (K = the tile's 16 pixels: `movmatrix.trans` turns the 8x8
//                                                blocks of the fragments above into the transposed operands; bf16 single
//                                                pass like the ldmatrix version in heads_bwd_fused_kernel)
// The CUDA-core kernel spent ~1000 FMAs per pixel plus a shared-memory round trip of every per-pixel vector (164 us for
// 1.18 M pixels against ~40 us of HBM time).
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// accumulate with a B fragment whose second half (k = 8 .. 15) is zero
__device__ __forceinline__ void mma_bf16_16816_b0(float (&c)[4], const uint32_t (&a)[4], uint32_t b0) {
  const uint32_t b[2] = {b0, 0u};
  mma_bf16_16816(c, a, b);
}

// LOSS: d_seg / d_heat are not read; the gradient of the training loss w.r.t. the two outputs is formed per pixel from the
// targets, the recomputed class probabilities, the stored heat-map predictions and the per-plane coefficients of
// loss_coef_kernel (closed forms of kernels_loss.cuh:loss_backward_kernel): inside the crop window
//   d_seg = A t + Bc p,   d_heat = ka y - kb x + kc;   zero outside.
template <int CF, int NC, int NF, int NL, bool LOSS = false>
__global__ void __launch_bounds__(128, 4) heads_bwd_mma_kernel(const bf16* __restrict__ feat, int ld, const float* wseg, const float* w1,
                                                               const float* w2, const float* __restrict__ d_seg,
                                                               const float* __restrict__ d_heat, bf16* d_feat, int d_ld,
                                                               float* g_acc /*[NL*(CF+NC) + NC*CF]*/, int P, int HW, FastDiv fd_hw,
                                                               int do_softmax, const HeadsLoss lossp) {
  pdl_wait(); pdl_trigger();
  typedef HeadsDims<CF, NC, NF, NL> D;
  static_assert(CF == 32 && NC <= 8 && NL <= 16, "fragment layout below: 32 features, <= 8 classes, <= 16 landmarks");
  constexpr int NCAT = D::NCAT, NCATP = D::NCATP;
  constexpr bool kL = NL > 0;
  constexpr int NG1 = kL ? NL * NCAT : 0, NG = NG1 + NC * CF;
  __shared__ __align__(16) float s_wseg[NC * CF];
  __shared__ __align__(16) float s_w21[kL ? NL * NCATP : 4];
  __shared__ float s_g[NG];
  for (int i = threadIdx.x; i < NC * CF; i += blockDim.x) s_wseg[i] = wseg[i];
  for (int i = threadIdx.x; i < NG; i += blockDim.x) s_g[i] = 0.f;
  heads_fold_w21<CF, NC, NF, NL>(w1, w2, s_w21);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // ---- weight fragments (hi | lo), resident in registers ----
  uint32_t bl_hi[2][2], bl_lo[2][2];                 // logits: K = features (permuted: step s, half h <-> channels 8t+4s+2h ..), n = class g
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 8 * t + 4 * ks + 2 * h;
      const float a = g < NC ? s_wseg[g * CF + c] : 0.f, b = g < NC ? s_wseg[g * CF + c + 1] : 0.f;
      float ra, rb;
      bl_hi[ks][h] = heads_pack_hi(a, b, ra, rb);
      bl_lo[ks][h] = heads_pack(ra, rb);
    }
  // output columns of the dfeat tiles: column n = g of tile j <-> channel 8 (g / 2) + 2 j + (g & 1), so that a thread's
  // accumulator columns 2t, 2t+1 of tiles 0..3 are channels 8t .. 8t+7
  uint32_t bd_hi[kL ? 5 : 1][2], bd_lo[kL ? 5 : 1][2];   // dcat: K = landmark, n = cat column (tile 4: class g)
  if (kL) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int col = j < 4 ? 8 * (g >> 1) + 2 * j + (g & 1) : CF + g;
      const bool cok = j < 4 || g < NC;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int l = 2 * t + 8 * h;
        const float a = (cok && l < NL) ? s_w21[l * NCATP + col] : 0.f, b = (cok && l + 1 < NL) ? s_w21[(l + 1) * NCATP + col] : 0.f;
        float ra, rb;
        bd_hi[j][h] = heads_pack_hi(a, b, ra, rb);
        bd_lo[j][h] = heads_pack(ra, rb);
      }
    }
  }
  uint32_t bf_hi[4], bf_lo[4];                       // dfeat += dlg Wseg: K = class (b1 = 0), n = channel as above
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 8 * (g >> 1) + 2 * j + (g & 1);
    const float a = (2 * t) < NC ? s_wseg[(2 * t) * CF + col] : 0.f, b = (2 * t + 1) < NC ? s_wseg[(2 * t + 1) * CF + col] : 0.f;
    float ra, rb;
    bf_hi[j] = heads_pack_hi(a, b, ra, rb);
    bf_lo[j] = heads_pack(ra, rb);
  }
  // The 36 fragment words go to shared memory ([word][lane]: the four warps hold identical values) and are re-read at their use:
  // with them in registers the kernel needed 160 registers = 3 blocks (12 warps) per SM, too few loads in flight for a pass
  // that moves 250 MB; now 4 blocks per SM.
  __shared__ uint32_t s_frag[36][32];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_frag[i][lane] = bl_hi[i >> 1][i & 1]; s_frag[4 + i][lane] = bl_lo[i >> 1][i & 1]; }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      s_frag[8 + i][lane] = kL ? bd_hi[kL ? i >> 1 : 0][i & 1] : 0u; s_frag[18 + i][lane] = kL ? bd_lo[kL ? i >> 1 : 0][i & 1] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_frag[28 + i][lane] = bf_hi[i]; s_frag[32 + i][lane] = bf_lo[i]; }
  }
  __syncthreads();
  auto frag2 = [&](int i, uint32_t (&b)[2]) { b[0] = s_frag[i][lane]; b[1] = s_frag[i + 1][lane]; };
  // ---- weight-gradient accumulators: G1 tiles (s, h) = features 8t+4s+2h.., tile 4 = logits; Gseg tiles (s, h) ----
  float g1[kL ? 5 : 1][4], gs[4][4];
#pragma unroll
  for (int j = 0; j < (kL ? 5 : 1); ++j) g1[j][0] = g1[j][1] = g1[j][2] = g1[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) gs[j][0] = gs[j][1] = gs[j][2] = gs[j][3] = 0.f;

  const int warp_g = blockIdx.x * 4 + (threadIdx.x >> 5), nwarps = gridDim.x * 4;
  const int ntiles = (P + 15) >> 4;
  auto load_rows = [&](int tile, uint4& q0, uint4& q1) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0;
    if (r0 < P) q0 = *reinterpret_cast<const uint4*>(feat + (long long)r0 * ld + 8 * t);
    if (r1 < P) q1 = *reinterpret_cast<const uint4*>(feat + (long long)r1 * ld + 8 * t);
  };
  uint4 nq0, nq1;
  load_rows(warp_g, nq0, nq1);
  const bool c0ok = 2 * t < NC, c1ok = 2 * t + 1 < NC;
  // upstream gradients of a tile in fragment order (a warp load covers 4 planes x 8 consecutive pixels); like the feature rows
  // they are requested one tile ahead (not in the LOSS variant, which derives them from targets and predictions)
  auto load_grads = [&](int tile, float (&ds)[4], float (&dh)[2][4]) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool ok0 = r0 < P, ok1 = r1 < P;
    const int n0 = fd_hw.div(ok0 ? r0 : 0), n1 = fd_hw.div(ok1 ? r1 : 0);
    const int hw0 = (ok0 ? r0 : 0) - n0 * HW, hw1 = (ok1 ? r1 : 0) - n1 * HW;
    ds[0] = ds[1] = ds[2] = ds[3] = 0.f;
    if (d_seg) {
      const float* p0 = d_seg + ((long long)n0 * NC + 2 * t) * HW + hw0;
      const float* p1 = d_seg + ((long long)n1 * NC + 2 * t) * HW + hw1;
      if (ok0 && c0ok) ds[0] = p0[0];
      if (ok0 && c1ok) ds[1] = p0[HW];
      if (ok1 && c0ok) ds[2] = p1[0];
      if (ok1 && c1ok) ds[3] = p1[HW];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int l = 2 * t + 8 * h + e;
        dh[0][2 * h + e] = (kL && d_heat && ok0 && l < NL) ? d_heat[((long long)n0 * NL + l) * HW + hw0] : 0.f;
        dh[1][2 * h + e] = (kL && d_heat && ok1 && l < NL) ? d_heat[((long long)n1 * NL + l) * HW + hw1] : 0.f;
      }
  };
  float nds[4] = {0.f, 0.f, 0.f, 0.f}, ndh[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  if (!LOSS && warp_g < ntiles) load_grads(warp_g, nds, ndh);
  for (int tile = warp_g; tile < ntiles; tile += nwarps) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool ok0 = r0 < P, ok1 = r1 < P;
    const uint4 q0 = nq0, q1 = nq1;
    load_rows(tile + nwarps, nq0, nq1);
    const int n0 = fd_hw.div(ok0 ? r0 : 0), n1 = fd_hw.div(ok1 ? r1 : 0);
    const int hw0 = (ok0 ? r0 : 0) - n0 * HW, hw1 = (ok1 ? r1 : 0) - n1 * HW;
    // LOSS: window coordinates of the two pixels
    bool in0 = false, in1 = false;
    long long toff0 = 0, toff1 = 0, hoff0 = 0, hoff1 = 0;
    const float* cf0 = nullptr; const float* cf1 = nullptr;
    if (LOSS) {
      int row, col;
      lossp.fd_w.divmod(hw0, row, col);
      row -= lossp.r0; col -= lossp.c0;
      in0 = ok0 && row >= 0 && row < lossp.Ht && col >= 0 && col < lossp.Wt;
      toff0 = (long long)n0 * lossp.mask_sb + (long long)row * lossp.mask_sr + col;
      hoff0 = (long long)n0 * lossp.heat_t_sb + (long long)row * lossp.heat_t_sr + col;
      lossp.fd_w.divmod(hw1, row, col);
      row -= lossp.r0; col -= lossp.c0;
      in1 = ok1 && row >= 0 && row < lossp.Ht && col >= 0 && col < lossp.Wt;
      toff1 = (long long)n1 * lossp.mask_sb + (long long)row * lossp.mask_sr + col;
      hoff1 = (long long)n1 * lossp.heat_t_sb + (long long)row * lossp.heat_t_sr + col;
      cf0 = lossp.coef + (long long)n0 * (NC * 2 + NL * 3);
      cf1 = lossp.coef + (long long)n1 * (NC * 2 + NL * 3);
    }
    // upstream gradients in fragment order (a warp load covers 4 planes x 8 consecutive pixels)
    float ds[4] = {nds[0], nds[1], nds[2], nds[3]};
    float dhp[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { dhp[0][i] = ndh[0][i]; dhp[1][i] = ndh[1][i]; }
    if (!LOSS && tile + nwarps < ntiles) load_grads(tile + nwarps, nds, ndh);
    uint32_t ah_hi[4] = {0u, 0u, 0u, 0u}, ah_lo[4] = {0u, 0u, 0u, 0u};      // dheat: rows = pixels, K = landmark
    if (kL && (LOSS || d_heat)) {
      float dh[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int l = 2 * t + 8 * h + e;
          if (LOSS) {
            float v0 = 0.f, v1 = 0.f;
            if (in0 && l < NL) {
              const float x = lossp.heat[((long long)n0 * NL + l) * HW + hw0], y = lossp.heat_t[hoff0 + (long long)l * lossp.heat_t_sc];
              const float* c3 = cf0 + NC * 2 + l * 3;
              v0 = fmaf(c3[0], y, fmaf(-c3[1], x, c3[2]));
            }
            if (in1 && l < NL) {
              const float x = lossp.heat[((long long)n1 * NL + l) * HW + hw1], y = lossp.heat_t[hoff1 + (long long)l * lossp.heat_t_sc];
              const float* c3 = cf1 + NC * 2 + l * 3;
              v1 = fmaf(c3[0], y, fmaf(-c3[1], x, c3[2]));
            }
            dh[0][2 * h + e] = v0; dh[1][2 * h + e] = v1;
            continue;
          }
          dh[0][2 * h + e] = dhp[0][2 * h + e];
          dh[1][2 * h + e] = dhp[1][2 * h + e];
        }
      float ra, rb;
      ah_hi[0] = heads_pack_hi(dh[0][0], dh[0][1], ra, rb); ah_lo[0] = heads_pack(ra, rb);
      ah_hi[1] = heads_pack_hi(dh[1][0], dh[1][1], ra, rb); ah_lo[1] = heads_pack(ra, rb);
      ah_hi[2] = heads_pack_hi(dh[0][2], dh[0][3], ra, rb); ah_lo[2] = heads_pack(ra, rb);
      ah_hi[3] = heads_pack_hi(dh[1][2], dh[1][3], ra, rb); ah_lo[3] = heads_pack(ra, rb);
    }
    const uint32_t a_k0[4] = {q0.x, q1.x, q0.y, q1.y}, a_k1[4] = {q0.z, q1.z, q0.w, q1.w};
    float lg[4] = {0.f, 0.f, 0.f, 0.f};
    {
      uint32_t b[2];
      frag2(0, b); mma_bf16_16816(lg, a_k0, b); frag2(2, b); mma_bf16_16816(lg, a_k1, b);
      frag2(4, b); mma_bf16_16816(lg, a_k0, b); frag2(6, b); mma_bf16_16816(lg, a_k1, b);
    }
    // dlg (softmax backward): lg[0], lg[1] / ds[0], ds[1]: pixel r0, classes 2t, 2t+1; [2], [3]: pixel r1
    float dlg[4] = {ds[0], ds[1], ds[2], ds[3]};
    auto loss_dseg = [&](const float (&p4)[4]) {       // LOSS: d_seg of the two pixels from the targets and the predictions p4
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cok = i ? c1ok : c0ok;
        const int c = 2 * t + i;
        ds[i] = (in0 && cok) ? fmaf(cf0[2 * c], lossp.mask[toff0 + (long long)c * lossp.mask_sc], cf0[2 * c + 1] * p4[i]) : 0.f;
        ds[2 + i] = (in1 && cok) ? fmaf(cf1[2 * c], lossp.mask[toff1 + (long long)c * lossp.mask_sc], cf1[2 * c + 1] * p4[2 + i]) : 0.f;
      }
    };
    if (LOSS && !do_softmax) { loss_dseg(lg); dlg[0] = ds[0]; dlg[1] = ds[1]; dlg[2] = ds[2]; dlg[3] = ds[3]; }
    if (do_softmax && (LOSS || d_seg)) {
      float m0 = fmaxf(c0ok ? lg[0] : -INFINITY, c1ok ? lg[1] : -INFINITY);
      float m1 = fmaxf(c0ok ? lg[2] : -INFINITY, c1ok ? lg[3] : -INFINITY);
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float pr[4] = {c0ok ? expf(lg[0] - m0) : 0.f, c1ok ? expf(lg[1] - m0) : 0.f,
                     c0ok ? expf(lg[2] - m1) : 0.f, c1ok ? expf(lg[3] - m1) : 0.f};
      float s0 = pr[0] + pr[1], s1 = pr[2] + pr[3];
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float i0 = 1.f / s0, i1 = 1.f / s1;
      pr[0] *= i0; pr[1] *= i0; pr[2] *= i1; pr[3] *= i1;
      if (LOSS) loss_dseg(pr);
      float d0 = fmaf(pr[0], ds[0], pr[1] * ds[1]), d1 = fmaf(pr[2], ds[2], pr[3] * ds[3]);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 2); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      dlg[0] = pr[0] * (ds[0] - d0); dlg[1] = pr[1] * (ds[1] - d0);
      dlg[2] = pr[2] * (ds[2] - d1); dlg[3] = pr[3] * (ds[3] - d1);
    }
    // dcat = dheat W21: tiles 0..3 -> dfeat (permuted columns), tile 4 -> into dlg
    float dc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dc[j][0] = dc[j][1] = dc[j][2] = dc[j][3] = 0.f;
    if (kL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bh[2], bl2[2];
        frag2(8 + 2 * j, bh); frag2(18 + 2 * j, bl2);
        mma_bf16_16816(dc[j], ah_hi, bh); mma_bf16_16816(dc[j], ah_hi, bl2); mma_bf16_16816(dc[j], ah_lo, bh);
      }
      {
        uint32_t bh[2], bl2[2];
        frag2(16, bh); frag2(26, bl2);
        mma_bf16_16816(dlg, ah_hi, bh); mma_bf16_16816(dlg, ah_hi, bl2); mma_bf16_16816(dlg, ah_lo, bh);
      }
    }
    // dfeat += dlg Wseg
    float r00, r01, r10, r11;
    const uint32_t al_hi[4] = {heads_pack_hi(dlg[0], dlg[1], r00, r01), heads_pack_hi(dlg[2], dlg[3], r10, r11), 0u, 0u};
    const uint32_t al_lo[4] = {heads_pack(r00, r01), heads_pack(r10, r11), 0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t fh = s_frag[28 + j][lane], fl = s_frag[32 + j][lane];
      mma_bf16_16816_b0(dc[j], al_hi, fh); mma_bf16_16816_b0(dc[j], al_hi, fl); mma_bf16_16816_b0(dc[j], al_lo, fh);
    }
    if (ok0) *reinterpret_cast<uint4*>(d_feat + (long long)r0 * d_ld + 8 * t) =
        make_uint4(heads_pack(dc[0][0], dc[0][1]), heads_pack(dc[1][0], dc[1][1]), heads_pack(dc[2][0], dc[2][1]), heads_pack(dc[3][0], dc[3][1]));
    if (ok1) *reinterpret_cast<uint4*>(d_feat + (long long)r1 * d_ld + 8 * t) =
        make_uint4(heads_pack(dc[0][2], dc[0][3]), heads_pack(dc[1][2], dc[1][3]), heads_pack(dc[2][2], dc[2][3]), heads_pack(dc[3][2], dc[3][3]));
    // ---- weight gradients: K = the 16 pixels of the tile ----
    // B' tiles: features (step s, half h): pixels 0..7 / 8..15 of that 8-channel block, transposed; tile 4: the logits
    uint32_t bt[kL ? 5 : 4][2];
    bt[0][0] = movm_trans(a_k0[0]); bt[0][1] = movm_trans(a_k0[1]);
    bt[1][0] = movm_trans(a_k0[2]); bt[1][1] = movm_trans(a_k0[3]);
    bt[2][0] = movm_trans(a_k1[0]); bt[2][1] = movm_trans(a_k1[1]);
    bt[3][0] = movm_trans(a_k1[2]); bt[3][1] = movm_trans(a_k1[3]);
    // A' (dlg^T): rows = classes (8 .. 15: zero), K = pixels
    const uint32_t at_s[4] = {movm_trans(al_hi[0]), 0u, movm_trans(al_hi[1]), 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) mma_bf16_16816(gs[j], at_s, bt[j]);
    if (kL) {
      bt[kL ? 4 : 0][0] = movm_trans(heads_pack(lg[0], lg[1])); bt[kL ? 4 : 0][1] = movm_trans(heads_pack(lg[2], lg[3]));
      // A' (dheat^T): rows = landmarks, K = pixels
      const uint32_t at_h[4] = {movm_trans(ah_hi[0]), movm_trans(ah_hi[2]), movm_trans(ah_hi[1]), movm_trans(ah_hi[3])};
#pragma unroll
      for (int j = 0; j < 5; ++j) mma_bf16_16816(g1[j], at_h, bt[j]);
    }
  }
  // ---- block reduction of the weight-gradient fragments, then one global atomic per element and block ----
  // fragment columns 2t, 2t+1 of feature tile (s, h) = 2s + h are channels 8t + 4s + 2h (+1); rows g, g + 8
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ch = 8 * t + 4 * (j >> 1) + 2 * (j & 1) + e;
      if (g < NC) atomicAdd(&s_g[NG1 + g * CF + ch], gs[j][e]);
      if (kL) {
        if (g < NL) atomicAdd(&s_g[g * NCAT + ch], g1[j][e]);
        if (g + 8 < NL) atomicAdd(&s_g[(g + 8) * NCAT + ch], g1[j][2 + e]);
      }
    }
  if (kL) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 2 * t + e;
      if (k < NC) {
        if (g < NL) atomicAdd(&s_g[g * NCAT + CF + k], g1[kL ? 4 : 0][e]);
        if (g + 8 < NL) atomicAdd(&s_g[(g + 8) * NCAT + CF + k], g1[kL ? 4 : 0][2 + e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NG; i += blockDim.x) atomicAdd(g_acc + i, s_g[i]);
}

// --------------------------------------------------------------------------
// First layer (C_in = 1, 32 output channels, bf16 storage) on warp-level tensor-core MMAs.  The CUDA-core kernels above spend
// 288 FMAs per pixel (3x3) and sit on the step's critical path at both ends: the 3x3 forward is the first kernel of the
// step, and its weight gradient can only start when the LAST data-gradient kernel of the backward has finished.
// A warp owns tiles of 16 consecutive pixels of an image row (W % 16 == 0):
//   forward:  y[16 pix][32] = patches[16 pix][9 taps -> K = 16] x W^T; the patch fragment is built from <= 3 scalar loads per
//             pixel of the (L1-resident) single-channel input; weights as split-bf16 pairs in registers; output columns permuted
//             so that a thread ends up with 8 contiguous channels of a pixel (one 16-byte store), bias / ReLU / BatchNorm
//             statistics on the accumulator fragments;
//   wgrad:    dW[32][9] = dY^T[32][16 pix] x patches[16 pix][9]; dY arrives as one 16-byte load per pixel row and is
//             transposed in registers (movmatrix), accumulators persist over the block's tiles.
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cin1_x_raw(const bf16* __restrict__ x, int x_ld, int n, int ih, int iw, int H, int W) {
  return (ih >= 0 && ih < H && iw >= 0 && iw < W)
             ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(x) + ((long long)(n * H + ih) * W + iw) * x_ld) : 0u;
}

// The (K x (16 + K - 1)) input window of a 16-pixel tile, zero padded, staged by the warp in its own shared-memory slice
// (<= 2 coalesced loads per lane instead of 8 scattered, predicated 2-byte gathers per lane and tile); read back as
// sw[r * 18 + c], c = pixel offset + kw.
template <int K>
__device__ __forceinline__ void cin1_stage_window(const bf16* __restrict__ x, int x_ld, unsigned short* sw, int n, int h, int w0, int H, int W,
                                                  int lane) {
  constexpr int PAD = K / 2, XW = 16 + 2 * PAD;
  __syncwarp();                                   // the previous tile's readers are done
#pragma unroll
  for (int q = 0; q < (K * XW + 31) / 32; ++q) {
    const int i = lane + 32 * q;
    if (i < K * XW) {
      const int r = i / XW, c = i - r * XW;
      sw[r * 18 + c] = (unsigned short)cin1_x_raw(x, x_ld, n, h + r - PAD, w0 + c - PAD, H, W);
    }
  }
  __syncwarp();
}

template <int K>
__global__ void __launch_bounds__(128) conv_cin1_mma_kernel(const bf16* __restrict__ x, int x_ld, bf16* __restrict__ y, int y_ld,
                                                            const float* __restrict__ w /* (32, 1, K, K) */, const float* bias,
                                                            int relu, double* stat, int B, int H, int W) {
  pdl_wait(); pdl_trigger();
  constexpr int KK = K * K, PAD = K / 2, CO = 32;
  static_assert(KK <= 16, "taps must fit one K step");
  __shared__ float part[4][2 * CO];
  __shared__ unsigned short s_win[4][3 * 18];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  unsigned short* sw = s_win[wrp];
  // B fragments: K = tap, column n = g of tile j <-> channel 8 (g / 2) + 2 j + (g & 1)
  uint32_t bw_hi[4][2], bw_lo[4][2];
  float bb[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = 8 * (g >> 1) + 2 * j + (g & 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int tap = 2 * t + 8 * h;
      const float a = tap < KK ? w[ch * KK + tap] : 0.f, b = tap + 1 < KK ? w[ch * KK + tap + 1] : 0.f;
      float ra, rb;
      bw_hi[j][h] = heads_pack_hi(a, b, ra, rb);
      bw_lo[j][h] = heads_pack(ra, rb);
    }
    // accumulator columns 2t, 2t+1 of tile j are channels 8t + 2j, 8t + 2j + 1
    bb[j][0] = bias ? bias[8 * t + 2 * j] : 0.f; bb[j][1] = bias ? bias[8 * t + 2 * j + 1] : 0.f;
  }
  float cs[4][2], cq[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { cs[j][0] = cs[j][1] = cq[j][0] = cq[j][1] = 0.f; }
  const int W16 = W >> 4;
  const int ntiles = B * H * W16;
  const int warp_g = blockIdx.x * 4 + wrp, nwarps = gridDim.x * 4;
  for (int tile = warp_g; tile < ntiles; tile += nwarps) {
    const int rr = tile / W16, w0 = (tile - rr * W16) * 16;
    const int n = rr / H, h = rr - n * H;
    // patch fragments: row g / g + 8 = pixels w0 + g / w0 + g + 8; K = taps 2t, 2t+1 | 2t+8, 2t+9
    uint32_t a[4];
    cin1_stage_window<K>(x, x_ld, sw, n, h, w0, H, W, lane);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int tap = 2 * t + (e & 1) + 8 * (e >> 1);
        v[e] = tap < KK ? (uint32_t)sw[(tap / K) * 18 + g + 8 * half + tap % K] : 0u;
      }
      a[half] = v[0] | (v[1] << 16);
      a[2 + half] = v[2] | (v[3] << 16);
    }
    const long long pix0 = ((long long)n * H + h) * W + w0;
    uint32_t o0[4], o1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(c, a, bw_hi[j]);
      mma_bf16_16816(c, a, bw_lo[j]);
      c[0] += bb[j][0]; c[1] += bb[j][1]; c[2] += bb[j][0]; c[3] += bb[j][1];
      if (relu) { c[0] = fmaxf(c[0], 0.f); c[1] = fmaxf(c[1], 0.f); c[2] = fmaxf(c[2], 0.f); c[3] = fmaxf(c[3], 0.f); }
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(c[0], c[1]), p1 = __floats2bfloat162_rn(c[2], c[3]);
      o0[j] = *reinterpret_cast<const uint32_t*>(&p0); o1[j] = *reinterpret_cast<const uint32_t*>(&p1);
      if (stat) {     // statistics of the STORED (rounded) values, as every other producer of a BatchNorm input does
        const float2 f0 = __bfloat1622float2(p0), f1 = __bfloat1622float2(p1);
        cs[j][0] += f0.x + f1.x; cs[j][1] += f0.y + f1.y;
        cq[j][0] = fmaf(f0.x, f0.x, fmaf(f1.x, f1.x, cq[j][0])); cq[j][1] = fmaf(f0.y, f0.y, fmaf(f1.y, f1.y, cq[j][1]));
      }
    }
    *reinterpret_cast<uint4*>(y + (pix0 + g) * y_ld + 8 * t) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
    *reinterpret_cast<uint4*>(y + (pix0 + g + 8) * y_ld + 8 * t) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
  }
  if (stat) {
    // lanes with equal t hold the same channels: fixed-order shuffle tree over g, then the block's four warps in order
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          cs[j][e] += __shfl_xor_sync(0xffffffffu, cs[j][e], o);
          cq[j][e] += __shfl_xor_sync(0xffffffffu, cq[j][e], o);
        }
    if (g == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          part[wrp][8 * t + 2 * j + e] = cs[j][e];
          part[wrp][CO + 8 * t + 2 * j + e] = cq[j][e];
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * CO) {
      const float v = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
      atomicAdd(&stat[threadIdx.x], (double)v);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) wgrad_cin1_mma_kernel(const bf16* __restrict__ x, int x_ld, const bf16* __restrict__ dy, int dy_ld,
                                                             float* dw /* (32, 1, K, K) */, int B, int H, int W) {
  pdl_wait(); pdl_trigger();
  constexpr int KK = K * K, PAD = K / 2, CO = 32, NT = KK > 8 ? 2 : 1;
  __shared__ float s_dw[CO * KK];
  __shared__ unsigned short s_win[4][3 * 18];
  for (int i = threadIdx.x; i < CO * KK; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  unsigned short* sw = s_win[wrp];
  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const int W16 = W >> 4;
  const int ntiles = B * H * W16;
  const int warp_g = blockIdx.x * 4 + wrp, nwarps = gridDim.x * 4;
  auto load_rows = [&](int tile, uint4& q0, uint4& q1) {
    q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0;
    if (tile < ntiles) {
      const long long pix0 = (long long)tile * 16;        // (tiles are numbered in pixel order: 16 consecutive pixels)
      q0 = *reinterpret_cast<const uint4*>(dy + (pix0 + g) * dy_ld + 8 * t);
      q1 = *reinterpret_cast<const uint4*>(dy + (pix0 + g + 8) * dy_ld + 8 * t);
    }
  };
  uint4 nq0, nq1;                             // dY rows one tile ahead (two ahead measured slower: 35.8 -> 47.5 us)
  load_rows(warp_g, nq0, nq1);
  for (int tile = warp_g; tile < ntiles; tile += nwarps) {
    const uint4 q0 = nq0, q1 = nq1;
    load_rows(tile + nwarps, nq0, nq1);
    const int rr = tile / W16, w0 = (tile - rr * W16) * 16;
    const int n = rr / H, h = rr - n * H;
    // dY^T: block j of a thread's 16-byte row = logical columns (2t, 2t+1) <-> channels 8t + 2j (+1); transposed, row g' of
    // block j is channel 8 (g' / 2) + 2 j + (g' & 1), K = pixels
    const uint32_t lo[4] = {movm_trans(q0.x), movm_trans(q0.y), movm_trans(q0.z), movm_trans(q0.w)};    // pixels 0..7
    const uint32_t hi[4] = {movm_trans(q1.x), movm_trans(q1.y), movm_trans(q1.z), movm_trans(q1.w)};    // pixels 8..15
    // patches as B operand: K = pixels 2t, 2t+1 | 2t+8, 2t+9, column n = tap g (tile 1: tap 8 in column 0)
    uint32_t b[NT][2];
    cin1_stage_window<K>(x, x_ld, sw, n, h, w0, H, W, lane);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int tap = 8 * j + g;
      uint32_t v[4] = {0u, 0u, 0u, 0u};
      if (tap < KK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (uint32_t)sw[(tap / K) * 18 + 2 * t + (e & 1) + 8 * (e >> 1) + tap % K];
      }
      b[j][0] = v[0] | (v[1] << 16); b[j][1] = v[2] | (v[3] << 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t a[4] = {lo[2 * i], lo[2 * i + 1], hi[2 * i], hi[2 * i + 1]};
#pragma unroll
      for (int j = 0; j < NT; ++j) mma_bf16_16816(acc[i][j], a, b[j]);
    }
  }
  // fragment (i, j): rows g (block 2i) and g + 8 (block 2i + 1) -> channels; columns 2t, 2t+1 -> taps 8j + 2t (+1)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int tap = 8 * j + 2 * t + e;
        if (tap < KK) {
          const int ch0 = 8 * (g >> 1) + 2 * (2 * i) + (g & 1), ch1 = 8 * (g >> 1) + 2 * (2 * i + 1) + (g & 1);
          atomicAdd(&s_dw[ch0 * KK + tap], acc[i][j][e]);
          atomicAdd(&s_dw[ch1 * KK + tap], acc[i][j][2 + e]);
        }
      }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * KK; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
}

// Per-plane coefficients of the loss gradient (closed forms of loss_backward_kernel) from the forward's sums and the
// upstream gradient of the loss: class planes (A, Bc), landmark planes (ka, kb, kc).  One thread per plane.
__global__ void loss_coef_kernel(const double* __restrict__ sums, const float* __restrict__ dloss, float* __restrict__ coef,
                                 int B, int NC, int NL, int Ht, int Wt, int skip_bg, float dice_wgt, float heat_wgt) {
  pdl_wait(); pdl_trigger();
  const int per = NC * 3 + NL * 5, cper = NC * 2 + NL * 3;
  const double up = (double)*dloss;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * (NC + NL); i += gridDim.x * blockDim.x) {
    const int b = i / (NC + NL), c = i - b * (NC + NL);
    if (c < NC) {
      const int c_first = skip_bg ? 1 : 0;
      const double* s = sums + (long long)b * per + c * 3;
      const double num = -2.0 * s[0] + 1.0e-4, den = s[1] + s[2] + 1.0e-4;
      const double k = c < c_first ? 0.0 : up * dice_wgt / ((double)(NC - c_first) * B) / (den * den);
      coef[(long long)b * cper + 2 * c] = (float)(-2.0 * k * den);
      coef[(long long)b * cper + 2 * c + 1] = (float)(-2.0 * k * num);
    } else {
      const int l = c - NC;
      const double N = (double)Ht * Wt;
      const double* m = sums + (long long)b * per + NC * 3 + l * 5;
      const double mx = m[0] / N, my = m[2] / N;
      const double vxx = fmax(m[1] - N * mx * mx, 0.0), vyy = fmax(m[3] - N * my * my, 0.0);
      const double sdx = sqrt(vxx / (N - 1.0)), sdy = sqrt(vyy / (N - 1.0));
      const double sxy_c = m[4] - N * mx * my, den = N * sdx * sdy + 1.0e-8;
      const double w = up * heat_wgt * -0.5 / ((double)B * NL);
      const double ka = w / den, kb = sdx > 0.0 ? w * sxy_c * N * sdy / ((N - 1.0) * sdx * den * den) : 0.0;
      float* o = coef + (long long)b * cper + NC * 2 + l * 3;
      o[0] = (float)ka; o[1] = (float)kb; o[2] = (float)(kb * mx - ka * my);
    }
  }
}

// dW2 = G1 W1^T, dW1 = W2^T G1, dWseg = Gseg  (G's accumulated by heads_bwd_fused_kernel)
__global__ void heads_bwd_finalize_kernel(const float* g_acc, const float* w1, const float* w2, float* dwseg, float* dw1,
                                          float* dw2, int CF, int NC, int NF, int NL) {
  pdl_wait(); pdl_trigger();
  const int NCAT = CF + NC;
  const float* g1 = g_acc;
  const float* gs = g_acc + (NL > 0 ? NL * NCAT : 0);
  const int n_seg = NC * CF, n_w1 = NL > 0 ? NF * NCAT : 0, n_w2 = NL > 0 ? NL * NF : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seg + n_w1 + n_w2; i += gridDim.x * blockDim.x) {
    if (i < n_seg) {
      dwseg[i] = gs[i];
    } else if (i < n_seg + n_w1) {
      const int j = i - n_seg, m = j / NCAT, c = j - m * NCAT;
      float a = 0.f;
      for (int l = 0; l < NL; ++l) a = fmaf(w2[l * NF + m], g1[l * NCAT + c], a);
      dw1[j] = a;
    } else {
      const int j = i - n_seg - n_w1, l = j / NF, m = j - l * NF;
      float a = 0.f;
      for (int c = 0; c < NCAT; ++c) a = fmaf(g1[l * NCAT + c], w1[m * NCAT + c], a);
      dw2[j] = a;
    }
  }
}

// --------------------------------------------------------------------------
// Weight gradient: dW[tap][cb][cs] += sum_m BIG[pix(m)@tap, cb] * SMALL[m, cs]
// (split over pixel ranges, fp32 atomics into a zeroed buffer, arbitrary output
// strides so the result lands directly in the torch parameter layout).
// --------------------------------------------------------------------------
struct WgradArgs {
  const void* big; int big_ld; int Hb, Wb, Cb;
  const void* small; int small_ld; int Hs, Ws, Cs;
  int B, KH, KW, stride, pad;
  float* dw; long long s_tap, s_big, s_small;
  int pix_per_split;   // multiple of 16
  int tiles_small;     // number of 64-wide tiles over Cs
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradArgs p) {
  pdl_wait(); pdl_trigger();
  constexpr int BK = 16, TP = 64 + 4;
  __shared__ __align__(16) float Bg[BK][TP];
  __shared__ __align__(16) float Sm[BK][TP];
  const int tid = threadIdx.x;
  const int tile_b = blockIdx.x / p.tiles_small, tile_s = blockIdx.x - tile_b * p.tiles_small;
  const int cb0 = tile_b * 64, cs0 = tile_s * 64;
  const int tap = blockIdx.y;
  const int kh = tap / p.KW, kw = tap - kh * p.KW;
  const long long M = (long long)p.B * p.Hs * p.Ws;
  const long long m_begin = (long long)blockIdx.z * p.pix_per_split;
  long long m_end = m_begin + p.pix_per_split;
  if (m_end > M) m_end = M;
  const T* bigp = reinterpret_cast<const T*>(p.big);
  const T* smallp = reinterpret_cast<const T*>(p.small);
  const int l_pix = tid >> 4, l_c = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 rb, rs;
  auto gload = [&](long long mb) {
    const long long m = mb + l_pix;
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    rs = rb;
    if (m < m_end) {
      const int ow = (int)(m % p.Ws);
      const long long t = m / p.Ws;
      const int oh = (int)(t % p.Hs);
      const int n = (int)(t / p.Hs);
      const int cs = cs0 + l_c;
      const T* ss = smallp + m * p.small_ld + cs;
      if (VEC) {
        if (cs < p.Cs) rs = ld4(ss);
      } else {
        if (cs < p.Cs) rs.x = ld1(ss);
        if (cs + 1 < p.Cs) rs.y = ld1(ss + 1);
        if (cs + 2 < p.Cs) rs.z = ld1(ss + 2);
        if (cs + 3 < p.Cs) rs.w = ld1(ss + 3);
      }
      const int ih = oh * p.stride + kh - p.pad, iw = ow * p.stride + kw - p.pad;
      if (ih >= 0 && ih < p.Hb && iw >= 0 && iw < p.Wb) {
        const int cb = cb0 + l_c;
        const T* bs = bigp + ((long long)(n * p.Hb + ih) * p.Wb + iw) * p.big_ld + cb;
        if (VEC) {
          if (cb < p.Cb) rb = ld4(bs);
        } else {
          if (cb < p.Cb) rb.x = ld1(bs);
          if (cb + 1 < p.Cb) rb.y = ld1(bs + 1);
          if (cb + 2 < p.Cb) rb.z = ld1(bs + 2);
          if (cb + 3 < p.Cb) rb.w = ld1(bs + 3);
        }
      }
    }
  };
  if (m_begin < m_end) gload(m_begin);
  for (long long mb = m_begin; mb < m_end; mb += BK) {
    *reinterpret_cast<float4*>(&Bg[l_pix][l_c]) = rb;
    *reinterpret_cast<float4*>(&Sm[l_pix][l_c]) = rs;
    __syncthreads();
    if (mb + BK < m_end) gload(mb + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Bg[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Sm[k][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cb = cb0 + ty * 4 + i;
    if (cb >= p.Cb) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cs = cs0 + tx * 4 + j;
      if (cs < p.Cs) atomicAdd(p.dw + tap * p.s_tap + cb * p.s_big + cs * p.s_small, acc[i][j]);
    }
  }
}

// --------------------------------------------------------------------------
// Weight packing: dst[t][k][nh*Ninner + nl] = src[tmap(t)*st + k*sk + nh*snh + nl*snl]
// (columns >= N are zero).  One kernel covers every layout in the engine.
// --------------------------------------------------------------------------
struct PackArgs {
  const float* src; float* dst;
  int T, K, N, Npad, Ninner, flip;
  long long st, sk, snh, snl;
};
__global__ void pack_weights_kernel(const PackArgs p) {
  pdl_wait(); pdl_trigger();
  const long long total = (long long)p.T * p.K * p.Npad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % p.Npad);
    const long long r = i / p.Npad;
    const int k = (int)(r % p.K);
    const int t = (int)(r / p.K);
    float v = 0.f;
    if (n < p.N) {
      const int tm = p.flip ? (p.T - 1 - t) : t;
      const int nh = n / p.Ninner, nl = n - nh * p.Ninner;
      v = p.src[tm * p.st + k * p.sk + nh * p.snh + nl * p.snl];
    }
    p.dst[i] = v;
  }
}

// --------------------------------------------------------------------------
// BatchNorm (nn.BatchNorm2d, unet.py:214-215,221-222)
// --------------------------------------------------------------------------
// stat[0:C] = sum x, stat[C:2C] = sum x^2 over the P = B*H*W positions (train) -> per-channel
// mean / invstd, affine fold a = gamma*invstd, b = beta - mean*a, running-stat update.
__global__ void bn_finalize_kernel(const double* stat, long long P, int C, int training,
                                   const float* gamma, const float* beta, float* rmean, float* rvar,
                                   long long* nbt, float momentum, float eps,
                                   float* mean_o, float* invstd_o, float* a_o, float* b_o) {
  pdl_wait(); pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = stat[c] / (double)P;
    double v = stat[C + c] / (double)P - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    const double unb = P > 1 ? v * ((double)P / (double)(P - 1)) : v;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  } else {
    mean = rmean[c];
    var = rvar[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float a = gamma[c] * invstd;
  mean_o[c] = mean;
  invstd_o[c] = invstd;
  a_o[c] = a;
  b_o[c] = beta[c] - mean * a;
}

// z = a[c]*r + b[c]   (BatchNorm apply with the folded scale / shift)
template <typename T>
__global__ void bn_apply_kernel(const T* r, int r_ld, T* z, int z_ld, const float* a, const float* b,
                                long long P, int C);
// BatchNorm forward finalise + apply in one launch: every thread derives mean / invstd / scale / shift of its
// channels from stat = [sum x | sum x^2] (train) or the running statistics (eval); block 0 also publishes
// mean / invstd (needed by the backward), and in training updates running_mean / running_var / num_batches_tracked.
struct BnFwdFin {
  const double* stat; const float* gamma; const float* beta; float* rmean; float* rvar; long long* nbt;
  float* mean_o; float* invstd_o; float* a_o; float* b_o; int training; float momentum, eps;
};
template <typename T>
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(const T* r, int r_ld, T* z, int z_ld, const BnFwdFin f,
                                                                long long P, int C);

// Eval mode: the folded scale / shift of EVERY BatchNorm layer of the network from its running statistics, one block per
// layer (nn.BatchNorm2d in eval(), unet.py:214-215,221-222): a = gamma / sqrt(running_var + eps), b = beta - mean * a.
// They are known before the convolutions run, so the tensor-core epilogues apply them right behind the ReLU.
struct BnEvalTable {
  enum { kMax = 48 };
  const float* gamma[kMax]; const float* beta[kMax]; const float* rmean[kMax]; const float* rvar[kMax];
  float* a[kMax]; float* b[kMax]; float* mean_o[kMax]; float* invstd_o[kMax];
  int C[kMax]; int count; float eps;
};
__global__ void bn_eval_coeffs_kernel(const BnEvalTable t) {
  pdl_wait(); pdl_trigger();
  const int e = blockIdx.x;
  if (e >= t.count) return;
  for (int i = threadIdx.x; i < t.C[e]; i += blockDim.x) {
    const float mean = t.rmean[e][i], invstd = rsqrtf(t.rvar[e][i] + t.eps);
    const float a = t.gamma[e][i] * invstd;
    t.a[e][i] = a; t.b[e][i] = t.beta[e][i] - mean * a;
    t.mean_o[e][i] = mean; t.invstd_o[e][i] = invstd;
  }
}

// 16-byte vector access per storage type: 4 fp32 or 8 bf16 channels per thread
// 16-byte vectors of the two storage types.  `raw`/`unpack` split a load from its conversion so that several
// loads can be in flight while costing 4 registers each.
__device__ __forceinline__ uint4 ld_raw16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ void load(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& u, float* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static __forceinline__ void load(const bf16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static __forceinline__ void store(bf16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__global__ void bn_apply_kernel(const T* r, int r_ld, T* z, int z_ld, const float* a, const float* b,
                                long long P, int C) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  const int cv = C / V;
  const long long total = P * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * V;
    const long long pix = i / cv;
    float v[V];
    Vec<T>::load(r + pix * r_ld + c, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = fmaf(a[c + k], v[k], b[c + k]);
    Vec<T>::store(z + pix * z_ld + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(const T* r, int r_ld, T* z, int z_ld, const BnFwdFin f,
                                                                long long P, int C) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  const int cvecs = C / V;
  const int lanes = cvecs < 256 ? cvecs : 256;
  const int rows = 256 / lanes;
  const int lane_c = threadIdx.x % lanes, prow = threadIdx.x / lanes;
  if (blockIdx.x == 0 && threadIdx.x == 0 && f.training && f.nbt) *f.nbt += 1;
  for (int cv = lane_c; cv < cvecs; cv += lanes) {
    const int c = cv * V;
    float a[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float mean, var;
      double unb = 0.0;
      if (f.training) {
        const double m = f.stat[c + i] / (double)P;
        double v = f.stat[C + c + i] / (double)P - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        unb = P > 1 ? v * ((double)P / (double)(P - 1)) : v;
      } else {
        mean = f.rmean[c + i]; var = f.rvar[c + i];
      }
      const float invstd = rsqrtf(var + f.eps);
      a[i] = f.gamma[c + i] * invstd;
      b[i] = f.beta[c + i] - mean * a[i];
      if (blockIdx.x == 0 && prow == 0) {
        f.mean_o[c + i] = mean; f.invstd_o[c + i] = invstd; f.a_o[c + i] = a[i]; f.b_o[c + i] = b[i];
      }
    }
    const long long step = (long long)gridDim.x * rows;
    long long pix = (long long)blockIdx.x * rows + prow;
    for (; pix + 3 * step < P; pix += 4 * step) {      // four 16-byte loads in flight per thread
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = ld_raw16(r + (pix + u * step) * r_ld + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[V];
        Vec<T>::unpack(q[u], v);
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = fmaf(a[i], v[i], b[i]);
        Vec<T>::store(z + (pix + u * step) * z_ld + c, v);
      }
    }
    for (; pix < P; pix += step) {
      float v[V];
      Vec<T>::load(r + pix * r_ld + c, v);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = fmaf(a[i], v[i], b[i]);
      Vec<T>::store(z + pix * z_ld + c, v);
    }
  }
  // running statistics: after every block has read them (eval never writes; train never reads them above)
  if (blockIdx.x == 0 && f.training) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double m = f.stat[c] / (double)P;
      double v = f.stat[C + c] / (double)P - m * m;
      if (v < 0.0) v = 0.0;
      const double unb = P > 1 ? v * ((double)P / (double)(P - 1)) : v;
      f.rmean[c] = (1.f - f.momentum) * f.rmean[c] + f.momentum * (float)m;
      f.rvar[c] = (1.f - f.momentum) * f.rvar[c] + f.momentum * (float)unb;
    }
  }
}

// Shared thread mapping of the per-channel reductions: lanes = min(C/VEC, kRedLanes) channel vectors across,
// 256/lanes pixel rows down; grid.y covers the channel groups beyond kRedLanes vectors.  (C and VEC are powers of two,
// C >= VEC.)  A block spans at most 8 vectors = one 128-byte line of channels: every block ends in one fp64 atomic
// per channel it covers, and with blocks that spanned all C channels the 512 / 1024-channel levels issued ~0.5 M
// atomics per kernel on 2-5 MB tensors (25-30 us per BatchNorm backward at 12x12 where the data moves in 3 us).
constexpr int kRedLanes = 8;
template <int VEC>
struct RedMap {
  int lanes, rows, cv, prow;
  __device__ RedMap(int C) {
    const int cvecs = C / VEC;
    lanes = cvecs < kRedLanes ? cvecs : kRedLanes;
    rows = 256 / lanes;
    cv = blockIdx.y * lanes + (threadIdx.x % lanes);
    prow = threadIdx.x / lanes;
  }
};

// Per-channel reduction targets are kept in kRedCopies copies, `stride` doubles apart: block b adds into copy
// b % kRedCopies and the consumer sums the copies.  With ONE copy every block's atomics on a channel hit the same
// address and serialise in L2 -- 1152 blocks x 64 addresses made that the tail of every reduction kernel (ncu: 12-16 us
// floors on 2-10 MB tensors, 2 TB/s at 96x96); eight copies cut the chain per address eightfold.
constexpr int kRedCopies = 8;
__device__ __forceinline__ double red_sum_copies(const double* p, int stride) {
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < kRedCopies; ++k) v += p[(size_t)k * stride];
  return v;
}

// block-level sum over the pixel rows of each channel, then one fp64 atomic per channel per block (into the block's copy)
template <int VEC>
__device__ __forceinline__ void block_channel_reduce(const RedMap<VEC>& mp, const float* s, double* out, int C, float* sm) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) sm[threadIdx.x * VEC + i] = s[i];
  __syncthreads();
  if (mp.prow == 0 && mp.cv * VEC < C) {
    float t[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = s[i];
    for (int r = 1; r < mp.rows; ++r) {
      const float* o = sm + (r * mp.lanes + (threadIdx.x % mp.lanes)) * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] += o[i];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomicAdd(out + mp.cv * VEC + i, (double)t[i]);
  }
  __syncthreads();
}
// copy of a [n]-double target this block adds into (copies are `stride` doubles apart)
__device__ __forceinline__ double* red_copy(double* base, int stride) { return base + (size_t)(blockIdx.x % kRedCopies) * stride; }

// out[0:C] += sum d ; out[C:2C] += sum d * xhat, xhat = (r - mean) * invstd
template <typename T>
__global__ void __launch_bounds__(256, 4) bn_bwd_reduce_kernel(const T* d, int d_ld, const T* r, int r_ld,
                                                            const float* mean, const float* invstd,
                                                            long long P, int C, double* out) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  __shared__ float sm[256 * V];
  RedMap<V> mp(C);
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  if (mp.cv * V < C) {
    const int c = mp.cv * V;
    // the loop accumulates sum d * (r - mean); invstd multiplies once at the end (fewer live registers: 4 blocks/SM)
    float mu[V];
#pragma unroll
    for (int i = 0; i < V; ++i) mu[i] = mean[c + i];
    const long long step = (long long)gridDim.x * mp.rows;
    long long pix = (long long)blockIdx.x * mp.rows + mp.prow;
    // two pixels in flight per thread (four measured slower: 514 -> 595 us per step over the 22 layers)
    for (; pix + step < P; pix += 2 * step) {
      float d0[V], r0[V], d1[V], r1[V];
      Vec<T>::load(d + pix * d_ld + c, d0); Vec<T>::load(r + pix * r_ld + c, r0);
      Vec<T>::load(d + (pix + step) * d_ld + c, d1); Vec<T>::load(r + (pix + step) * r_ld + c, r1);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s1[i] += d0[i] + d1[i];
        s2[i] = fmaf(d0[i], r0[i] - mu[i], fmaf(d1[i], r1[i] - mu[i], s2[i]));
      }
    }
    for (; pix < P; pix += step) {
      float d0[V], r0[V];
      Vec<T>::load(d + pix * d_ld + c, d0); Vec<T>::load(r + pix * r_ld + c, r0);
#pragma unroll
      for (int i = 0; i < V; ++i) { s1[i] += d0[i]; s2[i] = fmaf(d0[i], r0[i] - mu[i], s2[i]); }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) s2[i] *= invstd[c + i];
  }
  double* oc = red_copy(out, 2 * C);
  block_channel_reduce<V>(mp, s1, oc, C, sm);
  block_channel_reduce<V>(mp, s2, oc + C, C, sm);
}

// bstat -> d(gamma), d(beta) [, d(res bias) = sum d], and the coefficients of the
// elementwise pass: ga = gamma*invstd, m1 = mean(d), m2 = mean(d*xhat) (0 in eval mode).
__global__ void bn_bwd_finalize_kernel(const double* bstat, long long P, int C, int training,
                                       const float* gamma, const float* invstd, float* g_gamma,
                                       float* g_beta, float* g_extra, float* ga, float* m1, float* m2) {
  pdl_wait(); pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = bstat[c], s2 = bstat[C + c];
  g_gamma[c] = (float)s2;
  g_beta[c] = (float)s1;
  if (g_extra) g_extra[c] = (float)s1;
  ga[c] = gamma[c] * invstd[c];
  m1[c] = training ? (float)(s1 / (double)P) : 0.f;
  m2[c] = training ? (float)(s2 / (double)P) : 0.f;
}

// dy = relu'(r) * (has_bn ? ga*(d - m1 - xhat*m2) : d); out[0:C] += sum dy  (= conv bias gradient).
// With has_bn the kernel finalises the BN backward itself from bstat = [sum d | sum d*xhat] (written by
// bn_bwd_reduce_kernel): ga = gamma*invstd, m1 = mean(d), m2 = mean(d*xhat) (0 in eval mode); block (0,y)
// also stores d(gamma) = sum d*xhat, d(beta) = sum d [and d(res bias) = sum d].
struct BnBwdFin {
  const double* bstat; const float* gamma; float* g_gamma; float* g_beta; float* g_extra; int training;
};
template <typename T>
__global__ void __launch_bounds__(256, 3) act_bwd_kernel(const T* d, int d_ld, const T* r, int r_ld,
                                                      T* dy, int dy_ld, int has_bn, const float* mean,
                                                      const float* invstd, const BnBwdFin fin, long long P, int C,
                                                      double* out) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  __shared__ float sm[3 * 256 * V];      // [3][lanes*V] constants, then the [256*V] reduction scratch
  RedMap<V> mp(C);
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  // dy = A*d + Bc*r + Cc where r > 0:  A = gamma*invstd, Bc = -A*invstd*m2, Cc = -A*m1 + A*invstd*m2*mean
  // (m1 = mean(d), m2 = mean(d*xhat)); without BN: A = 1, Bc = Cc = 0.  Three constants per channel, computed once
  // per block by the pixel-row-0 threads (they sum the kRedCopies copies of the statistics) and shared.
  const bool active = mp.cv * V < C;
  const int c = mp.cv * V;
  float A[V], Bc[V], Cc[V];
  {
    float* sk = sm;                                  // [3][lanes*V] (sm is reused by the reduction at the end)
    const int nl = mp.lanes * V, li = (threadIdx.x % mp.lanes) * V;
    if (has_bn && active && mp.prow == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float mu = mean[c + i], is = invstd[c + i];
        const double s1 = red_sum_copies(fin.bstat + c + i, 2 * C), s2 = red_sum_copies(fin.bstat + C + c + i, 2 * C);
        const float a1 = fin.training ? (float)(s1 / (double)P) : 0.f;
        const float a2 = fin.training ? (float)(s2 / (double)P) : 0.f;
        const float a = fin.gamma[c + i] * is, b = -a * is * a2;
        sk[li + i] = a; sk[nl + li + i] = b; sk[2 * nl + li + i] = -a * a1 - b * mu;
        if (blockIdx.x == 0) {
          fin.g_gamma[c + i] = (float)s2;
          fin.g_beta[c + i] = (float)s1;
          if (fin.g_extra) fin.g_extra[c + i] = (float)s1;
        }
      }
    }
    if (has_bn) __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      A[i] = has_bn ? sk[li + i] : 1.f; Bc[i] = has_bn ? sk[nl + li + i] : 0.f; Cc[i] = has_bn ? sk[2 * nl + li + i] : 0.f;
    }
    if (has_bn) __syncthreads();                     // sm is written again by block_channel_reduce
  }
  const long long step = (long long)gridDim.x * mp.rows;
  const long long pix0 = (long long)blockIdx.x * mp.rows + mp.prow;
  if (active && pix0 < P) {
    auto one = [&](const float* dv, const float* rv, T* dst) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        o[i] = rv[i] > 0.f ? fmaf(A[i], dv[i], fmaf(Bc[i], rv[i], Cc[i])) : 0.f;
      }
      Vec<T>::store(dst, o);
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += rnd(o[i], dst);
    };
    // pixels pix0 + k*step, k < n, walked from the LAST to the first: bn_bwd_reduce_kernel (same grid, ascending
    // order) has just read this tensor pair, and what it read last is what the L2 still holds
    const long long n = (P - pix0 + step - 1) / step;
    long long k = n - 1;
    for (; k >= 1; k -= 2) {
      const long long pa = pix0 + k * step, pb = pa - step;
      float d0[V], r0[V], d1[V], r1[V];
      Vec<T>::load(d + pa * d_ld + c, d0); Vec<T>::load(r + pa * r_ld + c, r0);
      Vec<T>::load(d + pb * d_ld + c, d1); Vec<T>::load(r + pb * r_ld + c, r1);
      one(d0, r0, dy + pa * dy_ld + c);
      one(d1, r1, dy + pb * dy_ld + c);
    }
    if (k == 0) {
      float d0[V], r0[V];
      Vec<T>::load(d + pix0 * d_ld + c, d0); Vec<T>::load(r + pix0 * r_ld + c, r0);
      one(d0, r0, dy + pix0 * dy_ld + c);
    }
  }
  block_channel_reduce<V>(mp, s, red_copy(out, C), C, sm);
}

// Grid-wide barrier for kernels whose whole grid is resident (the launcher sizes the grid from the occupancy query).
// bar[0] = arrival count, bar[1] = generation; the last block to arrive resets the count and bumps the generation, so
// the same two words serve every launch of a stream (and every replay of a captured graph) without a host-side reset.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned* genp = bar + 1;
    const unsigned gen = *genp;
    __threadfence();                                  // this block's atomics / stores before its arrival
    if (atomicAdd(bar, 1u) == nblocks - 1u) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*genp == gen) __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

// BatchNorm backward + ReLU backward in ONE launch (bn_bwd_reduce_kernel + act_bwd_kernel with a grid barrier between
// them): phase 1 adds [sum d | sum d*xhat] into bstat, phase 2 finalises the coefficients from it and writes dy,
// walking the pixels in the OPPOSITE order so that what phase 1 read last (and the L2 still holds) is read first.
// Two launches per BatchNorm cost 12-17 us each on the 24x24 ... 6x6 levels (2-10 MB tensors, latency bound), and on
// the large levels the second pass found nothing of the first in L2.  The grid must be co-resident.
template <typename T>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_coop_kernel(const T* d, int d_ld, const T* r, int r_ld,
                                                              T* dy, int dy_ld, const float* mean, const float* invstd,
                                                              const BnBwdFin fin, double* bstat, long long P, int C,
                                                              double* out, unsigned* bar) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  __shared__ float sm[3 * 256 * V];
  RedMap<V> mp(C);
  const bool active = mp.cv * V < C;
  const int c = mp.cv * V;
  const long long step = (long long)gridDim.x * mp.rows;
  const long long pix0 = (long long)blockIdx.x * mp.rows + mp.prow;
  {
    float s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    if (active) {
      float mu[V];
#pragma unroll
      for (int i = 0; i < V; ++i) mu[i] = mean[c + i];
      long long pix = pix0;
      for (; pix + step < P; pix += 2 * step) {
        float d0[V], r0[V], d1[V], r1[V];
        Vec<T>::load(d + pix * d_ld + c, d0); Vec<T>::load(r + pix * r_ld + c, r0);
        Vec<T>::load(d + (pix + step) * d_ld + c, d1); Vec<T>::load(r + (pix + step) * r_ld + c, r1);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s1[i] += d0[i] + d1[i];
          s2[i] = fmaf(d0[i], r0[i] - mu[i], fmaf(d1[i], r1[i] - mu[i], s2[i]));
        }
      }
      for (; pix < P; pix += step) {
        float d0[V], r0[V];
        Vec<T>::load(d + pix * d_ld + c, d0); Vec<T>::load(r + pix * r_ld + c, r0);
#pragma unroll
        for (int i = 0; i < V; ++i) { s1[i] += d0[i]; s2[i] = fmaf(d0[i], r0[i] - mu[i], s2[i]); }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) s2[i] *= invstd[c + i];
    }
    double* oc = red_copy(bstat, 2 * C);
    block_channel_reduce<V>(mp, s1, oc, C, sm);
    block_channel_reduce<V>(mp, s2, oc + C, C, sm);
  }
  grid_barrier(bar, gridDim.x * gridDim.y);
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  float A[V], Bc[V], Cc[V];
  {
    // (see act_bwd_kernel) dy = A*d + Bc*r + Cc where r > 0
    float* sk = sm;
    const int nl = mp.lanes * V, li = (threadIdx.x % mp.lanes) * V;
    if (active && mp.prow == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float mu = mean[c + i], is = invstd[c + i];
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < kRedCopies; ++k) {        // written by other SMs' atomics: read at the L2
          t1 += __ldcg(bstat + (size_t)k * 2 * C + c + i);
          t2 += __ldcg(bstat + (size_t)k * 2 * C + C + c + i);
        }
        const float a1 = fin.training ? (float)(t1 / (double)P) : 0.f;
        const float a2 = fin.training ? (float)(t2 / (double)P) : 0.f;
        const float a = fin.gamma[c + i] * is, b = -a * is * a2;
        sk[li + i] = a; sk[nl + li + i] = b; sk[2 * nl + li + i] = -a * a1 - b * mu;
        if (blockIdx.x == 0) {
          fin.g_gamma[c + i] = (float)t2;
          fin.g_beta[c + i] = (float)t1;
          if (fin.g_extra) fin.g_extra[c + i] = (float)t1;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) { A[i] = sk[li + i]; Bc[i] = sk[nl + li + i]; Cc[i] = sk[2 * nl + li + i]; }
    __syncthreads();
  }
  if (active && pix0 < P) {
    auto one = [&](const float* dv, const float* rv, T* dst) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = rv[i] > 0.f ? fmaf(A[i], dv[i], fmaf(Bc[i], rv[i], Cc[i])) : 0.f;
      Vec<T>::store(dst, o);
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += rnd(o[i], dst);
    };
    const long long n = (P - pix0 + step - 1) / step;          // pixels of this thread: pix0 + k*step, k < n
    long long k = n - 1;
    for (; k >= 1; k -= 2) {
      const long long pa = pix0 + k * step, pb = pa - step;
      float d0[V], r0[V], d1[V], r1[V];
      Vec<T>::load(d + pa * d_ld + c, d0); Vec<T>::load(r + pa * r_ld + c, r0);
      Vec<T>::load(d + pb * d_ld + c, d1); Vec<T>::load(r + pb * r_ld + c, r1);
      one(d0, r0, dy + pa * dy_ld + c);
      one(d1, r1, dy + pb * dy_ld + c);
    }
    if (k == 0) {
      float d0[V], r0[V];
      Vec<T>::load(d + pix0 * d_ld + c, d0); Vec<T>::load(r + pix0 * r_ld + c, r0);
      one(d0, r0, dy + pix0 * dy_ld + c);
    }
  }
  block_channel_reduce<V>(mp, s, red_copy(out, C), C, sm);
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* d, int d_ld, long long P, int C,
                                                          double* out) {
  pdl_wait(); pdl_trigger();
  constexpr int V = Vec<T>::N;
  __shared__ float sm[256 * V];
  RedMap<V> mp(C);
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  if (mp.cv * V < C) {
    const int c = mp.cv * V;
    for (long long pix = (long long)blockIdx.x * mp.rows + mp.prow; pix < P;
         pix += (long long)gridDim.x * mp.rows) {
      float dv[V];
      Vec<T>::load(d + pix * d_ld + c, dv);
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += dv[i];
    }
  }
  block_channel_reduce<V>(mp, s, red_copy(out, C), C, sm);
}

// all bias-gradient accumulators of a backward pass -> fp32 gradients, one launch
struct SumTable {
  enum { kMax = 96 };
  const double* src[kMax]; float* dst[kMax]; int n[kMax]; int copies[kMax]; int count;   // copies: 1 or kRedCopies (n doubles apart)
};
__global__ void sums_to_float_kernel(const SumTable t) {
  pdl_wait(); pdl_trigger();
  const int e = blockIdx.x;
  if (e >= t.count) return;
  for (int i = threadIdx.x; i < t.n[e]; i += blockDim.x) {
    double v = 0.0;
    for (int k = 0; k < t.copies[e]; ++k) v += t.src[e][(size_t)k * t.n[e] + i];
    t.dst[e][i] = (float)v;
  }
}

__global__ void sum_to_float_kernel(const double* src, float* dst, int C) {
  pdl_wait(); pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dst[c] = (float)src[c];
}

// --------------------------------------------------------------------------
// F.max_pool2d(x, 2) (unet.py:169) and its backward: the first maximum in
// row-major window order receives the gradient.
// --------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_fwd_kernel(const T* x, int x_ld, T* y, int y_ld, int B, int Ho, int Wo, int C) {
  pdl_wait(); pdl_trigger();
  const int cv = C >> 2;
  const long long total = (long long)B * Ho * Wo * cv;
  const int Wi = Wo * 2, Hi = Ho * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 4;
    long long pix = i / cv;
    const int ow = (int)(pix % Wo);
    const long long t = pix / Wo;
    const int oh = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const T* s = x + ((long long)(n * Hi + 2 * oh) * Wi + 2 * ow) * x_ld + c;
    const float4 v0 = ld4(s), v1 = ld4(s + x_ld), v2 = ld4(s + (long long)Wi * x_ld),
                 v3 = ld4(s + (long long)(Wi + 1) * x_ld);
    float4 m;
    m.x = fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x));
    m.y = fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y));
    m.z = fmaxf(fmaxf(v0.z, v1.z), fmaxf(v2.z, v3.z));
    m.w = fmaxf(fmaxf(v0.w, v1.w), fmaxf(v2.w, v3.w));
    st4(y + pix * y_ld + c, m);
  }
}

__device__ __forceinline__ int first_max4(float a, float b, float c, float d) {
  int k = 0; float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  return k;
}

// dx (B,2Ho,2Wo,C) (+)= route(dy); accumulate=1 adds onto the existing dx (the skip gradient).
template <typename T>
__global__ void maxpool_bwd_kernel(const T* x, int x_ld, const T* dy, int dy_ld, T* dx, int dx_ld,
                                   int B, int Ho, int Wo, int C, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int cv = C >> 2;
  const long long total = (long long)B * Ho * Wo * cv;
  const int Wi = Wo * 2, Hi = Ho * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 4;
    long long pix = i / cv;
    const int ow = (int)(pix % Wo);
    const long long t = pix / Wo;
    const int oh = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const long long base = (long long)(n * Hi + 2 * oh) * Wi + 2 * ow;
    const long long off[4] = {base, base + 1, base + Wi, base + Wi + 1};
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld4(x + off[k] * x_ld + c);
    const float4 g = ld4(dy + pix * dy_ld + c);
    const int kx = first_max4(v[0].x, v[1].x, v[2].x, v[3].x);
    const int ky = first_max4(v[0].y, v[1].y, v[2].y, v[3].y);
    const int kz = first_max4(v[0].z, v[1].z, v[2].z, v[3].z);
    const int kw = first_max4(v[0].w, v[1].w, v[2].w, v[3].w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f,
                             kw == k ? g.w : 0.f);
      T* dst = dx + off[k] * dx_ld + c;
      if (accumulate) {
        const float4 old = ld4(dst);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      st4(dst, o);
    }
  }
}

// --------------------------------------------------------------------------
// Boundary layout changes and the softmax head (nn.Softmax2d, unet.py:103-104,179)
// --------------------------------------------------------------------------
// fp32 NCHW (B,C,H,W) -> T NHWC with pixel stride ld
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* src, T* dst, int ld, int B, int C, long long HW) {
  pdl_wait(); pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    for (int c = 0; c < C; ++c) st1(dst + i * ld + c, src[(n * C + c) * HW + hw]);
  }
}

// T NHWC (pixel stride ld) -> fp32 NCHW (debug / per-layer parity read-back)
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* src, int ld, float* dst, int B, int C, long long HW) {
  pdl_wait(); pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    for (int c = 0; c < C; ++c) dst[(n * C + c) * HW + hw] = ld1(src + i * ld + c);
  }
}

constexpr int kMaxClasses = 32;

// logits (T NHWC, stride ld) -> seg (fp32 NCHW) [softmax over channels], optional fp32 NCHW logits copy
template <typename T>
__global__ void softmax_fwd_kernel(const T* logits, int ld, int B, int ncls, long long HW, int do_softmax,
                                   float* seg, float* logits_out) {
  pdl_wait(); pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    float v[kMaxClasses];
    float mx = -INFINITY;
    for (int c = 0; c < ncls; ++c) {
      v[c] = ld1(logits + i * ld + c);
      mx = fmaxf(mx, v[c]);
    }
    if (logits_out)
      for (int c = 0; c < ncls; ++c) logits_out[(n * ncls + c) * HW + hw] = v[c];
    if (do_softmax) {
      float s = 0.f;
      for (int c = 0; c < ncls; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
      const float inv = 1.f / s;
      for (int c = 0; c < ncls; ++c) v[c] *= inv;
    }
    for (int c = 0; c < ncls; ++c) seg[(n * ncls + c) * HW + hw] = v[c];
  }
}

// d_logits[c] (+)= p_c * (d_seg_c - sum_k d_seg_k p_k)   (p recomputed from the stored logits)
template <typename T>
__global__ void softmax_bwd_kernel(const T* logits, int ld, const float* d_seg, T* d_logits, int d_ld,
                                   int B, int ncls, long long HW, int do_softmax, int accumulate) {
  pdl_wait(); pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    float p[kMaxClasses], g[kMaxClasses];
    for (int c = 0; c < ncls; ++c) g[c] = d_seg[(n * ncls + c) * HW + hw];
    if (do_softmax) {
      float mx = -INFINITY;
      for (int c = 0; c < ncls; ++c) { p[c] = ld1(logits + i * ld + c); mx = fmaxf(mx, p[c]); }
      float s = 0.f;
      for (int c = 0; c < ncls; ++c) { p[c] = expf(p[c] - mx); s += p[c]; }
      const float inv = 1.f / s;
      float dot = 0.f;
      for (int c = 0; c < ncls; ++c) { p[c] *= inv; dot += p[c] * g[c]; }
      for (int c = 0; c < ncls; ++c) g[c] = p[c] * (g[c] - dot);
    }
    for (int c = 0; c < ncls; ++c) {
      T* dst = d_logits + i * d_ld + c;
      st1(dst, accumulate ? ld1(dst) + g[c] : g[c]);
    }
  }
}

}  // namespace fu
