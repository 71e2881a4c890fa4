// Fused training loss of the reference: DiceLoss2D / DiceAndHeatMapLoss2D
// (dice.py:14-86) over ncc_2d (ncc.py:12-38), with the centre crop of the network outputs
// (util.py:92-114, called at train.py:414-417) folded into the indexing.
//
// The PyTorch formulation makes ~25 elementwise / reduction passes over the (B,7+14,H,W) outputs and
// their cropped views (0.9 ms of a 8.5 ms step at B=32, 192x192); here the forward is ONE pass that
// reduces every (image, channel) plane to 3 (Dice) or 5 (NCC) fp64 sums, a one-block kernel turns the
// sums into the scalar loss, and the backward is ONE pass that writes the full-size (uncropped)
// gradients of both outputs, zeros outside the crop window, straight from the closed-form derivative.
#pragma once
#include "common.cuh"

namespace fu {

struct LossArgs {
  const float* seg; long long seg_sb, seg_sc; int seg_sr;        // prediction strides: batch, channel, row (elements)
  const float* mask; long long mask_sb, mask_sc; int mask_sr;
  const float* heat; long long heat_sb, heat_sc; int heat_sr;    // heat == nullptr: Dice only
  const float* heat_t; long long heat_t_sb, heat_t_sc; int heat_t_sr;
  int B, NC, NL;
  int Ht, Wt;                   // window (= target) size; prediction pointers already address the window origin
  int skip_bg;
  float dice_wgt, heat_wgt;
  double* sums;                 // [B][NC*3 + NL*5]
  int rows;                     // image rows per block (loss_rows_per_block): warp w takes rows w, w+8, ...; lanes stride the columns
};

// Rows per block: whole planes when there are enough planes to fill the machine (B = 32 @192x192: 672 planes), else chunks
// sized for ~4 blocks per SM.  16-row blocks (8064 of them at B = 32) spent most of their time in the five block-level
// fp64 reductions and atomics each block ends with: 61 + 95 us for two passes over 180 + 250 MB.
// (Measured later and not kept: chunks sized for ~6 waves -- sums pass unchanged at 60 us, gradient pass 46 -> 50 us.  What the
// sums pass did respond to was wider loads: loss_sums_kernel<VEC2>.)
inline int loss_rows_per_block(int rows_total, long long planes, int sms = 148) {
  long long r = ((long long)rows_total * planes + (long long)sms * 4 - 1) / ((long long)sms * 4);
  r = (r + 7) / 8 * 8;
  if (r < 8) r = 8;
  if (r > rows_total) r = rows_total;
  return (int)r;
}

__device__ __forceinline__ double block_sum_double(double v, double* sm) {
  // 256 threads: warp shuffle, then one value per warp through shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 8) t = sm[threadIdx.x];
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;    // valid in thread 0
}

// N sums at once: one shuffle tree per value, one shared-memory exchange, threads 0..N-1 end up with the totals
template <int N>
__device__ __forceinline__ double block_sum_n(double (&v)[N], double (*sm)[8]) {
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) sm[k][w] = v[k];
  }
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[threadIdx.x][i];
  }
  return t;    // thread k < N: total of value k
}

// grid: (row chunks of the window, B*(NC+NL) planes)
// VEC2: two columns per lane and load (8-byte loads; the window origin of the predictions is 8- but not 16-byte aligned for the
// reference's crops: (192 - 180) / 2 = 6 columns).  The pass is bound by bytes in flight -- 3.3 TB/s with 4-byte loads, measured
// alone under ncu -- not by arithmetic.  Preconditions (launcher): Wt, all strides even, all plane / window bases 8-byte aligned.
template <bool VEC2>
__global__ void __launch_bounds__(256) loss_sums_kernel(const LossArgs p) {
  __shared__ double sm[5][8];
  const int plane = blockIdx.y;
  const int b = plane / (p.NC + p.NL), c = plane - b * (p.NC + p.NL);
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int r_begin = blockIdx.x * p.rows, r_end = min(r_begin + p.rows, p.Ht);
  const int per = p.NC * 3 + p.NL * 5;
  if (c < p.NC) {
    const float* x = p.seg + b * p.seg_sb + c * p.seg_sc;
    const float* t = p.mask + b * p.mask_sb + c * p.mask_sc;
    double s_tp = 0.0, s_tt = 0.0, s_pp = 0.0;
    // four rows per warp at a time: eight independent loads per lane and column step (one row at a time left the pass
    // latency bound: 58 us for 180 MB)
    for (int r = r_begin + wrp; r < r_end; r += 32) {
      const float* xr[4]; const float* tr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ru = min(r + 8 * u, r_end - 1);        // (clamped rows are loaded and not counted)
        xr[u] = x + (long long)ru * p.seg_sr; tr[u] = t + (long long)ru * p.mask_sr;
      }
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};   // <= 8 columns per lane and row
      if (VEC2) {
        for (int col = lane; col < (p.Wt >> 1); col += 32) {
          float2 xv[4], tv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { xv[u] = reinterpret_cast<const float2*>(xr[u])[col]; tv[u] = reinterpret_cast<const float2*>(tr[u])[col]; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            a0[u] = fmaf(tv[u].x, xv[u].x, fmaf(tv[u].y, xv[u].y, a0[u])); a1[u] = fmaf(tv[u].x, tv[u].x, fmaf(tv[u].y, tv[u].y, a1[u]));
            a2[u] = fmaf(xv[u].x, xv[u].x, fmaf(xv[u].y, xv[u].y, a2[u]));
          }
        }
      } else
      for (int col = lane; col < p.Wt; col += 32) {
        float xv[4], tv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { xv[u] = xr[u][col]; tv[u] = tr[u][col]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0[u] = fmaf(tv[u], xv[u], a0[u]); a1[u] = fmaf(tv[u], tv[u], a1[u]); a2[u] = fmaf(xv[u], xv[u], a2[u]); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + 8 * u < r_end) { s_tp += (double)a0[u]; s_tt += (double)a1[u]; s_pp += (double)a2[u]; }
    }
    double v3[3] = {s_tp, s_tt, s_pp};
    const double tot = block_sum_n<3>(v3, sm);
    if (threadIdx.x < 3) atomicAdd(p.sums + (long long)b * per + c * 3 + threadIdx.x, tot);
  } else {
    const int l = c - p.NC;
    const float* x = p.heat + b * p.heat_sb + l * p.heat_sc;
    const float* y = p.heat_t + b * p.heat_t_sb + l * p.heat_t_sc;
    double sx = 0.0, sxx = 0.0, sy = 0.0, syy = 0.0, sxy = 0.0;
    for (int r = r_begin + wrp; r < r_end; r += 32) {
      const float* xr[4]; const float* yr[4];
      bool on[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        on[u] = r + 8 * u < r_end;
        const int ru = min(r + 8 * u, r_end - 1);
        xr[u] = x + (long long)ru * p.heat_sr; yr[u] = y + (long long)ru * p.heat_t_sr;
      }
      // fp32 over the <= 8 columns a lane sees of a row, fp64 from there on (the variances are differences of sums: a
      // partial of <= 8 terms carries <= 5e-7 relative error and the ~6000 partials of a plane average it out; element-wise
      // fp64 arithmetic -- 7 fp64-pipe operations per element -- was what bounded this pass, 58 us for 180 MB)
      float m[4][5];
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u][0] = m[u][1] = m[u][2] = m[u][3] = m[u][4] = 0.f;
      if (VEC2) {
        for (int col = lane; col < (p.Wt >> 1); col += 32) {
          float2 xf[4], yf[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { xf[u] = reinterpret_cast<const float2*>(xr[u])[col]; yf[u] = reinterpret_cast<const float2*>(yr[u])[col]; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            m[u][0] += xf[u].x + xf[u].y; m[u][1] = fmaf(xf[u].x, xf[u].x, fmaf(xf[u].y, xf[u].y, m[u][1]));
            m[u][2] += yf[u].x + yf[u].y; m[u][3] = fmaf(yf[u].x, yf[u].x, fmaf(yf[u].y, yf[u].y, m[u][3]));
            m[u][4] = fmaf(xf[u].x, yf[u].x, fmaf(xf[u].y, yf[u].y, m[u][4]));
          }
        }
      } else
      for (int col = lane; col < p.Wt; col += 32) {
        float xf[4], yf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { xf[u] = xr[u][col]; yf[u] = yr[u][col]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          m[u][0] += xf[u]; m[u][1] = fmaf(xf[u], xf[u], m[u][1]); m[u][2] += yf[u]; m[u][3] = fmaf(yf[u], yf[u], m[u][3]);
          m[u][4] = fmaf(xf[u], yf[u], m[u][4]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (on[u]) { sx += (double)m[u][0]; sxx += (double)m[u][1]; sy += (double)m[u][2]; syy += (double)m[u][3]; sxy += (double)m[u][4]; }
    }
    double v5[5] = {sx, sxx, sy, syy, sxy};
    const double tot = block_sum_n<5>(v5, sm);
    if (threadIdx.x < 5) atomicAdd(p.sums + (long long)b * per + p.NC * 3 + l * 5 + threadIdx.x, tot);
  }
}

// per-plane NCC pieces from the raw moments (ncc.py:12-38: sample standard deviations, 1e-8 in the denominator)
struct NccTerms { double mx, my, sdx, sdy, sxy_c, den; };
__device__ __forceinline__ NccTerms ncc_terms(const double* m, double N) {
  NccTerms t;
  t.mx = m[0] / N; t.my = m[2] / N;
  const double vxx = fmax(m[1] - N * t.mx * t.mx, 0.0), vyy = fmax(m[3] - N * t.my * t.my, 0.0);
  t.sdx = sqrt(vxx / (N - 1.0)); t.sdy = sqrt(vyy / (N - 1.0));
  t.sxy_c = m[4] - N * t.mx * t.my;
  t.den = N * t.sdx * t.sdy + 1.0e-8;
  return t;
}

// one block: the scalar loss (dice.py:48-55, 74-86)
__global__ void __launch_bounds__(256) loss_finalize_kernel(const LossArgs p, float* loss_out) {
  __shared__ double sm[8];
  const int per = p.NC * 3 + p.NL * 5;
  const int c_first = p.skip_bg ? 1 : 0;
  const int n_cls = p.NC - c_first;
  const double N = (double)p.Ht * p.Wt;
  double dice = 0.0, ncc = 0.0;
  for (int i = threadIdx.x; i < p.B * p.NC; i += 256) {
    const int b = i / p.NC, c = i - b * p.NC;
    if (c < c_first) continue;
    const double* s = p.sums + (long long)b * per + c * 3;
    dice += (-2.0 * s[0] + 1.0e-4) / (s[1] + s[2] + 1.0e-4);
  }
  for (int i = threadIdx.x; i < p.B * p.NL; i += 256) {
    const int b = i / p.NL, l = i - b * p.NL;
    const NccTerms t = ncc_terms(p.sums + (long long)b * per + p.NC * 3 + l * 5, N);
    ncc += (t.sxy_c / t.den + 1.0) * -0.5;
  }
  dice = block_sum_double(dice, sm);
  ncc = block_sum_double(ncc, sm);
  if (threadIdx.x == 0) {
    double loss = (double)p.dice_wgt * dice / ((double)n_cls * p.B);
    if (p.NL > 0) loss += (double)p.heat_wgt * ncc / ((double)p.B * p.NL);
    *loss_out = (float)loss;
  }
}

struct LossBwdArgs {
  LossArgs a;
  const float* dloss;           // device scalar: upstream gradient of the loss
  float* d_seg; float* d_heat;  // full-size contiguous (B,NC,H,W) / (B,NL,H,W) gradients
  int H, W, r0, c0;             // full output size and window origin
};

// Four columns per thread: the block's rows x (W / 4) column groups are dealt to the threads as one flat index space, the
// gradient leaves as aligned 16-byte stores and the predictions arrive as aligned 16-byte loads (full-plane addresses);
// the targets sit at window coordinates (misaligned by the crop origin) and are read as scalars.  v = f(target, prediction).
// Preconditions (checked by the launcher): W % 4 == 0, prediction row stride % 4 == 0, plane bases 16-byte aligned.
template <typename F>
__device__ __forceinline__ void loss_bwd_rows_vec4(const LossBwdArgs& q, const float* x_full /* plane origin (full tensor) */, int x_sr,
                                                   const float* t, int t_sr, float* g, int R_begin, int R_end, bool plane_on, F f) {
  const LossArgs& p = q.a;
  const int W4 = q.W >> 2;
  const int items = (R_end - R_begin) * W4;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int rr = i / W4, c4 = i - rr * W4;
    const int R = R_begin + rr, C = c4 * 4;
    const int r = R - q.r0, col = C - q.c0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (plane_on && r >= 0 && r < p.Ht && col + 3 >= 0 && col < p.Wt) {
      const float4 xv = *reinterpret_cast<const float4*>(x_full + (long long)R * x_sr + C);
      const float* tr = t + (long long)r * t_sr;
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = col + k;
        o[k] = (cc >= 0 && cc < p.Wt) ? f(tr[cc], xs[k]) : 0.f;
      }
      v = make_float4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<float4*>(g + (long long)R * q.W + C) = v;
  }
}

// grid: (row chunks of the FULL plane, B*(NC+NL) planes); zeros outside the window
template <bool VEC4>
__global__ void __launch_bounds__(256) loss_backward_kernel(const LossBwdArgs q) {
  const LossArgs& p = q.a;
  const int plane = blockIdx.y;
  const int b = plane / (p.NC + p.NL), c = plane - b * (p.NC + p.NL);
  const int per = p.NC * 3 + p.NL * 5;
  const long long n_full = (long long)q.H * q.W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int R_begin = blockIdx.x * p.rows, R_end = min(R_begin + p.rows, q.H);
  const float up = *q.dloss;
  if (c < p.NC) {
    float* g = q.d_seg + ((long long)b * p.NC + c) * n_full;
    const int c_first = p.skip_bg ? 1 : 0;
    const double* s = p.sums + (long long)b * per + c * 3;
    const double num = -2.0 * s[0] + 1.0e-4, den = s[1] + s[2] + 1.0e-4;
    // d(num/den)/dp = (-2 t den - 2 p num) / den^2
    const float k = c < c_first ? 0.f : (float)((double)up * p.dice_wgt / ((double)(p.NC - c_first) * p.B) / (den * den));
    const float fden = (float)den, fnum = (float)num;
    const float* x = p.seg + b * p.seg_sb + c * p.seg_sc;
    const float* t = p.mask + b * p.mask_sb + c * p.mask_sc;
    if (VEC4) {
      // (p.seg addresses the window origin: step back to the plane origin)
      loss_bwd_rows_vec4(q, x - ((long long)q.r0 * p.seg_sr + q.c0), p.seg_sr, t, p.mask_sr, g, R_begin, R_end, k != 0.f,
                         [=](float tv, float xv) { return k * (-2.f * tv * fden - 2.f * xv * fnum); });
      return;
    }
    for (int R = R_begin + wrp; R < R_end; R += 8) {
      const int r = R - q.r0;
      const bool row_in = r >= 0 && r < p.Ht && k != 0.f;
      const float* xr = x + (long long)r * p.seg_sr;
      const float* tr = t + (long long)r * p.mask_sr;
      for (int C = lane; C < q.W; C += 32) {
        const int col = C - q.c0;
        float v = 0.f;
        if (row_in && col >= 0 && col < p.Wt) v = k * (-2.f * tr[col] * fden - 2.f * xr[col] * fnum);
        g[(long long)R * q.W + C] = v;
      }
    }
  } else {
    const int l = c - p.NC;
    float* g = q.d_heat + ((long long)b * p.NL + l) * n_full;
    const double N = (double)p.Ht * p.Wt;
    const NccTerms t = ncc_terms(p.sums + (long long)b * per + p.NC * 3 + l * 5, N);
    // ncc = Sxy / D, D = N sdx sdy + 1e-8:  d ncc / dx_i = (y_i - my)/D - Sxy N sdy (x_i - mx) / ((N-1) sdx D^2)
    const double w = (double)up * p.heat_wgt * -0.5 / ((double)p.B * p.NL);
    const float ka = (float)(w / t.den);
    const float kb = t.sdx > 0.0 ? (float)(w * t.sxy_c * N * t.sdy / ((N - 1.0) * t.sdx * t.den * t.den)) : 0.f;
    const float mx = (float)t.mx, my = (float)t.my;
    const float* x = p.heat + b * p.heat_sb + l * p.heat_sc;
    const float* y = p.heat_t + b * p.heat_t_sb + l * p.heat_t_sc;
    if (VEC4) {
      loss_bwd_rows_vec4(q, x - ((long long)q.r0 * p.heat_sr + q.c0), p.heat_sr, y, p.heat_t_sr, g, R_begin, R_end, true,
                         [=](float yv, float xv) { return ka * (yv - my) - kb * (xv - mx); });
      return;
    }
    for (int R = R_begin + wrp; R < R_end; R += 8) {
      const int r = R - q.r0;
      const bool row_in = r >= 0 && r < p.Ht;
      const float* xr = x + (long long)r * p.heat_sr;
      const float* yr = y + (long long)r * p.heat_t_sr;
      for (int C = lane; C < q.W; C += 32) {
        const int col = C - q.c0;
        float v = 0.f;
        if (row_in && col >= 0 && col < p.Wt) v = ka * (yr[col] - my) - kb * (xr[col] - mx);
        g[(long long)R * q.W + C] = v;
      }
    }
  }
}

}  // namespace fu
