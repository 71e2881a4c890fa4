// C ABI of the sample-preparation and inference post-processing kernels (include/fluoro_unet.h, "callers either
// side of the path"; kernels in kernels_io.cuh).  Stateless: no engine handle; errors go to fu_last_error(NULL).
namespace {

void launch_ens_combine(const EnsArgs& a, dim3 grid, cudaStream_t st) {
  switch (a.n_nets) {
    case 1: ens_combine_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: ens_combine_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: ens_combine_kernel<3><<<grid, 256, 0, st>>>(a); break;
    case 4: ens_combine_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 5: ens_combine_kernel<5><<<grid, 256, 0, st>>>(a); break;
    case 6: ens_combine_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 8: ens_combine_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: ens_combine_kernel<0><<<grid, 256, 0, st>>>(a); break;
  }
}

int io_fail(int rc, const std::string& msg) { g_create_error = msg; return rc; }

int io_launch_check(const char* who) {
  cudaError_t ce = cudaPeekAtLastError();
  if (ce != cudaSuccess) return io_fail(FU_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(ce));
  return FU_OK;
}

// the reference evaluates these scalars in fp32 (0-dim float tensors): sigma*sigma*-2 and 2*pi*sigma*sigma
void gauss_consts(float sigma, float* neg2ss, float* norm) {
  volatile float ss = sigma * sigma;
  *neg2ss = ss * -2.0f;
  volatile float t = (float)(2.0 * 3.14159265358979323846) * sigma;
  *norm = t * sigma;
}

}  // namespace

extern "C" {

int fu_prep_tiles(const float* tiles, int B, int h, int w, int pad, int normalize, double* sums, float* out,
                  void* stream) {
  if (!tiles || !out || B < 1 || h < 1 || w < 1 || pad < 0) return io_fail(FU_ERR_ARG, "fu_prep_tiles: bad argument");
  if (pad >= h || pad >= w) return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_prep_tiles: reflect padding needs pad < tile size");
  if (normalize && !sums) return io_fail(FU_ERR_ARG, "fu_prep_tiles: normalisation needs the 2*B-double workspace");
  PrepArgs a;
  a.src = tiles; a.out = out; a.sums = sums; a.B = B; a.h = h; a.w = w; a.pad = pad;
  a.Hp = h + 2 * pad; a.Wp = w + 2 * pad; a.normalize = normalize ? 1 : 0;
  const long long n = (long long)a.Hp * a.Wp;
  if (n < 2 || n > 0x7ff00000LL) return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_prep_tiles: tile size");
  if (B > 65535) return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_prep_tiles: at most 65535 tiles per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a.normalize) {
    if (cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(double), st) != cudaSuccess)
      return io_fail(FU_ERR_CUDA, "fu_prep_tiles: memset failed");
    // enough blocks to fill the machine (~16 per SM over all tiles), each with at least 8 pixels per thread
    const long long want = std::max<long long>(1, (16LL * 148 + B - 1) / B);
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(std::min<long long>((n + 256 * 8 - 1) / (256 * 8), want), 128));
    prep_stats_kernel<<<dim3(gx, (unsigned)B), 256, 0, st>>>(a);
  }
  a.vec = ((a.Wp & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  const long long groups = (long long)a.Hp * ((a.Wp + 3) / 4);
  prep_apply_kernel<<<dim3((unsigned)((groups + 255) / 256), (unsigned)B), 256, 0, st>>>(a);
  return io_launch_check("fu_prep_tiles");
}

int fu_heatmap_targets(const float* lands, int B, int num_lands, int H, int W, float sigma, float* out, void* stream) {
  if (!lands || !out || B < 1 || num_lands < 1 || H < 1 || W < 1 || !(sigma > 0.f))
    return io_fail(FU_ERR_ARG, "fu_heatmap_targets: bad argument");
  if ((long long)B * num_lands > 65535 || (long long)H * W > 0x7fffffffLL)
    return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_heatmap_targets: B*num_lands <= 65535");
  HeatArgs a;
  a.lands = lands; a.out = out; a.B = B; a.L = num_lands; a.H = H; a.W = W;
  gauss_consts(sigma, &a.neg2ss, &a.norm);
  a.far_r2 = -a.neg2ss * 105.0f;
  a.R = (int)ceilf(sqrtf(a.far_r2)) + 1;
  const long long n = (long long)H * W;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(out, 0, (size_t)B * num_lands * n * sizeof(float), st) != cudaSuccess)
    return io_fail(FU_ERR_CUDA, "fu_heatmap_targets: memset failed");
  const long long box = (long long)(2 * a.R + 2) * (2 * a.R + 2);
  // a few fat blocks per plane: the per-block prologue (landmark load, box origin) is amortised over ~6 pixels a thread
  const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((std::min(box, n) + 256 * 6 - 1) / (256 * 6), 64));
  heatmap_targets_kernel<<<dim3(gx, (unsigned)(B * num_lands)), 256, 0, st>>>(a);
  return io_launch_check("fu_heatmap_targets");
}

int64_t fu_ensemble_workspace_words(int n_nets, int B) { return 2 * (int64_t)n_nets * B; }

int fu_ensemble_combine(const float* const* seg, const float* const* heat, int n_nets, int B, int n_classes,
                        int num_lands, int H, int W, int r0, int c0, int h, int w, uint32_t* workspace,
                        uint8_t* labels, float* avg_heat, void* stream) {
  if (!seg || !labels || n_nets < 1 || B < 1 || n_classes < 1 || num_lands < 0 || h < 1 || w < 1)
    return io_fail(FU_ERR_ARG, "fu_ensemble_combine: bad argument");
  if (n_nets > kMaxNets) return io_fail(FU_ERR_ARG, "fu_ensemble_combine: at most 16 networks per call");
  if (n_classes > 256) return io_fail(FU_ERR_ARG, "fu_ensemble_combine: labels are u1 (util.py:301), at most 256 classes");
  if (r0 < 0 || c0 < 0 || r0 + h > H || c0 + w > W) return io_fail(FU_ERR_ARG, "fu_ensemble_combine: window outside the output");
  if ((num_lands > 0) != (heat != nullptr) || (num_lands > 0) != (avg_heat != nullptr) || (num_lands > 0 && !workspace))
    return io_fail(FU_ERR_ARG, "fu_ensemble_combine: heat / avg_heat / workspace / num_lands disagree");
  EnsArgs a;
  memset(&a, 0, sizeof(a));
  for (int n = 0; n < n_nets; ++n) {
    if (!seg[n] || (heat && !heat[n])) return io_fail(FU_ERR_ARG, "fu_ensemble_combine: null network output");
    a.seg[n] = seg[n];
    a.heat[n] = heat ? heat[n] : nullptr;
  }
  a.n_nets = n_nets; a.B = B; a.C = n_classes; a.L = num_lands; a.H = H; a.W = W;
  a.r0 = r0; a.c0 = c0; a.h = h; a.w = w; a.labels = labels; a.avg_heat = avg_heat;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long hw = (long long)h * w;
  const unsigned gx_c = (unsigned)((hw + 256 * kEnsPix - 1) / (256 * kEnsPix));
  if (num_lands == 0) {
    for (int b0 = 0; b0 < B; b0 += 65535) {
      a.b0 = b0; a.nb = std::min(B - b0, 65535);
      launch_ens_combine(a, dim3(gx_c, (unsigned)a.nb, 1), st);
    }
    return io_launch_check("fu_ensemble_combine");
  }
  const int nbt = n_nets * B;
  a.mn = workspace; a.mx = workspace + nbt;
  if (cudaMemsetAsync(a.mn, 0xff, (size_t)nbt * 4, st) != cudaSuccess || cudaMemsetAsync(a.mx, 0, (size_t)nbt * 4, st) != cudaSuccess)
    return io_fail(FU_ERR_CUDA, "fu_ensemble_combine: memset failed");
  // images per chunk: all networks' heat-maps of a chunk stay within ~48 MB so the second pass hits L2
  const long long per_img = (long long)n_nets * num_lands * H * W * 4;
  const int chunk = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(B, 65535 / n_nets),
                                                                    (48LL << 20) / std::max<long long>(per_img, 1)));
  const long long row_groups = ((long long)num_lands * h + 7) / 8;   // a block pass covers 8 rows
  const unsigned gx_m = (unsigned)std::max<long long>(1, std::min<long long>((row_groups + 3) / 4, 64));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    a.b0 = b0; a.nb = std::min(B - b0, chunk);
    ens_minmax_kernel<<<dim3(gx_m, (unsigned)(n_nets * a.nb)), 256, 0, st>>>(a);
    launch_ens_combine(a, dim3(gx_c, (unsigned)a.nb, (unsigned)(num_lands + 1)), st);
  }
  return io_launch_check("fu_ensemble_combine");
}

int fu_extract_landmarks(const float* heats, const uint8_t* segs, const int32_t* seg_labels, int P, int num_lands,
                         int h, int w, int tmpl_dim, float sigma, float min_ncc, int32_t* out_rc, float* out_ncc,
                         void* stream) {
  if (!heats || !out_rc || P < 1 || num_lands < 1 || h < 1 || w < 1 || !(sigma > 0.f))
    return io_fail(FU_ERR_ARG, "fu_extract_landmarks: bad argument");
  if (num_lands > kMaxLands) return io_fail(FU_ERR_ARG, "fu_extract_landmarks: at most 64 landmarks");
  if (tmpl_dim < 3 || (tmpl_dim & 1) == 0) return io_fail(FU_ERR_ARG, "fu_extract_landmarks: the template size must be odd and >= 3");
  if (tmpl_dim / 2 >= h || tmpl_dim / 2 >= w)
    return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_extract_landmarks: reflect padding needs tmpl_dim/2 < heat-map size");
  if (segs && !seg_labels) return io_fail(FU_ERR_ARG, "fu_extract_landmarks: segs without per-landmark labels");
  if ((long long)h * w > 0x7fffffffLL) return io_fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_extract_landmarks: heat-map size");
  LandArgs a;
  memset(&a, 0, sizeof(a));
  a.heats = heats; a.segs = segs; a.out = out_rc; a.ncc_out = out_ncc;
  for (int l = 0; l < num_lands; ++l) a.label[l] = seg_labels ? seg_labels[l] : -1;
  a.P = P; a.L = num_lands; a.h = h; a.w = w; a.D = tmpl_dim; a.min_ncc = min_ncc;
  gauss_consts(sigma, &a.neg2ss, &a.norm);
  a.vec = (((long long)h * w & 3) == 0 && (reinterpret_cast<uintptr_t>(heats) & 15) == 0 &&
           (!segs || (reinterpret_cast<uintptr_t>(segs) & 3) == 0)) ? 1 : 0;
  extract_landmarks_kernel<<<(unsigned)((long long)P * num_lands), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return io_launch_check("fu_extract_landmarks");
}

}  // extern "C"
