// Kernel-level test hook (include/fluoro_unet.h: fu_test_conv).  Test infrastructure only.
namespace {

template <typename T>
int test_conv_simt(fu_engine* e, int mode, int B, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                   int relu, const void* x, const float* w, const float* bias, void* y_or_dx, const void* dy,
                   float* dw, double* stats) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int kk = k * k;
  int rc = FU_OK;
  float* wp = nullptr;
  if (mode == 0) {
    const int npad = pad_to(Cout, 4);
    CUDA_TRY(e, cudaMalloc(&wp, (size_t)kk * Cin * npad * sizeof(float)));
    rc = pack_one(e, w, wp, kk, Cin, Cout, npad, Cout, 0, 1, kk, 0, (long long)Cin * kk);
    if (!rc) {
      ConvCall c;
      c.x.p = (char*)x; c.x.C = Cin; c.x.ld = Cin; c.B = B; c.Hi = H; c.Wi = W;
      c.y.p = (char*)y_or_dx; c.y.C = Cout; c.y.ld = Cout; c.Ho = Ho; c.Wo = Wo;
      c.w = wp; c.N = Cout; c.Npad = npad; c.bias = bias; c.bias_mod = Cout;
      c.KH = k; c.stride = stride; c.pad = pad; c.relu = relu; c.stat = stats;
      rc = run_igemm<T>(e, c);
    }
  } else if (mode == 1) {
    if (stride != 1) return e->fail(FU_ERR_ARG, "fu_test_conv: dgrad hook covers stride 1 only");
    const int npad = pad_to(Cin, 4);
    CUDA_TRY(e, cudaMalloc(&wp, (size_t)kk * Cout * npad * sizeof(float)));
    rc = pack_one(e, w, wp, kk, Cout, Cin, npad, Cin, 1, 1, (long long)Cin * kk, 0, kk);
    if (!rc) {
      ConvCall c;
      c.x.p = (char*)dy; c.x.C = Cout; c.x.ld = Cout; c.B = B; c.Hi = Ho; c.Wi = Wo;
      c.y.p = (char*)y_or_dx; c.y.C = Cin; c.y.ld = Cin; c.Ho = H; c.Wo = W;
      c.w = wp; c.N = Cin; c.Npad = npad; c.KH = k; c.stride = 1; c.pad = k - 1 - pad;
      rc = run_igemm<T>(e, c);
    }
  } else {
    CUDA_TRY(e, cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * kk * sizeof(float), e->stream));
    WgradCall c;
    c.big.p = (char*)x; c.big.C = Cin; c.big.ld = Cin; c.Hb = H; c.Wb = W;
    c.small.p = (char*)dy; c.small.C = Cout; c.small.ld = Cout; c.Hs = Ho; c.Ws = Wo; c.B = B;
    c.KH = k; c.stride = stride; c.pad = pad;
    c.dw = dw; c.s_tap = 1; c.s_big = kk; c.s_small = (long long)Cin * kk;
    rc = run_wgrad<T>(e, c);
  }
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (wp) cudaFree(wp);
  if (!rc && ce != cudaSuccess) return e->fail(FU_ERR_CUDA, "fu_test_conv: %s", cudaGetErrorString(ce));
  return rc;
}

}  // namespace

extern "C" int fu_test_conv(int precision, int impl, int mode, int B, int H, int W, int Cin, int Cout, int ksize,
                            int stride, int pad, int relu, const void* x, const float* w, const float* bias,
                            void* y_or_dx, const void* dy, float* dw, double* stats, void* stream) {
  fu_engine e;
  memset(&e.cfg, 0, sizeof(e.cfg));
  memset(&e.cnt, 0, sizeof(e.cnt));
  e.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaGetDevice(&e.device);
  cudaDeviceGetAttribute(&e.num_sms, cudaDevAttrMultiProcessorCount, e.device);
  int rc;
  if (impl == 0) {
    if (precision == FU_PRECISION_BF16)
      rc = test_conv_simt<bf16>(&e, mode, B, H, W, Cin, Cout, ksize, stride, pad, relu, x, w, bias, y_or_dx, dy, dw, stats);
    else
      rc = test_conv_simt<float>(&e, mode, B, H, W, Cin, Cout, ksize, stride, pad, relu, x, w, bias, y_or_dx, dy, dw, stats);
  } else {
    // impl 1: bf16 tensors; impl 2: fp32 tensors through the split-bf16 x3 parity path (FU_PRECISION_FP32_TC)
    rc = tc_test_conv(mode, B, H, W, Cin, Cout, ksize, stride, pad, relu, x, w, bias, y_or_dx, dy, dw, stats,
                      e.stream, &e.cnt, impl == 2 ? 1 : 0);
    if (rc) e.err = tc_last_error();
  }
  if (rc) g_create_error = e.err;
  return rc;
}
