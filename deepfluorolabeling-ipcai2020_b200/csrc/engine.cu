// The U-Net forward/backward engine behind include/fluoro_unet.h.
//
// Host side: builds the layer table of the reference network (unet.py:41-159),
// plans an NHWC activation arena for a given (B,H,W), and enqueues the kernels of
// kernels_simt.cuh / kernels_tc.cuh on the caller's stream.  No host
// synchronisation on the step path, no CPU fallback.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/fluoro_unet.h"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_loss.cuh"
#include "kernels_io.cuh"

using namespace fu;

namespace {

thread_local std::string g_create_error;

// the C entry points make the engine's device current for their own launches and put the caller's device back on
// return (PyTorch tracks the current device itself; changing it behind its back breaks its next allocation)
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

struct View {
  char* p = nullptr;  // address of channel 0 of this view
  int C = 0;          // channels in the view
  int ld = 0;         // pixel stride, in elements
  char* sp = nullptr; // parity_tc mode: channel 0 of the hi half of this view's split-bf16 twin (kernels_tc.cuh:
                      // split_f32_kernel); twin pixel stride 2*ld bf16 elements, lo half ld elements after the hi half
  // planar buffer (plane > 0): channels [k*planeC, (k+1)*planeC) are a contiguous (pixels, planeC) tensor at p + k*plane.
  // Used for the gradient of the 32-channel skip concat in bf16 storage: its two halves are read separately (up-conv
  // backward / encoder block backward), and a 64-byte half of a 128-byte pixel costs a full line of DRAM traffic per
  // read (ncu: 226 MB instead of 151 MB for the BN backward of the first encoder block).
  size_t plane = 0;
  int planeC = 0;
};

struct TensorSlot {
  fu_tensor_info info;
  void* data = nullptr;
  int64_t numel = 0;
  bool late = false;   // gradient produced after the mid-backward point (shallow encoder levels): second all-reduce bucket
};

struct ConvW {
  int Cin = 0, Cout = 0, k = 0;
  bool transposed = false;
  int w_idx = -1, b_idx = -1;
  float* wp_fwd = nullptr;    // CUDA-core packing for the forward GEMM
  float* wp_dgrad = nullptr;  // CUDA-core packing for the data-gradient GEMM
  int npad_fwd = 0, npad_dgrad = 0, n_fwd = 0, n_dgrad = 0;
  bool small_cin = false;     // first-layer kernels (Cin <= 2, Cout % 8 == 0)
  double* bsum = nullptr;     // [Cout] bias-gradient accumulator
  TcConv tc;                  // tensor-core packings / descriptors (throughput mode)
};

struct BNL {
  int C = 0;
  int i_gamma = -1, i_beta = -1, i_rm = -1, i_rv = -1, i_nbt = -1;
  double *stat = nullptr, *bstat = nullptr;
  float *mean = nullptr, *invstd = nullptr, *a = nullptr, *b = nullptr, *ga = nullptr, *m1 = nullptr,
        *m2 = nullptr;
};

struct Block {
  int Cin = 0, C = 0;
  bool has_res = false;
  ConvW res;
  std::vector<ConvW> convs;
  std::vector<BNL> bns;
  double* dstat = nullptr;      // [2*Cin]: per-channel sums of the block's input gradient (from the last dgrad's epilogue)
  // plan-dependent
  std::vector<View> r, z, dy, dz;
};

struct Plan {
  int B = 0, H = 0, W = 0;
  bool valid = false;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  char* twin = nullptr;         // parity_tc mode: second arena of the same size holding the split-bf16 twins
  View xin;
  std::vector<View> cat, d_cat, down, d_down, decout, d_decout;
  View bott, d_bott, hcat, d_hcat, dheat;
  std::vector<View> hmid, d_hmid;
};

}  // namespace

struct fu_engine {
  fu_config cfg;
  int device = 0;
  int esz = 4;
  std::string err;
  std::vector<TensorSlot> tensors;
  int64_t grad_numel = 0;
  bool bound = false;
  std::vector<int> chans;
  std::vector<Block> enc, dec;
  std::vector<ConvW> downc, upc;  // downsample convs (depth-1 used), up convs
  ConvW seg;
  std::vector<ConvW> lands;
  int Cf = 0, Cpad = 0;
  // persistent device memory
  char* wmem = nullptr;
  size_t wmem_bytes = 0;
  double* dscr_fwd = nullptr; size_t dscr_fwd_bytes = 0;
  double* dscr_bwd = nullptr; size_t dscr_bwd_bytes = 0;
  char* wgrad_scr = nullptr; size_t wgrad_scr_bytes = 0;   // tensor-core weight-gradient accumulators
  bool wgrad_scr_clean = false;  // all zeros: set by a backward that ran to its end (tc_unpack_batched_kernel reads and clears)
  // Slices of the flat gradient buffer that the unpack launches OVERWRITE (97 % of it) need no zeroing: after the first
  // backward of a plan only the gaps between them (biases, BatchNorm, first layer, heads: accumulated with atomics) are
  // zeroed, by one kernel over a table of ranges.
  std::vector<std::pair<long long, long long>> cover_now;   // (offset, floats) of this backward's unpack jobs
  bool gaps_valid = false; unsigned long long cover_hash = 0;
  long long* gap_tbl = nullptr; int n_gaps = 0; int gap_cap = 0;
  std::vector<long long> gap_host;          // host copy of the table (source of the stream-ordered upload)
  float *ones = nullptr, *zeros = nullptr;
  float* heads_gacc = nullptr;   // [NL*(CF+NC) + NC*CF] accumulators of the fused heads backward
  // the training loss inside the heads kernels (fu_forward_loss / fu_backward_loss): set for the duration of such a call
  const HeadsLoss* lossf = nullptr;
  const float* lossf_dloss = nullptr;
  bool lossf_noseg = false;      // fu_forward_loss was given no seg output: the class probabilities stay in the head kernel
  LossArgs lossf_args;           // sizes / weights of the loss (finalisation, coefficient kernel)
  float* loss_coef = nullptr;    // [B][NC*2 + NL*3] per-plane gradient coefficients (own allocation, grows with B)
  size_t loss_coef_floats = 0;
  unsigned* coop_bar = nullptr;  // {arrival count, generation} of the grid barrier in bn_act_bwd_coop_kernel (self-resetting)
  int coop_blocks_per_sm = -1;   // resident blocks per SM of that kernel (occupancy query, once); 0 = do not use it
  // batched weight pack / weight-gradient unpack (kernels_tc.cuh): device job tables and what they hold
  static constexpr int kJobCap = 512;
  TcPackJob* pack_tbl = nullptr; TcUnpackJob* unpack_tbl = nullptr;
  TcPackJob* pack_pin = nullptr; TcUnpackJob* unpack_pin = nullptr;      // page-locked staging copies
  std::vector<TcPackJob> pack_uploaded, pack_uploaded2; std::vector<TcUnpackJob> unpack_uploaded, unpack_uploaded2;
  TcBatch batch;
  TcBatch batch_deep;            // pack jobs of the layers first used at level >= split_level(): packed on the side stream
  int split_level() const { return cfg.depth / 2 > 1 ? cfg.depth / 2 : 1; }
  int64_t packed_version = -1;
  bool packed_once = false;
  Plan plan;
  bool saved = false;
  int saved_training = 0;
  fu_counters cnt;
  cudaStream_t stream = nullptr;
  // Second stream of the backward pass: weight gradients only produce dW, nothing downstream in the step reads them
  // before the final unpack, so they run beside the dgrad -> BN-backward chain (HBM-bound, small shared memory) instead
  // of in line with it.  Fork = event on the caller's stream, join = event before the unpack; inside a CUDA-graph
  // capture this becomes a parallel branch of the graph.  FU_STREAMS=1 keeps everything on the caller's stream.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_used = false;
  int use_side = 1;
  int num_sms = 148;
  // Gradient buckets (data parallelism, fu_set_bucket_callback): flat[0, early_numel) holds every gradient that is
  // final once the backward pass has left encoder level split_level() -- heads, the whole decoder, the deep encoder
  // levels: 97 % of the parameters of the paper network -- so their all-reduce can run beside the shallow encoder
  // levels' backward (the most expensive 40 % of the pass); the rest forms the tail bucket.
  int64_t early_numel = 0;
  fu_bucket_callback bucket_cb = nullptr;
  void* bucket_user = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_mid_main = nullptr, ev_mid_side = nullptr;
  bool creating_late = false;
  bool split = false;            // parity_tc: fp32 storage, tensor-core layers read split-bf16 twins (three MMA passes)
  // twins that are up to date: (view address, channels).  Forward activations stay valid until the next forward;
  // gradient tensors are produced once per backward.
  std::vector<std::pair<const void*, int>> fresh_fwd, fresh_bwd;
  bool in_backward = false;
  // optional per-launch CUDA-event profiling (fu_profile_enable)
  struct DeferredSum { const double* src; float* dst; int n; int copies; };
  std::vector<DeferredSum> deferred_sums;
  bool prof = false;
  struct ProfRec { std::string tag; const char* kern; cudaEvent_t a, b; double flops, bytes; };
  std::vector<ProfRec> prof_recs;
  std::string tag = "other";
  double tag_flops = 0, tag_bytes = 0;
  void set_tag(double flops, double bytes, const char* fmt, ...) {
    if (!prof) return;
    char buf[160];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tag = buf; tag_flops = flops; tag_bytes = bytes;
  }
  void prof_begin(const char* kern) {
    ProfRec r; r.tag = tag; r.kern = kern; r.flops = tag_flops; r.bytes = tag_bytes;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream);
    prof_recs.push_back(r);
  }
  void prof_end() { cudaEventRecord(prof_recs.back().b, stream); }

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CUDA_TRY(e, expr)                                                                  \
  do {                                                                                     \
    cudaError_t _ce = (expr);                                                              \
    if (_ce != cudaSuccess)                                                                \
      return (e)->fail(FU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_ce), \
                       __FILE__, __LINE__);                                                \
  } while (0)

#define LAUNCH(e, kern, grid, block, ...)                                                   \
  do {                                                                                      \
    auto _kfn = kern;                                                                       \
    if ((e)->prof) (e)->prof_begin(#kern);                                                  \
    fu_launch(_kfn, dim3(grid), dim3(block), 0, (e)->stream, fu_pdl_enabled(), __VA_ARGS__); \
    if ((e)->prof) (e)->prof_end();                                                         \
    (e)->cnt.kernel_launches++;                                                             \
    cudaError_t _ce = cudaPeekAtLastError();                                                \
    if (_ce != cudaSuccess)                                                                 \
      return (e)->fail(FU_ERR_CUDA, "launch %s failed: %s (%s:%d)", #kern,                  \
                       cudaGetErrorString(_ce), __FILE__, __LINE__);                        \
  } while (0)

// same, with dynamic shared memory (> 48 KB is opted into once per kernel)
#define LAUNCH_SMEM(e, kern, grid, block, smem, ...)                                        \
  do {                                                                                      \
    auto _kfn = kern;                                                                       \
    static bool _attr[64] = {};                                                             \
    if (tc_attr_needed(_attr)) {     /* per device; opted in to the largest size any launch may ask for */ \
      cudaFuncAttributes _fa;                                                               \
      int _stat = cudaFuncGetAttributes(&_fa, _kfn) == cudaSuccess ? (int)_fa.sharedSizeBytes : 0; \
      cudaFuncSetAttribute(_kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - _stat); \
    }                                                                                       \
    if ((e)->prof) (e)->prof_begin(#kern);                                                  \
    fu_launch(_kfn, dim3(grid), dim3(block), (smem), (e)->stream, fu_pdl_enabled(), __VA_ARGS__); \
    if ((e)->prof) (e)->prof_end();                                                         \
    (e)->cnt.kernel_launches++;                                                             \
    cudaError_t _ce = cudaPeekAtLastError();                                                \
    if (_ce != cudaSuccess)                                                                 \
      return (e)->fail(FU_ERR_CUDA, "launch %s failed: %s (%s:%d)", #kern,                  \
                       cudaGetErrorString(_ce), __FILE__, __LINE__);                        \
  } while (0)

namespace {

// launches issued while a SideScope is alive go to the engine's side stream, ordered after everything enqueued on the
// caller's stream so far (no-op when profiling, or with FU_STREAMS=1)
struct SideScope {
  fu_engine* e; cudaStream_t saved; bool active;
  explicit SideScope(fu_engine* e_) : e(e_), saved(e_->stream), active(e_->use_side && !e_->prof && e_->side != nullptr) {
    if (active) {
      cudaEventRecord(e->ev_fork, saved);
      cudaStreamWaitEvent(e->side, e->ev_fork, 0);
      e->stream = e->side;
      e->side_used = true;
    }
  }
  ~SideScope() { if (active) e->stream = saved; }
};
inline void side_join(fu_engine* e) {
  if (!e->side_used) return;
  cudaEventRecord(e->ev_join, e->side);
  cudaStreamWaitEvent(e->stream, e->ev_join, 0);
  e->side_used = false;
}

inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }
inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline View slice(const View& v, int c0, int C, int esz) {
  View o;
  if (v.plane) {          // whole planes only
    o.p = v.p + (size_t)(c0 / v.planeC) * v.plane;
    o.C = C;
    o.ld = v.planeC;
    return o;
  }
  o.p = v.p + (size_t)c0 * esz;
  o.C = C;
  o.ld = v.ld;
  o.sp = v.sp ? v.sp + (size_t)c0 * 2 : nullptr;
  return o;
}
inline unsigned grid1d(long long total, int block, int sms) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)sms * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ---------------------------------------------------------------------------
// schema (state_dict order of unet.py; must match oracle/unet_oracle.py:param_schema)
// ---------------------------------------------------------------------------
int add_tensor(fu_engine* e, const std::string& name, std::vector<int64_t> shape, int kind, int dtype,
               bool has_grad) {
  TensorSlot s;
  memset(&s.info, 0, sizeof(s.info));
  snprintf(s.info.name, sizeof(s.info.name), "%s", name.c_str());
  s.info.ndim = (int)shape.size();
  int64_t n = 1;
  for (size_t i = 0; i < shape.size(); ++i) {
    s.info.shape[i] = shape[i];
    n *= shape[i];
  }
  s.numel = n;
  s.info.kind = kind;
  s.info.dtype = dtype;
  s.info.grad_offset = has_grad ? 0 : -1;   // offsets are assigned by assign_grad_offsets (early bucket first)
  s.late = e->creating_late;
  e->tensors.push_back(s);
  return (int)e->tensors.size() - 1;
}

ConvW make_conv(fu_engine* e, const std::string& prefix, int cin, int cout, int k, bool bias, bool transposed,
                bool reachable) {
  ConvW c;
  c.Cin = cin; c.Cout = cout; c.k = k; c.transposed = transposed;
  if (transposed)
    c.w_idx = add_tensor(e, prefix + ".weight", {cin, cout, k, k}, FU_KIND_PARAM, FU_DTYPE_F32, reachable);
  else
    c.w_idx = add_tensor(e, prefix + ".weight", {cout, cin, k, k}, FU_KIND_PARAM, FU_DTYPE_F32, reachable);
  if (bias) c.b_idx = add_tensor(e, prefix + ".bias", {cout}, FU_KIND_PARAM, FU_DTYPE_F32, reachable);
  return c;
}

void make_block(fu_engine* e, Block& b, const std::string& prefix, int cin, int cout) {
  const fu_config& c = e->cfg;
  b.Cin = cin; b.C = cout; b.has_res = c.do_res != 0;
  if (b.has_res) b.res = make_conv(e, prefix + ".res_conv1x1", cin, cout, 1, true, false, true);
  int idx = 0, ci = cin;
  for (int d = 0; d < c.block_depth; ++d) {
    b.convs.push_back(make_conv(e, prefix + ".block." + std::to_string(idx), ci, cout, 3, true, false, true));
    idx += 2;
    if (c.batch_norm) {
      BNL bn;
      bn.C = cout;
      const std::string p = prefix + ".block." + std::to_string(idx);
      bn.i_gamma = add_tensor(e, p + ".weight", {cout}, FU_KIND_PARAM, FU_DTYPE_F32, true);
      bn.i_beta = add_tensor(e, p + ".bias", {cout}, FU_KIND_PARAM, FU_DTYPE_F32, true);
      bn.i_rm = add_tensor(e, p + ".running_mean", {cout}, FU_KIND_BUFFER, FU_DTYPE_F32, false);
      bn.i_rv = add_tensor(e, p + ".running_var", {cout}, FU_KIND_BUFFER, FU_DTYPE_F32, false);
      bn.i_nbt = add_tensor(e, p + ".num_batches_tracked", {}, FU_KIND_BUFFER, FU_DTYPE_I64, false);
      b.bns.push_back(bn);
      idx += 1;
    }
    ci = cout;
  }
}

int build_schema(fu_engine* e) {
  const fu_config& c = e->cfg;
  e->chans.clear();
  for (int i = 0; i < c.depth; ++i) e->chans.push_back(1 << (c.wf + i));
  // "late" tensors: their gradients are written after backward_t's mid point (encoder levels below split_level() and
  // the downsample convs feeding those levels' gradients back: downc[i] is processed at the end of level i + 1)
  const int Ls = e->split_level();
  if (!c.max_pool)
    for (int i = 0; i < c.depth; ++i) {
      e->creating_late = i + 1 < Ls;
      e->downc.push_back(make_conv(e, "downsample_convs." + std::to_string(i), e->chans[i], e->chans[i], 2,
                                   true, false, i != c.depth - 1));
    }
  int prev = c.in_channels;
  e->enc.resize(c.depth);
  for (int i = 0; i < c.depth; ++i) {
    e->creating_late = i < Ls;
    make_block(e, e->enc[i], "down_path." + std::to_string(i), prev, e->chans[i]);
    prev = e->chans[i];
  }
  e->creating_late = false;
  e->dec.resize(c.depth - 1);
  for (int j = 0; j < c.depth - 1; ++j) {
    const int lvl = c.depth - 2 - j;
    const int co = e->chans[lvl];
    e->upc.push_back(make_conv(e, "up_path." + std::to_string(j) + ".up", prev, co, 2, true, true, true));
    make_block(e, e->dec[j], "up_path." + std::to_string(j) + ".conv_block", prev, co);
    prev = co;
  }
  e->Cf = prev;
  e->seg = make_conv(e, "seg_conv", prev, c.n_classes, 1, false, false, true);
  if (c.num_lands > 0) {
    int nf = c.lands_num_1x1 > 1 ? c.num_lands + c.n_classes : c.num_lands;
    e->lands.push_back(make_conv(e, "lands_1x1.0", prev + c.n_classes, nf, 1, false, false, true));
    for (int k = 0; k < c.lands_num_1x1 - 1; ++k) {
      e->lands.push_back(make_conv(e, "lands_1x1." + std::to_string(k + 1), nf, c.num_lands, 1, false, false, true));
      nf = c.num_lands;
    }
  }
  e->Cpad = pad_to(e->Cf + c.n_classes, 8);
  // flat gradient layout: early bucket first, each gradient 16-byte aligned
  e->grad_numel = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (auto& t : e->tensors)
      if (t.info.grad_offset >= 0 && t.late == (pass == 1)) {
        t.info.grad_offset = e->grad_numel;
        e->grad_numel += (t.numel + 3) / 4 * 4;
      }
    if (pass == 0) e->early_numel = e->grad_numel;
  }
  return FU_OK;
}

// ---------------------------------------------------------------------------
// persistent device memory: packed weights, BN scratch
// ---------------------------------------------------------------------------

void conv_pack_sizes(ConvW& c, bool shuffle_fwd, bool shuffle_dgrad) {
  // forward GEMM columns / data-gradient GEMM columns
  if (c.transposed) {           // ConvTranspose2d(Cin,Cout,2,2): fwd = shuffle GEMM, dgrad = 2x2/s2 conv
    c.n_fwd = 4 * c.Cout; c.n_dgrad = c.Cin;
  } else if (c.k == 2) {        // Conv2d(C,C,2,stride 2): fwd = 2x2/s2 conv, dgrad = shuffle GEMM
    c.n_fwd = c.Cout; c.n_dgrad = 4 * c.Cin;
  } else {
    c.n_fwd = c.Cout; c.n_dgrad = c.Cin;
  }
  c.npad_fwd = pad_to(c.n_fwd, 4);
  c.npad_dgrad = pad_to(c.n_dgrad, 4);
  (void)shuffle_fwd; (void)shuffle_dgrad;
}

template <typename F>
void for_each_conv(fu_engine* e, F f) {
  for (auto& c : e->downc) f(c);
  for (auto& b : e->enc) { if (b.has_res) f(b.res); for (auto& c : b.convs) f(c); }
  for (auto& c : e->upc) f(c);
  for (auto& b : e->dec) { if (b.has_res) f(b.res); for (auto& c : b.convs) f(c); }
  f(e->seg);
  for (auto& c : e->lands) f(c);
}
template <typename F>
void for_each_bn(fu_engine* e, F f) {
  for (auto& b : e->enc) for (auto& bn : b.bns) f(bn);
  for (auto& b : e->dec) for (auto& bn : b.bns) f(bn);
}

void carve_persistent(fu_engine* e, Bump& w, Bump& df, Bump& db, Bump& ws) {
  int maxc = 4;
  for_each_conv(e, [&](ConvW& c) {
    conv_pack_sizes(c, false, false);
    const int taps_f = c.transposed ? 1 : c.k * c.k;
    const int k_f = c.Cin;
    c.wp_fwd = w.take<float>((size_t)taps_f * k_f * c.npad_fwd);
    const int taps_d = c.transposed ? 4 : (c.k == 2 ? 1 : c.k * c.k);
    c.wp_dgrad = w.take<float>((size_t)taps_d * c.Cout * c.npad_dgrad);
    c.bsum = db.take<double>((size_t)kRedCopies * c.Cout);      // kRedCopies copies (kernels_simt.cuh)
    c.small_cin = !c.transposed && (c.k == 3 || c.k == 1) && c.Cin <= 2 && (c.Cout % 8 == 0) && c.Cout <= 256 &&
                  (c.Cout & (c.Cout - 1)) == 0 && c.Cout / 8 <= 32;
    if (c.Cout > maxc) maxc = c.Cout;
    if (c.Cin > maxc) maxc = c.Cin;
    tc_carve(c.tc, c.Cin, c.Cout, c.k, c.transposed, e->cfg.precision != FU_PRECISION_FP32, w, ws, e->split);
  });
  for (auto& b : e->enc) b.dstat = db.take<double>(2 * (size_t)b.Cin);
  for (auto& b : e->dec) b.dstat = db.take<double>(2 * (size_t)b.Cin);
  for_each_bn(e, [&](BNL& b) {
    b.stat = df.take<double>(2 * b.C);
    b.bstat = db.take<double>((size_t)kRedCopies * 2 * b.C);
    b.mean = w.take<float>(b.C); b.invstd = w.take<float>(b.C);
    b.a = w.take<float>(b.C); b.b = w.take<float>(b.C);
    b.ga = w.take<float>(b.C); b.m1 = w.take<float>(b.C); b.m2 = w.take<float>(b.C);
  });
  e->ones = w.take<float>(maxc);
  e->zeros = w.take<float>(maxc);
  e->heads_gacc = w.take<float>((size_t)(e->cfg.num_lands + e->cfg.n_classes) * (e->Cf + e->cfg.n_classes) + 64);
  e->coop_bar = reinterpret_cast<unsigned*>(w.take<float>(4));     // (wmem is zeroed once at allocation)
}

int alloc_persistent(fu_engine* e) {
  Bump w, df, db, ws;
  carve_persistent(e, w, df, db, ws);
  e->wgrad_scr_bytes = ws.off + 256;
  CUDA_TRY(e, cudaMalloc(&e->wgrad_scr, e->wgrad_scr_bytes));
  e->wmem_bytes = w.off + 256;
  e->dscr_fwd_bytes = df.off + 256;
  e->dscr_bwd_bytes = db.off + 256;
  CUDA_TRY(e, cudaMalloc(&e->wmem, e->wmem_bytes));
  CUDA_TRY(e, cudaMalloc(&e->dscr_fwd, e->dscr_fwd_bytes));
  CUDA_TRY(e, cudaMalloc(&e->dscr_bwd, e->dscr_bwd_bytes));
  CUDA_TRY(e, cudaMemset(e->wmem, 0, e->wmem_bytes));
  CUDA_TRY(e, cudaMalloc(&e->pack_tbl, fu_engine::kJobCap * sizeof(TcPackJob)));
  CUDA_TRY(e, cudaMalloc(&e->unpack_tbl, fu_engine::kJobCap * sizeof(TcUnpackJob)));
  {
    const char* s = getenv("FU_STREAMS");
    e->use_side = (s && atoi(s) == 1) ? 0 : 1;
    if (e->use_side) {
      CUDA_TRY(e, cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
      CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    }
  }
  CUDA_TRY(e, cudaMallocHost(&e->pack_pin, fu_engine::kJobCap * sizeof(TcPackJob)));
  CUDA_TRY(e, cudaMallocHost(&e->unpack_pin, fu_engine::kJobCap * sizeof(TcUnpackJob)));
  Bump w2, df2, db2, ws2;
  w2.base = e->wmem; df2.base = reinterpret_cast<char*>(e->dscr_fwd); db2.base = reinterpret_cast<char*>(e->dscr_bwd);
  ws2.base = e->wgrad_scr;
  carve_persistent(e, w2, df2, db2, ws2);
  // ones / zeros
  int maxc = 4;
  for_each_conv(e, [&](ConvW& c) { if (c.Cout > maxc) maxc = c.Cout; if (c.Cin > maxc) maxc = c.Cin; });
  std::vector<float> h(maxc, 1.f);
  CUDA_TRY(e, cudaMemcpy(e->ones, h.data(), maxc * sizeof(float), cudaMemcpyHostToDevice));
  return FU_OK;
}

// ---------------------------------------------------------------------------
// plan: activation arena for (B,H,W)
// ---------------------------------------------------------------------------
View take_view(Bump& b, long long pixels, int C, int ld, int esz, const Plan* pl = nullptr) {
  View v;
  v.p = b.take<char>((size_t)pixels * ld * esz);
  v.C = C;
  v.ld = ld;
  // the twin of a buffer sits at the same offset of the twin arena (same bytes per pixel: 2 x ld bf16 = ld fp32)
  if (pl && pl->twin && v.p) v.sp = pl->twin + (v.p - pl->arena);
  return v;
}

void carve_plan(fu_engine* e, Plan& pl, Bump& b, int B, int H, int W) {
  const fu_config& c = e->cfg;
  const int esz = e->esz;
  const int D = c.depth;
  auto pix = [&](int lvl) { return (long long)B * (H >> lvl) * (W >> lvl); };
  pl.xin = take_view(b, pix(0), c.in_channels, c.in_channels, esz, &pl);
  pl.cat.assign(D, View()); pl.d_cat.assign(D, View());
  pl.down.assign(D, View()); pl.d_down.assign(D, View());
  pl.decout.assign(D, View()); pl.d_decout.assign(D, View());
  pl.hcat = take_view(b, pix(0), e->Cpad, e->Cpad, esz, &pl);
  pl.d_hcat = take_view(b, pix(0), e->Cpad, e->Cpad, esz, &pl);
  for (int l = 0; l < D - 1; ++l) {
    pl.cat[l] = take_view(b, pix(l), 2 * e->chans[l], 2 * e->chans[l], esz, &pl);
    pl.d_cat[l] = take_view(b, pix(l), 2 * e->chans[l], 2 * e->chans[l], esz, &pl);
    // halves narrower than a 128-byte line, written by the fused tensor-core data gradient of the decoder block (which
    // can store plane by plane): two planes instead of interleaved halves
    if (e->cfg.precision == FU_PRECISION_BF16 && e->chans[l] * esz < 128 && e->chans[l] % 32 == 0 && c.do_res &&
        !e->dec.empty() && e->dec[D - 2 - l].convs[0].tc.enabled && e->dec[D - 2 - l].res.tc.enabled &&
        (W >> l) >= tc_env_int("FU_TC_V2_MINW", 48) && (H >> l) >= 8 && tc_env_int("FU_DCAT_PLANAR", 1)) {
      pl.d_cat[l].plane = (size_t)pix(l) * e->chans[l] * esz;
      pl.d_cat[l].planeC = e->chans[l];
    }
  }
  for (int l = 1; l < D; ++l) {
    pl.down[l] = take_view(b, pix(l), e->chans[l - 1], e->chans[l - 1], esz, &pl);
    pl.d_down[l] = take_view(b, pix(l), e->chans[l - 1], e->chans[l - 1], esz, &pl);
  }
  if (D > 1) {
    pl.bott = take_view(b, pix(D - 1), e->chans[D - 1], e->chans[D - 1], esz, &pl);
    pl.d_bott = take_view(b, pix(D - 1), e->chans[D - 1], e->chans[D - 1], esz, &pl);
  }
  for (int l = 1; l < D - 1; ++l) {
    pl.decout[l] = take_view(b, pix(l), e->chans[l], e->chans[l], esz, &pl);
    pl.d_decout[l] = take_view(b, pix(l), e->chans[l], e->chans[l], esz, &pl);
  }
  // the last block of the network writes its output straight into the head's concat buffer
  pl.decout[0] = slice(pl.hcat, 0, e->Cf, esz);
  pl.d_decout[0] = slice(pl.d_hcat, 0, e->Cf, esz);
  auto carve_block = [&](Block& blk, int lvl, const View& outv) {
    const int nd = (int)blk.convs.size();
    blk.r.assign(nd, View()); blk.z.assign(nd, View()); blk.dy.assign(nd, View()); blk.dz.assign(nd, View());
    for (int i = 0; i < nd; ++i) {
      if (i == nd - 1 && blk.bns.empty() && !blk.has_res) blk.r[i] = outv;  // ReLU output IS the block output
      else blk.r[i] = take_view(b, pix(lvl), blk.C, blk.C, esz, &pl);
      if (!blk.bns.empty() && i < nd - 1) blk.z[i] = take_view(b, pix(lvl), blk.C, blk.C, esz, &pl);
      blk.dy[i] = take_view(b, pix(lvl), blk.C, blk.C, esz, &pl);
      if (i > 0) blk.dz[i] = take_view(b, pix(lvl), blk.C, blk.C, esz, &pl);
    }
  };
  for (int l = 0; l < D; ++l)
    carve_block(e->enc[l], l, D == 1 ? pl.decout[0] : (l < D - 1 ? slice(pl.cat[l], e->chans[l], e->chans[l], esz) : pl.bott));
  for (int j = 0; j < D - 1; ++j) carve_block(e->dec[j], D - 2 - j, pl.decout[D - 2 - j]);
  pl.hmid.clear(); pl.d_hmid.clear();
  for (size_t k = 0; k + 1 < e->lands.size(); ++k) {
    const int n = e->lands[k].Cout;
    pl.hmid.push_back(take_view(b, pix(0), n, pad_to(n, 8), esz, &pl));
    pl.d_hmid.push_back(take_view(b, pix(0), n, pad_to(n, 8), esz, &pl));
  }
  if (c.num_lands > 0) pl.dheat = take_view(b, pix(0), c.num_lands, pad_to(c.num_lands, 8), esz, &pl);
}

int ensure_plan(fu_engine* e, int B, int H, int W) {
  Plan& pl = e->plan;
  if (pl.valid && pl.B == B && pl.H == H && pl.W == W) return FU_OK;
  e->gaps_valid = false;          // a new plan may send other layers through the tensor-core weight gradients
  const int D = e->cfg.depth;
  if (B < 1 || H < 1 || W < 1 || (H % (1 << (D - 1))) || (W % (1 << (D - 1))))
    return e->fail(FU_ERR_UNSUPPORTED_SHAPE,
                   "input %dx%dx%d: H and W must be positive multiples of 2^(depth-1)=%d (the reference "
                   "pads tiles to --unet-img-dim, dataset.py:26-40)", B, H, W, 1 << (D - 1));
  if ((long long)B * H * W > (1ll << 31) - 1)
    return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "B*H*W too large");
  Bump dry;
  Plan tmp;
  carve_plan(e, tmp, dry, B, H, W);
  const size_t need = dry.off + 256;
  if (need > pl.arena_bytes) {
    if (pl.arena) {
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      CUDA_TRY(e, cudaFree(pl.arena));
      if (pl.twin) CUDA_TRY(e, cudaFree(pl.twin));
      pl.arena = nullptr; pl.twin = nullptr;
      pl.arena_bytes = 0;
    }
    CUDA_TRY(e, cudaMalloc(&pl.arena, need));
    if (e->split) CUDA_TRY(e, cudaMalloc(&pl.twin, need));
    pl.arena_bytes = need;
  }
  Bump real;
  real.base = pl.arena;
  carve_plan(e, pl, real, B, H, W);
  e->fresh_fwd.clear(); e->fresh_bwd.clear();
  pl.B = B; pl.H = H; pl.W = W;
  pl.valid = true;
  e->saved = false;
  e->cnt.arena_bytes = (int64_t)pl.arena_bytes;
  return FU_OK;
}

// ---------------------------------------------------------------------------
// weight packing (CUDA-core layouts); tensor-core layouts are packed in kernels_tc.cuh
// ---------------------------------------------------------------------------
int pack_one(fu_engine* e, const float* src, float* dst, int T, int K, int N, int Npad, int Ninner, int flip,
             long long st, long long sk, long long snh, long long snl) {
  e->set_tag(0, 0, "weight_pack");
  PackArgs a;
  a.src = src; a.dst = dst; a.T = T; a.K = K; a.N = N; a.Npad = Npad; a.Ninner = Ninner; a.flip = flip;
  a.st = st; a.sk = sk; a.snh = snh; a.snl = snl;
  const long long total = (long long)T * K * Npad;
  LAUNCH(e, pack_weights_kernel, grid1d(total, 256, e->num_sms), 256, a);
  return FU_OK;
}

int pack_conv(fu_engine* e, ConvW& c) {
  const float* w = reinterpret_cast<const float*>(e->tensors[c.w_idx].data);
  int rc;
  if (c.tc.enabled) {
    // tensor-core layer: only the bf16 packings are ever read
    if ((rc = tc_pack(c.tc, w, e->stream, &e->cnt))) return e->fail(FU_ERR_CUDA, "tensor-core weight pack failed");
    return FU_OK;
  }
  if (c.small_cin) return FU_OK;   // first-layer kernels read the torch layout directly
  if (c.transposed) {
    // W[ci][co][ab].  fwd: [1][Cin][(ab)*Cout+co];  dgrad (2x2/s2 conv over dY): [tap=ab][Cout][Cin]
    if ((rc = pack_one(e, w, c.wp_fwd, 1, c.Cin, 4 * c.Cout, c.npad_fwd, c.Cout, 0, 0, (long long)c.Cout * 4, 1, 4))) return rc;
    if ((rc = pack_one(e, w, c.wp_dgrad, 4, c.Cout, c.Cin, c.npad_dgrad, c.Cin, 0, 1, 4, 0, (long long)c.Cout * 4))) return rc;
  } else if (c.k == 2) {
    // W[co][ci][ab].  fwd: [tap=ab][Cin][Cout];  dgrad (shuffle GEMM): [1][Cout][(ab)*Cin+ci]
    if ((rc = pack_one(e, w, c.wp_fwd, 4, c.Cin, c.Cout, c.npad_fwd, c.Cout, 0, 1, 4, 0, (long long)c.Cin * 4))) return rc;
    if ((rc = pack_one(e, w, c.wp_dgrad, 1, c.Cout, 4 * c.Cin, c.npad_dgrad, c.Cin, 0, 0, (long long)c.Cin * 4, 1, 4))) return rc;
  } else {
    const int kk = c.k * c.k;
    // W[co][ci][tap].  fwd: [tap][Cin][Cout];  dgrad: [flip(tap)][Cout][Cin]
    if ((rc = pack_one(e, w, c.wp_fwd, kk, c.Cin, c.Cout, c.npad_fwd, c.Cout, 0, 1, kk, 0, (long long)c.Cin * kk))) return rc;
    if ((rc = pack_one(e, w, c.wp_dgrad, kk, c.Cout, c.Cin, c.npad_dgrad, c.Cin, 1, 1, (long long)c.Cin * kk, 0, kk))) return rc;
  }
  return FU_OK;
}

int pack_all(fu_engine* e, bool training) {
  int rc = FU_OK;
  const int last = e->cfg.depth - 1;
  int di = 0;
  // Tensor-core layers only register their job; one launch packs a whole list.  Two lists: the layers the forward
  // needs first (encoder levels below split_level(): a few hundred KB of weights) are packed in line, everything else
  // (97 % of the bytes: deep encoder levels, the whole decoder, the heads) on the side stream while the shallow
  // encoder levels run; forward_t joins before the first deep layer.
  const int Ls = e->split_level();
  e->batch.pack.clear();
  e->batch_deep.pack.clear();
  auto sink = [&](bool deep) { tc_batch() = deep ? &e->batch_deep : &e->batch; };
  for (auto& c : e->downc) { const int i = di++; sink(i >= Ls); if (i != last && rc == FU_OK) rc = pack_conv(e, c); }
  auto blk = [&](Block& b) {
    if (b.has_res && rc == FU_OK) rc = pack_conv(e, b.res);
    for (auto& c : b.convs) if (rc == FU_OK) rc = pack_conv(e, c);
  };
  for (int l = 0; l < (int)e->enc.size(); ++l) { sink(l >= Ls); blk(e->enc[l]); }
  sink(true);
  for (auto& c : e->upc) if (rc == FU_OK) rc = pack_conv(e, c);
  for (auto& b : e->dec) blk(b);
  if (rc == FU_OK) rc = pack_conv(e, e->seg);
  for (auto& c : e->lands) if (rc == FU_OK) rc = pack_conv(e, c);
  tc_batch() = nullptr;
  (void)training;
  if (rc != FU_OK) return rc;
  e->set_tag(0, 0, "weight_pack");
  if (!e->batch_deep.pack.empty()) {
    SideScope side(e);
    if (e->prof) e->prof_begin("tc_pack_batched_kernel");
    const int trc = tc_flush_jobs(e->batch_deep.pack, e->pack_uploaded2, e->pack_tbl + fu_engine::kJobCap / 2,
                                  fu_engine::kJobCap / 2, tc_pack_launcher(e->stream), e->stream, &e->cnt,
                                  e->pack_pin + fu_engine::kJobCap / 2);
    if (e->prof) e->prof_end();
    if (trc) return e->fail(FU_ERR_CUDA, "batched weight pack failed");
  }
  if (!e->batch.pack.empty()) {
    if (e->prof) e->prof_begin("tc_pack_batched_kernel");
    const int trc = tc_flush_jobs(e->batch.pack, e->pack_uploaded, e->pack_tbl, fu_engine::kJobCap / 2,
                                  tc_pack_launcher(e->stream), e->stream, &e->cnt, e->pack_pin);
    if (e->prof) e->prof_end();
    if (trc) return e->fail(FU_ERR_CUDA, "batched weight pack failed");
  }
  return rc;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
struct ConvCall {
  View x; int B = 0, Hi = 0, Wi = 0;
  View y; int Ho = 0, Wo = 0;
  const float* w = nullptr; int N = 0, Npad = 0;
  const float* bias = nullptr; int bias_mod = 1;
  int KH = 1, stride = 1, pad = 0;
  int relu = 0, accumulate = 0, shuffle = 0;
  float* nchw_out = nullptr;
  View t; const float* bn_a = nullptr; const float* bn_b = nullptr; bool has_t = false;
  double* stat = nullptr;
};

template <typename T>
int run_igemm(fu_engine* e, const ConvCall& c) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = c.x.p; a.x_ld = c.x.ld;
  a.y = c.nchw_out ? (void*)c.nchw_out : (void*)c.y.p; a.y_ld = c.y.ld;
  a.w = c.w; a.bias = c.bias; a.bias_mod = c.bias_mod > 0 ? c.bias_mod : 1;
  a.B = c.B; a.Hi = c.Hi; a.Wi = c.Wi; a.Cin = c.x.C;
  a.Ho = c.Ho; a.Wo = c.Wo; a.N = c.N; a.Npad = c.Npad;
  a.KH = c.KH; a.KW = c.KH; a.stride = c.stride; a.pad = c.pad;
  a.relu = c.relu; a.accumulate = c.accumulate; a.shuffle = c.shuffle;
  a.nchw_out = c.nchw_out ? 1 : 0;
  if (c.has_t) { a.t = c.t.p; a.t_ld = c.t.ld; a.bn_a = c.bn_a; a.bn_b = c.bn_b; }
  a.stat = c.stat;
  const size_t va = 4 * sizeof(T);
  const bool vec_in = (c.x.C % 4 == 0) && (c.x.ld % 4 == 0) && aligned(c.x.p, va);
  bool vec_out = !c.nchw_out && (c.N % 4 == 0) && (c.y.ld % 4 == 0) && aligned(c.y.p, va);
  if (c.shuffle && ((c.N / 4) % 4 != 0)) vec_out = false;
  if (c.has_t && !((c.t.ld % 4 == 0) && aligned(c.t.p, va))) vec_out = false;
  a.vec_out = vec_out ? 1 : 0;
  const long long M = (long long)c.B * c.Ho * c.Wo;
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((c.N + 63) / 64));
  if (vec_in) LAUNCH(e, (igemm_simt_kernel<T, true>), grid, 256, a);
  else LAUNCH(e, (igemm_simt_kernel<T, false>), grid, 256, a);
  return FU_OK;
}

struct WgradCall {
  View big; int Hb = 0, Wb = 0;
  View small; int Hs = 0, Ws = 0;
  int B = 0, KH = 1, stride = 1, pad = 0;
  float* dw = nullptr; long long s_tap = 0, s_big = 0, s_small = 0;
};

template <typename T>
int run_wgrad(fu_engine* e, const WgradCall& c) {
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.big = c.big.p; a.big_ld = c.big.ld; a.Hb = c.Hb; a.Wb = c.Wb; a.Cb = c.big.C;
  a.small = c.small.p; a.small_ld = c.small.ld; a.Hs = c.Hs; a.Ws = c.Ws; a.Cs = c.small.C;
  a.B = c.B; a.KH = c.KH; a.KW = c.KH; a.stride = c.stride; a.pad = c.pad;
  a.dw = c.dw; a.s_tap = c.s_tap; a.s_big = c.s_big; a.s_small = c.s_small;
  const int tb = (c.big.C + 63) / 64, ts = (c.small.C + 63) / 64, taps = c.KH * c.KH;
  a.tiles_small = ts;
  const long long M = (long long)c.B * c.Hs * c.Ws;
  long long tiles = (long long)tb * ts * taps;
  long long splits = ((long long)e->num_sms * 4 + tiles - 1) / tiles;
  const long long max_splits = (M + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = (M + splits - 1) / splits;
  pps = (pps + 15) / 16 * 16;
  splits = (M + pps - 1) / pps;
  a.pix_per_split = (int)pps;
  const size_t va = 4 * sizeof(T);
  const bool vec = (c.big.C % 4 == 0) && (c.big.ld % 4 == 0) && aligned(c.big.p, va) &&
                   (c.small.C % 4 == 0) && (c.small.ld % 4 == 0) && aligned(c.small.p, va);
  dim3 grid((unsigned)(tb * ts), (unsigned)taps, (unsigned)splits);
  if (vec) LAUNCH(e, (wgrad_simt_kernel<T, true>), grid, 256, a);
  else LAUNCH(e, (wgrad_simt_kernel<T, false>), grid, 256, a);
  return FU_OK;
}

// bps: resident blocks per SM of the kernel the grid is for (its __launch_bounds__): the grid is capped at whole waves of them
inline dim3 red_grid(fu_engine* e, long long P, int C, int max_lanes = kRedLanes, int bps = 4) {
  const int cvecs = C / (e->esz == 2 ? 8 : 4);     // Vec<T>::N channels per thread
  const int lanes = cvecs < max_lanes ? cvecs : max_lanes;
  const int rows = 256 / lanes;
  // 16 x `rows` pixel rows per block on the big levels; fewer (down to 2) on the small, wide ones until ~1.5 blocks per
  // SM exist.  A block covers at most kRedLanes channel vectors (grid.y = channel groups), so more blocks no longer
  // mean proportionally more fp64 atomics: round 1 had blocks spanning all C channels, where finer grids were 1.6x
  // slower, and an 8-block cluster / DSMEM pre-reduction made the big levels pay for cluster scheduling.
  long long per = 16;
  static const int per_env = tc_env_int("FU_RED_PER_MIN", 2);
  const int per_min = max_lanes == kRedLanes ? per_env : 16;      // (the forward apply kernel keeps its round-1 grid)
  const unsigned gy = (unsigned)((cvecs + lanes - 1) / lanes);
  while (per > per_min && (P + rows * per - 1) / (rows * per) * gy < (long long)e->num_sms * 3 / 2) per >>= 1;
  long long gx = (P + (long long)rows * per - 1) / ((long long)rows * per);
  static const int waves = tc_env_int("FU_RED_WAVES", 1);   // measured @192x192: apply 47.4 (2 waves) -> 42.4 us (1), reduce 38.3 -> 35.4
  const long long cap = std::max<long long>((long long)e->num_sms * bps * (max_lanes == kRedLanes ? waves : 2) / gy, 1);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, gy);
}

float* tdata(fu_engine* e, int idx) { return idx < 0 ? nullptr : reinterpret_cast<float*>(e->tensors[idx].data); }

// ---------------------------------------------------------------------------
// parity_tc mode: MMA operands are read from split-bf16 twins of the fp32 tensors
// ---------------------------------------------------------------------------
struct Opnd { const void* p; int ld; };
// operand descriptor of a view as the tensor-core kernels take it: the view itself, or its twin (hi half, twin stride)
inline Opnd opnd(const fu_engine* e, const View& v) {
  return e->split ? Opnd{v.sp, 2 * v.ld} : Opnd{v.p, v.ld};
}
// make the twin of `v` (P pixels) current on the engine's CURRENT stream.  Call it on the main stream before forking
// work that reads the twin to the side stream.
int ensure_split(fu_engine* e, const View& v, long long P) {
  if (!e->split || !v.sp || (v.C % 4) || (v.ld % 4)) return FU_OK;
  const std::pair<const void*, int> key(v.p, v.C);
  for (auto& k : e->fresh_fwd) if (k == key) return FU_OK;
  for (auto& k : e->fresh_bwd) if (k == key) return FU_OK;
  e->set_tag(0, 8.0 * P * v.C, "split_twin");
  if (e->prof) e->prof_begin("split_f32_kernel");
  const int rc = tc_split(reinterpret_cast<const float*>(v.p), v.ld, v.C, reinterpret_cast<bf16*>(v.sp), 2 * v.ld, v.ld, P,
                          e->num_sms, e->stream, &e->cnt);
  if (e->prof) e->prof_end();
  if (rc) return e->fail(FU_ERR_CUDA, "split_f32_kernel launch failed");
  (e->in_backward ? e->fresh_bwd : e->fresh_fwd).push_back(key);
  return FU_OK;
}
float* gptr(fu_engine* e, float* flat, int idx) {
  if (idx < 0) return nullptr;
  const int64_t off = e->tensors[idx].info.grad_offset;
  return off < 0 ? nullptr : flat + off;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <typename T>
int conv_forward(fu_engine* e, ConvW& cw, const View& x, const View& y, int B, int H, int W, int relu,
                 double* stat, const View* t, const float* bn_a, const float* bn_b, const TcBnFin* fin = nullptr,
                 int post = 0) {
  // 3x3/pad1 or 1x1 convolution, stride 1
  {
    const double M = (double)B * H * W;
    e->set_tag(2.0 * M * cw.Cin * cw.Cout * cw.k * cw.k, (M * (x.C + y.C)) * e->esz + 2.0 * cw.Cin * cw.Cout * cw.k * cw.k,
               "conv%d_fwd %dx%d %d->%d", cw.k, H, W, cw.Cin, cw.Cout);
  }
  const Opnd xo = opnd(e, x);
  if (tc_conv_eligible(cw.tc, xo.p, xo.ld, y.p, y.ld, t ? t->p : nullptr, t ? t->ld : 0)) {
    { const int src = ensure_split(e, x, (long long)B * H * W); if (src) return src; }
    if (e->prof) e->prof_begin("tc_conv_kernel");
    int rc = tc_conv_forward(cw.tc, xo.p, xo.ld, y.p, y.ld, B, H, W, tdata(e, cw.b_idx), relu, stat,
                             t ? t->p : nullptr, t ? t->ld : 0, bn_a, bn_b, 0, e->stream, &e->cnt, fin, post);
    if (e->prof) e->prof_end();
    if (rc) return e->fail(FU_ERR_CUDA, "tensor-core conv launch failed: %s", tc_last_error());
    return FU_OK;
  }
  if (cw.tc.enabled) return e->fail(FU_ERR_STATE, "tensor-core layer with a misaligned view (internal error)");
  if (post) return e->fail(FU_ERR_STATE, "post-ReLU BatchNorm folding is a tensor-core epilogue feature (internal error)");
  const size_t va = 4 * sizeof(T);
  if (cw.small_cin && (y.ld % 4 == 0) && aligned(y.p, va) && (!t || ((t->ld % 4 == 0) && aligned(t->p, va)))) {
    SmallCinArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x.p; a.x_ld = x.ld; a.Cin = cw.Cin; a.y = y.p; a.y_ld = y.ld; a.Cout = cw.Cout;
    a.w = tdata(e, cw.w_idx); a.bias = tdata(e, cw.b_idx);
    a.B = B; a.H = H; a.W = W; a.k = cw.k; a.pad = cw.k / 2; a.relu = relu;
    if (t) { a.t = t->p; a.t_ld = t->ld; a.bn_a = bn_a; a.bn_b = bn_b; }
    a.stat = stat;
    if constexpr (sizeof(T) == 2) {
      // bf16 storage, 3x3, 32 output channels, no second epilogue operand: warp-level tensor-core MMAs (conv_cin1_mma_kernel)
      if (cw.Cin == 1 && cw.Cout == 32 && cw.k == 3 && !t && W % 16 == 0 && y.ld % 8 == 0 && aligned(y.p, 16) &&
          (long long)B * H * W < (1ll << 30) && tc_env_int("FU_CIN1_MMA", 1)) {
        const long long tiles = (long long)B * H * (W / 16);
        const unsigned grid = (unsigned)std::min<long long>((tiles + 3) / 4, (long long)e->num_sms * tc_env_int("FU_CIN1_BPS", 16));      // blocks per SM: 8 -> 48.7 us, 16 -> 43.4 us (finer tail)
        LAUNCH(e, (conv_cin1_mma_kernel<3>), grid, 128, reinterpret_cast<const bf16*>(x.p), x.ld, reinterpret_cast<bf16*>(y.p), y.ld,
               tdata(e, cw.w_idx), tdata(e, cw.b_idx), relu, stat, B, H, W);
        return FU_OK;
      }
    }
    if (cw.Cin == 1) {
      // C_in = 1 fast path: persistent blocks, two per SM
      const int TW = 256 / (cw.Cout / 8);
      const long long tiles = (long long)((W + TW - 1) / TW) * ((H + cin1_tile_h(cw.k) - 1) / cin1_tile_h(cw.k)) * B;
      // whole waves of the kernel's resident blocks (3 per SM for the 1x1, 2 for the 3x3 variant): 4 x SMs was 1.33 waves
      // of the 1x1 kernel
      const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)e->num_sms * (cw.k == 1 ? 3 : 2) * tc_env_int("FU_CIN1_WAVES", 1));
      const size_t sm1 = cin1_smem_bytes(cw.k, cw.Cout, false);
      if (cw.k == 3) LAUNCH_SMEM(e, (conv_cin1_kernel<T, 3>), grid, 256, sm1, a);
      else LAUNCH_SMEM(e, (conv_cin1_kernel<T, 1>), grid, 256, sm1, a);
      return FU_OK;
    }
    const long long total = (long long)B * H * W * (cw.Cout / 8);
    const size_t smem = ((size_t)cw.k * cw.k * cw.Cin * cw.Cout + 3 * cw.Cout + 16 * cw.Cout) * sizeof(float);
    auto kfn = conv_small_cin_kernel<T>;
    if (e->prof) e->prof_begin("conv_small_cin_kernel");
    fu_launch(kfn, dim3(grid1d(total, 256, e->num_sms)), dim3(256), smem, e->stream, fu_pdl_enabled(), a);
    if (e->prof) e->prof_end();
    e->cnt.kernel_launches++;
    if (cudaPeekAtLastError() != cudaSuccess) return e->fail(FU_ERR_CUDA, "conv_small_cin_kernel launch failed");
    return FU_OK;
  }
  ConvCall c;
  c.x = x; c.B = B; c.Hi = H; c.Wi = W; c.y = y; c.Ho = H; c.Wo = W;
  c.w = cw.wp_fwd; c.N = cw.n_fwd; c.Npad = cw.npad_fwd;
  c.bias = tdata(e, cw.b_idx); c.bias_mod = cw.Cout;
  c.KH = cw.k; c.stride = 1; c.pad = cw.k / 2; c.relu = relu; c.stat = stat;
  if (t) { c.t = *t; c.has_t = true; c.bn_a = bn_a; c.bn_b = bn_b; }
  return run_igemm<T>(e, c);
}

template <typename T>
int block_forward(fu_engine* e, Block& blk, const View& x_in, const View& out, int B, int H, int W, int training,
                  bool eval_fast = false) {
  const int nd = (int)blk.convs.size();
  const bool bn = !blk.bns.empty();
  const long long P = (long long)B * H * W;
  View cur = x_in;
  int rc;
  TcBnFin res_fin;
  memset(&res_fin, 0, sizeof(res_fin));
  bool use_res_fin = false;
  for (int i = 0; i < nd; ++i) {
    View r = blk.r[i];   // (aliases `out` when the block ends in a bare ReLU, see carve_plan)
    double* stat = (bn && training) ? blk.bns[i].stat : nullptr;
    if (eval_fast && bn && (i < nd - 1 || !blk.has_res)) {
      // Inference fast path (no gradient will be asked for): the BatchNorm's folded running-statistics scale / shift
      // (bn_eval_coeffs_kernel, once per forward) ride in the convolution's epilogue, z = a * relu(conv) + b; the
      // post-ReLU tensor r is never written and the separate normalisation pass (one read + one write of the tensor) is
      // gone.  Only for tensor-core layers; the C_in = 1 first layer keeps the two-pass form.
      View z = (i == nd - 1) ? out : blk.z[i];
      const Opnd xo = opnd(e, cur);
      if (tc_conv_eligible(blk.convs[i].tc, xo.p, xo.ld, z.p, z.ld, nullptr, 0)) {
        if ((rc = conv_forward<T>(e, blk.convs[i], cur, z, B, H, W, 1, nullptr, nullptr, blk.bns[i].a, blk.bns[i].b, nullptr, 1))) return rc;
        cur = z;
        continue;
      }
    }
    if ((rc = conv_forward<T>(e, blk.convs[i], cur, r, B, H, W, 1, stat, nullptr, nullptr, nullptr))) return rc;
    if (bn) {
      BNL& b = blk.bns[i];
      e->set_tag(0, 2.0 * P * b.C * e->esz, "bn_fwd %dx%d C%d", H, W, b.C);
      if (i < nd - 1 || !blk.has_res) {
        // finalise + apply in one launch
        View z = (i == nd - 1) ? out : blk.z[i];
        BnFwdFin f;
        f.stat = b.stat; f.gamma = tdata(e, b.i_gamma); f.beta = tdata(e, b.i_beta); f.rmean = tdata(e, b.i_rm);
        f.rvar = tdata(e, b.i_rv); f.nbt = reinterpret_cast<long long*>(e->tensors[b.i_nbt].data);
        f.mean_o = b.mean; f.invstd_o = b.invstd; f.a_o = b.a; f.b_o = b.b; f.training = training;
        f.momentum = 0.1f; f.eps = 1e-5f;
        LAUNCH(e, (bn_finalize_apply_kernel<T>), red_grid(e, P, b.C, 256).x, 256, reinterpret_cast<const T*>(r.p), r.ld,
               reinterpret_cast<T*>(z.p), z.ld, f, P, b.C);
        cur = z;
      } else {
        // the last BN of a residual block is applied inside the residual conv's epilogue; on the tensor-core path it
        // is also finalised there (res_fin), otherwise by a one-block kernel
        if (tc_conv_eligible(blk.res.tc, opnd(e, x_in).p, opnd(e, x_in).ld, out.p, out.ld, r.p, r.ld)) {
          res_fin.stat = b.stat; res_fin.P = P; res_fin.training = training;
          res_fin.gamma = tdata(e, b.i_gamma); res_fin.beta = tdata(e, b.i_beta);
          res_fin.rmean = tdata(e, b.i_rm); res_fin.rvar = tdata(e, b.i_rv);
          res_fin.nbt = reinterpret_cast<long long*>(e->tensors[b.i_nbt].data);
          res_fin.momentum = 0.1f; res_fin.eps = 1e-5f; res_fin.mean_o = b.mean; res_fin.invstd_o = b.invstd;
          use_res_fin = true;
        } else {
          LAUNCH(e, bn_finalize_kernel, (b.C + 127) / 128, 128, b.stat, P, b.C, training, tdata(e, b.i_gamma),
                 tdata(e, b.i_beta), tdata(e, b.i_rm), tdata(e, b.i_rv),
                 reinterpret_cast<long long*>(e->tensors[b.i_nbt].data), 0.1f, 1e-5f, b.mean, b.invstd, b.a, b.b);
        }
        cur = r;
      }
    } else {
      cur = r;
    }
  }
  if (blk.has_res) {
    // out = BN_last(r_last) + res_conv1x1(x_in)   (unet.py:229-231), one pass
    const View& t = blk.r[nd - 1];
    const float* a = bn ? blk.bns[nd - 1].a : e->ones;
    const float* b = bn ? blk.bns[nd - 1].b : e->zeros;
    if ((rc = conv_forward<T>(e, blk.res, x_in, out, B, H, W, 0, nullptr, &t, a, b, use_res_fin ? &res_fin : nullptr))) return rc;
  }
  return FU_OK;
}

template <typename T>
int forward_t(fu_engine* e, const float* x, int B, int H, int W, int training, float* seg, float* logits,
              float* heat, bool eval_fast = false) {
  const fu_config& c = e->cfg;
  Plan& pl = e->plan;
  const int D = c.depth;
  const int esz = e->esz;
  int rc;
  if (training && c.batch_norm) CUDA_TRY(e, cudaMemsetAsync(e->dscr_fwd, 0, e->dscr_fwd_bytes, e->stream));
  if (eval_fast && c.batch_norm) {
    BnEvalTable t;
    t.count = 0; t.eps = 1e-5f;
    bool fits = true;
    for_each_bn(e, [&](BNL& b) {
      if (t.count >= BnEvalTable::kMax) { fits = false; return; }
      const int k = t.count++;
      t.gamma[k] = tdata(e, b.i_gamma); t.beta[k] = tdata(e, b.i_beta); t.rmean[k] = tdata(e, b.i_rm); t.rvar[k] = tdata(e, b.i_rv);
      t.a[k] = b.a; t.b[k] = b.b; t.mean_o[k] = b.mean; t.invstd_o[k] = b.invstd; t.C[k] = b.C;
    });
    if (!fits) eval_fast = false;
    else {
      e->set_tag(0, 0, "bn_eval_coeffs");
      LAUNCH(e, bn_eval_coeffs_kernel, t.count, 256, t);
    }
  }
  const long long HW = (long long)H * W;
  e->set_tag(0, 0, "input_cast");
  LAUNCH(e, (nchw_to_nhwc_kernel<T>), grid1d((long long)B * HW, 256, e->num_sms), 256, x,
         reinterpret_cast<T*>(pl.xin.p), pl.xin.ld, B, c.in_channels, HW);
  View cur = pl.xin;
  for (int l = 0; l < D; ++l) {
    const int h = H >> l, w = W >> l;
    View outv = (D == 1) ? pl.decout[0] : (l < D - 1 ? slice(pl.cat[l], e->chans[l], e->chans[l], esz) : pl.bott);
    if (l == e->split_level()) side_join(e);        // the deep layers' weight packs (side stream) are needed from here on
    if ((rc = block_forward<T>(e, e->enc[l], cur, outv, B, h, w, training, eval_fast))) return rc;
    if (l < D - 1) {
      View dn = pl.down[l + 1];
      e->set_tag(2.0 * B * (h / 2) * (w / 2) * 4.0 * outv.C * outv.C, 0, "down_fwd %dx%d C%d", h, w, outv.C);
      if (c.max_pool) {
        LAUNCH(e, (maxpool_fwd_kernel<T>), grid1d((long long)B * (h / 2) * (w / 2) * (outv.C / 4), 256, e->num_sms),
               256, reinterpret_cast<const T*>(outv.p), outv.ld, reinterpret_cast<T*>(dn.p), dn.ld, B, h / 2,
               w / 2, outv.C);
      } else {
        ConvW& cw = e->downc[l];
        const Opnd oo = opnd(e, outv);
        if (tc_down_eligible(cw.tc, oo.p, oo.ld, dn.p, dn.ld)) {
          if ((rc = ensure_split(e, outv, (long long)B * h * w))) return rc;
          if (e->prof) e->prof_begin("tc_conv_kernel");
          const int trc = tc_down_forward(cw.tc, oo.p, oo.ld, dn.p, dn.ld, B, h, w, tdata(e, cw.b_idx), e->stream, &e->cnt);
          if (e->prof) e->prof_end();
          if (trc)
            return e->fail(FU_ERR_CUDA, "tensor-core downsample launch failed: %s", tc_last_error());
        } else {
          ConvCall cc;
          cc.x = outv; cc.B = B; cc.Hi = h; cc.Wi = w; cc.y = dn; cc.Ho = h / 2; cc.Wo = w / 2;
          cc.w = cw.wp_fwd; cc.N = cw.n_fwd; cc.Npad = cw.npad_fwd; cc.bias = tdata(e, cw.b_idx);
          cc.bias_mod = cw.Cout; cc.KH = 2; cc.stride = 2; cc.pad = 0;
          if ((rc = run_igemm<T>(e, cc))) return rc;
        }
      }
      cur = dn;
    } else {
      cur = outv;
    }
  }
  side_join(e);                                     // (no-op unless the network is shallower than the split level)
  for (int j = 0; j < D - 1; ++j) {
    const int l = D - 2 - j;
    const int h = H >> l, w = W >> l;
    ConvW& up = e->upc[j];
    View upv = slice(pl.cat[l], 0, e->chans[l], esz);
    e->set_tag(2.0 * B * (h / 2) * (w / 2) * 4.0 * up.Cin * up.Cout, 0, "up_fwd %dx%d %d->%d", h, w, up.Cin, up.Cout);
    const Opnd co = opnd(e, cur);
    if (tc_up_eligible(up.tc, co.p, co.ld, upv.p, upv.ld)) {
      if ((rc = ensure_split(e, cur, (long long)B * (h / 2) * (w / 2)))) return rc;
      if (e->prof) e->prof_begin("tc_conv_kernel");
      const int trc = tc_up_forward(up.tc, co.p, co.ld, upv.p, upv.ld, B, h / 2, w / 2, tdata(e, up.b_idx), e->stream, &e->cnt);
      if (e->prof) e->prof_end();
      if (trc)
        return e->fail(FU_ERR_CUDA, "tensor-core upconv launch failed: %s", tc_last_error());
    } else {
      ConvCall cc;
      cc.x = cur; cc.B = B; cc.Hi = h / 2; cc.Wi = w / 2; cc.y = upv; cc.Ho = h / 2; cc.Wo = w / 2;
      cc.w = up.wp_fwd; cc.N = up.n_fwd; cc.Npad = up.npad_fwd; cc.bias = tdata(e, up.b_idx);
      cc.bias_mod = up.Cout; cc.KH = 1; cc.stride = 1; cc.pad = 0; cc.shuffle = 1;
      if ((rc = run_igemm<T>(e, cc))) return rc;
    }
    if ((rc = block_forward<T>(e, e->dec[j], pl.cat[l], pl.decout[l], B, h, w, training, eval_fast))) return rc;
    cur = pl.decout[l];
  }
  // ---- heads (unet.py:176-191) ----
  e->set_tag(0, 0, "heads_fwd");
  View feat = slice(pl.hcat, 0, e->Cf, esz);
  View lg = slice(pl.hcat, e->Cf, c.n_classes, esz);
  if (e->Cf == 32 && c.n_classes == 7 && (c.num_lands == 0 || (c.num_lands == 14 && e->lands.size() == 2 && e->lands[0].Cout == 21))) {
    // paper heads (7 classes, 39 -> 21 -> 14): one fused pass
    const long long P = (long long)B * HW;
    const unsigned gridf = (unsigned)std::min<long long>((P + 255) / 256, (long long)e->num_sms * 4);   // 4 resident blocks/SM (126 regs)
    if constexpr (sizeof(T) == 2) {
      // bf16 storage: the two 1x1 products as warp-level tensor-core MMAs (heads_fwd_mma_kernel); FU_HEADS_MMA=0: CUDA cores
      const bool use_mma = tc_env_int("FU_HEADS_MMA", 1) != 0;
      if (use_mma && P < (1ll << 30) && feat.ld % 8 == 0 && reinterpret_cast<uintptr_t>(feat.p) % 16 == 0) {
        const unsigned gridm = (unsigned)std::min<long long>((P + 63) / 64, (long long)e->num_sms * 6);      // 6 resident blocks per SM (launch bounds)
        HeadsLoss hl;
        memset(&hl, 0, sizeof(hl));
        if (e->lossf) {
          // the loss sums leave the head kernel (3 resident blocks per SM: 26 fp64 accumulators per thread)
          hl = *e->lossf;
          if (e->lossf_noseg) seg = nullptr;
          const unsigned gridl = (unsigned)std::min<long long>((P + 63) / 64, (long long)e->num_sms * 3);
          if (c.num_lands == 14)
            LAUNCH(e, (heads_fwd_mma_kernel<32, 7, 21, 14, true>), gridl, 128,
                   reinterpret_cast<const bf16*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx),
                   tdata(e, e->lands[1].w_idx), reinterpret_cast<bf16*>(lg.p), seg, logits, heat, (int)P, (int)HW, FastDiv((int)HW),
                   c.do_soft_max, hl);
          else
            LAUNCH(e, (heads_fwd_mma_kernel<32, 7, 1, 0, true>), gridl, 128,
                   reinterpret_cast<const bf16*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), (const float*)nullptr,
                   (const float*)nullptr, reinterpret_cast<bf16*>(lg.p), seg, logits, (float*)nullptr, (int)P, (int)HW,
                   FastDiv((int)HW), c.do_soft_max, hl);
          return FU_OK;
        }
        if (c.num_lands == 14)
          LAUNCH(e, (heads_fwd_mma_kernel<32, 7, 21, 14>), gridm, 128,
                 reinterpret_cast<const bf16*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx),
                 tdata(e, e->lands[1].w_idx), reinterpret_cast<bf16*>(lg.p), seg, logits, heat, (int)P, (int)HW, FastDiv((int)HW),
                 c.do_soft_max, hl);
        else
          LAUNCH(e, (heads_fwd_mma_kernel<32, 7, 1, 0>), gridm, 128,
                 reinterpret_cast<const bf16*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), (const float*)nullptr,
                 (const float*)nullptr, reinterpret_cast<bf16*>(lg.p), seg, logits, (float*)nullptr, (int)P, (int)HW,
                 FastDiv((int)HW), c.do_soft_max, hl);
        return FU_OK;
      }
    }
    if (e->lossf) return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_forward_loss: the fused loss needs bf16 storage and the tensor-core heads (FU_HEADS_MMA)");
    if (c.num_lands == 14)
      LAUNCH(e, (heads_fwd_fused_kernel<T, 32, 7, 21, 14>), gridf, 128,
             reinterpret_cast<const T*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx),
             tdata(e, e->lands[1].w_idx), reinterpret_cast<T*>(lg.p), seg, logits, heat, B, HW, c.do_soft_max);
    else
      LAUNCH(e, (heads_fwd_fused_kernel<T, 32, 7, 1, 0>), gridf, 128,
             reinterpret_cast<const T*>(feat.p), feat.ld, tdata(e, e->seg.w_idx), (const float*)nullptr,
             (const float*)nullptr, reinterpret_cast<T*>(lg.p), seg, logits, (float*)nullptr, B, HW,
             c.do_soft_max);
    return FU_OK;
  }
  {
    ConvCall cc;
    cc.x = feat; cc.B = B; cc.Hi = H; cc.Wi = W; cc.y = lg; cc.Ho = H; cc.Wo = W;
    cc.w = e->seg.wp_fwd; cc.N = e->seg.n_fwd; cc.Npad = e->seg.npad_fwd;
    if ((rc = run_igemm<T>(e, cc))) return rc;
  }
  LAUNCH(e, (softmax_fwd_kernel<T>), grid1d((long long)B * HW, 128, e->num_sms), 128,
         reinterpret_cast<const T*>(lg.p), lg.ld, B, c.n_classes, HW, c.do_soft_max, seg, logits);
  if (c.num_lands > 0) {
    View hc = slice(pl.hcat, 0, e->Cf + c.n_classes, esz);
    for (size_t k = 0; k < e->lands.size(); ++k) {
      ConvCall cc;
      cc.x = hc; cc.B = B; cc.Hi = H; cc.Wi = W; cc.Ho = H; cc.Wo = W;
      cc.w = e->lands[k].wp_fwd; cc.N = e->lands[k].n_fwd; cc.Npad = e->lands[k].npad_fwd;
      if (k + 1 == e->lands.size()) cc.nchw_out = heat;
      else cc.y = pl.hmid[k];
      if ((rc = run_igemm<T>(e, cc))) return rc;
      if (k + 1 < e->lands.size()) hc = pl.hmid[k];
    }
  }
  return FU_OK;
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
template <typename T>
int channel_sum_to(fu_engine* e, const View& d, long long P, double* scratch, float* dst) {
  LAUNCH(e, (channel_sum_kernel<T>), red_grid(e, P, d.C), 256, reinterpret_cast<const T*>(d.p), d.ld, P, d.C, scratch);
  e->deferred_sums.push_back({scratch, dst, d.C, kRedCopies});     // converted to fp32 by one launch at the end of backward
  return FU_OK;
}

__global__ void __launch_bounds__(256) zero_ranges_kernel(float* base, const long long* __restrict__ ranges /* (offset, floats) pairs */, int n) {
  pdl_wait(); pdl_trigger();
  // (every block walks every range: the ranges are few (~100) and of very different lengths)
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < n; ++r) {
    float* p = base + ranges[2 * r];
    const long long len = ranges[2 * r + 1];
    for (long long i = t0; i < len; i += step) p[i] = 0.f;
  }
}

// After a backward: which slices of the flat gradient did the unpack launches overwrite?  The first time (per plan) the gaps
// between them are tabulated for zero_ranges_kernel; afterwards the coverage must not change (same plan, same kernels) --
// a backward that covered something else than its zeroing assumed fails loudly.
int update_flat_gaps(fu_engine* e) {
  std::sort(e->cover_now.begin(), e->cover_now.end());
  unsigned long long h = 1469598103934665603ull;
  for (const auto& c : e->cover_now) { h = (h ^ (unsigned long long)c.first) * 1099511628211ull; h = (h ^ (unsigned long long)c.second) * 1099511628211ull; }
  if (e->gaps_valid) {
    if (h != e->cover_hash) {
      e->gaps_valid = false;
      return e->fail(FU_ERR_STATE, "internal: the weight-gradient unpack coverage changed between two backward passes of one plan");
    }
    return FU_OK;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(e->stream, &cap);
  if (cap != cudaStreamCaptureStatusNone || !tc_env_int("FU_FLAT_GAPS", 1)) return FU_OK;   // (tabulated outside captures only)
  std::vector<long long> tbl;
  long long pos = 0;
  for (const auto& c : e->cover_now) {
    if (c.first < pos) return FU_OK;                      // overlapping slices: keep the full memset
    if (c.first > pos) { tbl.push_back(pos); tbl.push_back(c.first - pos); }
    pos = c.first + c.second;
  }
  if (pos > e->grad_numel) return FU_OK;
  if (pos < e->grad_numel) { tbl.push_back(pos); tbl.push_back(e->grad_numel - pos); }
  const int n = (int)(tbl.size() / 2);
  long long gap_total = 0;
  for (int i = 0; i < n; ++i) gap_total += tbl[2 * i + 1];
  if (gap_total * 4 > e->grad_numel || n > 4096) return FU_OK;      // little is overwritten (CUDA-core modes): keep the plain memset
  if (n > e->gap_cap) {
    if (e->gap_tbl) { cudaStreamSynchronize(e->stream); cudaFree(e->gap_tbl); }
    e->gap_tbl = nullptr; e->gap_cap = 0;
    CUDA_TRY(e, cudaMalloc(&e->gap_tbl, (size_t)(n + 16) * 2 * sizeof(long long)));
    e->gap_cap = n + 16;
  }
  // (ordered on the engine's stream: an earlier backward's zero_ranges_kernel may still be reading the old table)
  e->gap_host = tbl;
  if (n > 0) CUDA_TRY(e, cudaMemcpyAsync(e->gap_tbl, e->gap_host.data(), e->gap_host.size() * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  e->n_gaps = n; e->cover_hash = h; e->gaps_valid = true;
  return FU_OK;
}

// [taps][M][N] accumulators of the tensor-core weight gradients registered so far -> torch layout, one launch.
// part 0 = the mid-backward flush (device table half 0), part 1 = the final one.
int flush_unpack(fu_engine* e, float* flat, int part) {
  if (e->batch.unpack.empty()) return FU_OK;
  for (const auto& j : e->batch.unpack) e->cover_now.push_back({j.dw_off, (long long)j.M * j.N * j.taps});
  const int half = fu_engine::kJobCap / 2;
  e->set_tag(0, 0, "wgrad_unpack");
  if (e->prof) e->prof_begin("tc_unpack_batched_kernel");
  const int trc = tc_flush_jobs(e->batch.unpack, part ? e->unpack_uploaded2 : e->unpack_uploaded, e->unpack_tbl + part * half,
                                half, tc_unpack_launcher(e->stream, flat), e->stream, &e->cnt, e->unpack_pin + part * half);
  if (e->prof) e->prof_end();
  e->batch.unpack.clear();
  if (trc) return e->fail(FU_ERR_CUDA, "batched weight-gradient unpack failed");
  return FU_OK;
}

int flush_deferred_sums(fu_engine* e) {
  size_t i = 0;
  while (i < e->deferred_sums.size()) {
    SumTable t;
    t.count = 0;
    int maxn = 1;
    for (; i < e->deferred_sums.size() && t.count < SumTable::kMax; ++i) {
      t.src[t.count] = e->deferred_sums[i].src; t.dst[t.count] = e->deferred_sums[i].dst;
      t.n[t.count] = e->deferred_sums[i].n; t.copies[t.count] = e->deferred_sums[i].copies;
      if (e->deferred_sums[i].n > maxn) maxn = e->deferred_sums[i].n;
      ++t.count;
    }
    e->set_tag(0, 0, "bias_grads");
    LAUNCH(e, sums_to_float_kernel, t.count, maxn < 256 ? 128 : 256, t);
  }
  e->deferred_sums.clear();
  return FU_OK;
}

// gradient of a stride-1 conv (3x3/pad1 or 1x1) w.r.t. its input
// stat / stat_done: see tc_conv_dgrad; *stat_done is set when the (tensor-core) kernel produced the sums
template <typename T>
int conv_dgrad(fu_engine* e, ConvW& cw, const View& dy, const View& dx, int B, int H, int W, int accumulate,
               double* stat = nullptr, bool* stat_done = nullptr) {
  {
    const double M = (double)B * H * W;
    e->set_tag(2.0 * M * cw.Cin * cw.Cout * cw.k * cw.k, (M * (dx.C + dy.C)) * e->esz + 2.0 * cw.Cin * cw.Cout * cw.k * cw.k,
               "conv%d_dgrad %dx%d %d->%d", cw.k, H, W, cw.Cout, cw.Cin);
  }
  const Opnd dyo = opnd(e, dy);
  if (tc_dgrad_eligible(cw.tc, dyo.p, dyo.ld, dx.p, dx.ld)) {
    { const int src = ensure_split(e, dy, (long long)B * H * W); if (src) return src; }
    if (e->prof) e->prof_begin("tc_conv_kernel");
    const int trc = tc_conv_dgrad(cw.tc, dyo.p, dyo.ld, dx.p, dx.ld, B, H, W, accumulate, e->stream, &e->cnt, stat);
    if (e->prof) e->prof_end();
    if (trc)
      return e->fail(FU_ERR_CUDA, "tensor-core dgrad launch failed: %s", tc_last_error());
    if (stat && stat_done) *stat_done = true;
    return FU_OK;
  }
  ConvCall c;
  c.x = dy; c.B = B; c.Hi = H; c.Wi = W; c.y = dx; c.Ho = H; c.Wo = W;
  c.w = cw.wp_dgrad; c.N = cw.n_dgrad; c.Npad = cw.npad_dgrad;
  c.KH = cw.k; c.stride = 1; c.pad = cw.k / 2; c.accumulate = accumulate;
  return run_igemm<T>(e, c);
}

// gradient of a stride-1 conv w.r.t. its weight, written in the torch layout (Cout,Cin,k,k)
template <typename T>
int conv_wgrad(fu_engine* e, ConvW& cw, const View& x, const View& dy, int B, int H, int W, float* dw) {
  {
    const double M = (double)B * H * W;
    e->set_tag(2.0 * M * cw.Cin * cw.Cout * cw.k * cw.k, (M * (x.C + dy.C)) * e->esz + 4.0 * cw.Cin * cw.Cout * cw.k * cw.k,
               "conv%d_wgrad %dx%d %d->%d", cw.k, H, W, cw.Cin, cw.Cout);
  }
  const Opnd xo = opnd(e, x), dyo = opnd(e, dy);
  if (tc_wgrad_eligible(cw.tc, xo.p, xo.ld, dyo.p, dyo.ld)) {
    // (twins are made on the main stream by the caller before this runs on the side stream: see prep_wgrad)
    { int src = ensure_split(e, x, (long long)B * H * W); if (!src) src = ensure_split(e, dy, (long long)B * H * W); if (src) return src; }
    if (e->prof) e->prof_begin("tc_wgrad_kernel");
    const int trc = tc_conv_wgrad(cw.tc, xo.p, xo.ld, dyo.p, dyo.ld, B, H, W, dw, e->stream, &e->cnt);
    if (e->prof) e->prof_end();
    if (trc)
      return e->fail(FU_ERR_CUDA, "tensor-core wgrad launch failed: %s", tc_last_error());
    return FU_OK;
  }
  if (cw.small_cin && (dy.ld % 4 == 0) && aligned(dy.p, 4 * sizeof(T))) {
    SmallCinWgradArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x.p; a.x_ld = x.ld; a.Cin = cw.Cin; a.dy = dy.p; a.dy_ld = dy.ld; a.Cout = cw.Cout;
    a.B = B; a.H = H; a.W = W; a.k = cw.k; a.pad = cw.k / 2; a.dw = dw;
    if constexpr (sizeof(T) == 2) {
      // bf16 storage, 32 output channels: dW = dY^T x patches on warp-level tensor-core MMAs (wgrad_cin1_mma_kernel)
      if (cw.Cin == 1 && cw.Cout == 32 && (cw.k == 3 || cw.k == 1) && W % 16 == 0 && dy.ld % 8 == 0 && aligned(dy.p, 16) &&
          (long long)B * H * W < (1ll << 30) && tc_env_int("FU_CIN1_MMA", 1)) {
        const long long tiles = (long long)B * H * (W / 16);
        const unsigned grid = (unsigned)std::min<long long>((tiles + 3) / 4, (long long)e->num_sms * 8);
        if (cw.k == 3)
          LAUNCH(e, (wgrad_cin1_mma_kernel<3>), grid, 128, reinterpret_cast<const bf16*>(x.p), x.ld, reinterpret_cast<const bf16*>(dy.p),
                 dy.ld, dw, B, H, W);
        else
          LAUNCH(e, (wgrad_cin1_mma_kernel<1>), grid, 128, reinterpret_cast<const bf16*>(x.p), x.ld, reinterpret_cast<const bf16*>(dy.p),
                 dy.ld, dw, B, H, W);
        return FU_OK;
      }
    }
    if (cw.Cin == 1) {
      const int TW = 256 / (cw.Cout / 8);
      const long long tiles = (long long)((W + TW - 1) / TW) * ((H + cin1_tile_h(cw.k) - 1) / cin1_tile_h(cw.k)) * B;
      const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)e->num_sms * (cw.k == 1 ? 3 : 2) * tc_env_int("FU_CIN1_WAVES", 1));
      const size_t sm1 = cin1_smem_bytes(cw.k, cw.Cout, true);
      if (cw.k == 3) LAUNCH_SMEM(e, (wgrad_cin1_kernel<T, 3>), grid, 256, sm1, a);
      else LAUNCH_SMEM(e, (wgrad_cin1_kernel<T, 1>), grid, 256, sm1, a);
      return FU_OK;
    }
    const int rows = 256 / (cw.Cout / 4);
    long long gx = ((long long)B * H * W + (long long)rows * 32 - 1) / ((long long)rows * 32);
    if (gx > (long long)e->num_sms * 8) gx = (long long)e->num_sms * 8;
    if (gx < 1) gx = 1;
    const int kk = cw.k * cw.k * cw.Cin;
    if (kk <= 2) LAUNCH(e, (wgrad_small_cin_kernel<T, 2>), (unsigned)gx, 256, a);
    else if (kk <= 9) LAUNCH(e, (wgrad_small_cin_kernel<T, 9>), (unsigned)gx, 256, a);
    else LAUNCH(e, (wgrad_small_cin_kernel<T, 18>), (unsigned)gx, 256, a);
    return FU_OK;
  }
  WgradCall c;
  c.big = x; c.Hb = H; c.Wb = W; c.small = dy; c.Hs = H; c.Ws = W; c.B = B;
  c.KH = cw.k; c.stride = 1; c.pad = cw.k / 2;
  c.dw = dw; c.s_tap = 1; c.s_big = (long long)cw.k * cw.k; c.s_small = (long long)cw.Cin * cw.k * cw.k;
  return run_wgrad<T>(e, c);
}

// in_sums (optional): set when blk.dstat received the per-channel sums of *d_in from the last data-gradient kernel
template <typename T>
int block_backward(fu_engine* e, Block& blk, const View& x_in, const View& g, const View* d_in, int B, int H,
                   int W, int training, float* flat, bool* in_sums = nullptr) {
  const int nd = (int)blk.convs.size();
  const bool bn = !blk.bns.empty();
  const long long P = (long long)B * H * W;
  int rc;
  if (blk.has_res) {
    // parity_tc: twins are written on the main stream BEFORE the weight gradient is forked to the side stream
    if (blk.res.tc.enabled && ((rc = ensure_split(e, x_in, P)) || (rc = ensure_split(e, g, P)))) return rc;
    {
      SideScope side(e);
      if ((rc = conv_wgrad<T>(e, blk.res, x_in, g, B, H, W, gptr(e, flat, blk.res.w_idx)))) return rc;
    }
    if (!bn && (rc = channel_sum_to<T>(e, g, P, blk.res.bsum, gptr(e, flat, blk.res.b_idx)))) return rc;
  }
  View d = g;
  for (int i = nd - 1; i >= 0; --i) {
    View r = blk.r[i];
    ConvW& cw = blk.convs[i];
    const T* dp = reinterpret_cast<const T*>(d.p);
    e->set_tag(0, 5.0 * P * blk.C * e->esz, "act_bwd %dx%d C%d", H, W, blk.C);
    if (bn) {
      BNL& b = blk.bns[i];
      BnBwdFin fin;
      fin.bstat = b.bstat; fin.gamma = tdata(e, b.i_gamma); fin.g_gamma = gptr(e, flat, b.i_gamma);
      fin.g_beta = gptr(e, flat, b.i_beta);
      fin.g_extra = (i == nd - 1 && blk.has_res) ? gptr(e, flat, blk.res.b_idx) : nullptr;
      fin.training = training;
      // one launch (reduce | grid barrier | apply) whenever the whole grid can be resident; two launches otherwise
      dim3 rg = red_grid(e, P, b.C);
      if (e->coop_blocks_per_sm < 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_act_bwd_coop_kernel<T>, 256, 0) != cudaSuccess) occ = 0;
        const int cap_env = tc_env_int("FU_BN_COOP_OCC", 8);
        e->coop_blocks_per_sm = occ < cap_env ? occ : cap_env;
      }
      // FU_BN_COOP: 0 (default) never, 1 always, 2 tensors of <= FU_BN_COOP_MB megabytes.  Measured in the graph-replayed
      // step (B=32 @192x192): 5.94 ms with two launches, 6.04 with mode 2, 6.12 with mode 1 -- the cooperative grid cannot
      // pass its barrier until every block is resident, and the persistent weight-gradient kernels on the side stream hold
      // SMs for 50-150 us at a time; per launch the fused kernel is no faster either (22 vs 12.6 + 10.7 us at 12x12).
      static const int coop_mode = tc_env_int("FU_BN_COOP", 0), coop_mb = tc_env_int("FU_BN_COOP_MB", 12);
      const bool coop_want = coop_mode == 1 || (coop_mode == 2 && (double)P * b.C * e->esz <= coop_mb * 1048576.0);
      const long long coop_cap = (long long)e->coop_blocks_per_sm * e->num_sms / rg.y;
      if (coop_want && coop_cap >= 1) {
        if ((long long)rg.x > coop_cap) rg.x = (unsigned)coop_cap;
        LAUNCH(e, (bn_act_bwd_coop_kernel<T>), rg, 256, dp, d.ld, reinterpret_cast<const T*>(r.p), r.ld,
               reinterpret_cast<T*>(blk.dy[i].p), blk.dy[i].ld, b.mean, b.invstd, fin, b.bstat, P, b.C, cw.bsum, e->coop_bar);
      } else {
        LAUNCH(e, (bn_bwd_reduce_kernel<T>), rg, 256, dp, d.ld, reinterpret_cast<const T*>(r.p),
               r.ld, b.mean, b.invstd, P, b.C, b.bstat);
        // (the apply kernel holds 3 blocks per SM: 8 x SMs blocks were 2.6 waves of it)
        LAUNCH(e, (act_bwd_kernel<T>), red_grid(e, P, b.C, kRedLanes, 3), 256, dp, d.ld, reinterpret_cast<const T*>(r.p), r.ld,
               reinterpret_cast<T*>(blk.dy[i].p), blk.dy[i].ld, 1, b.mean, b.invstd, fin, P, b.C, cw.bsum);
      }
    } else {
      BnBwdFin fin;
      memset(&fin, 0, sizeof(fin));
      LAUNCH(e, (act_bwd_kernel<T>), red_grid(e, P, blk.C, kRedLanes, 3), 256, dp, d.ld,
             reinterpret_cast<const T*>(blk.r[i].p), blk.r[i].ld, reinterpret_cast<T*>(blk.dy[i].p),
             blk.dy[i].ld, 0, nullptr, nullptr, fin, P, blk.C, cw.bsum);
    }
    e->deferred_sums.push_back({cw.bsum, gptr(e, flat, cw.b_idx), blk.C, kRedCopies});
    View conv_in = (i == 0) ? x_in : (bn ? blk.z[i - 1] : blk.r[i - 1]);
    if (cw.tc.enabled && ((rc = ensure_split(e, conv_in, P)) || (rc = ensure_split(e, blk.dy[i], P)))) return rc;
    {
      SideScope side(e);
      if ((rc = conv_wgrad<T>(e, cw, conv_in, blk.dy[i], B, H, W, gptr(e, flat, cw.w_idx)))) return rc;
    }
    if (i > 0) {
      if ((rc = conv_dgrad<T>(e, cw, blk.dy[i], blk.dz[i], B, H, W, 0))) return rc;
      d = blk.dz[i];
    } else if (d_in) {
      // the LAST writer of *d_in also sums it per channel: that is the bias gradient of the up / downsample conv
      // whose output this block consumed (no separate pass over the tensor)
      double* st = in_sums ? blk.dstat : nullptr;
      bool fused = false;
      const Opnd dy0 = opnd(e, blk.dy[0]), go = opnd(e, g);
      if (blk.has_res && tc_dgrad_eligible(cw.tc, dy0.p, dy0.ld, d_in->p, d_in->ld) &&
          tc_dgrad_can_fuse_res(cw.tc, blk.res.tc, H, W, go.p, go.ld)) {
        if ((rc = ensure_split(e, blk.dy[0], P)) || (rc = ensure_split(e, g, P))) return rc;
        // d_in = conv3x3^T(dy_0) + conv1x1^T(g) in ONE launch: the shortcut's data gradient rides along as extra,
        // centre-tap-only K chunks (no write + read-modify-write of d_in, one launch less)
        const double M = (double)B * H * W;
        e->set_tag(2.0 * M * cw.Cin * cw.Cout * 10.0, (M * (d_in->C + 2.0 * blk.C)) * e->esz, "conv3_dgrad %dx%d %d->%d",
                   H, W, cw.Cout, cw.Cin);
        if (e->prof) e->prof_begin("tc_conv_kernel");
        const int trc = tc_conv_dgrad(cw.tc, dy0.p, dy0.ld, d_in->p, d_in->plane ? d_in->planeC : d_in->ld, B, H, W, 0, e->stream, &e->cnt,
                                      st, &blk.res.tc, go.p, go.ld, d_in->plane / e->esz, d_in->planeC);
        if (e->prof) e->prof_end();
        if (trc == 0) { fused = true; if (st && in_sums) *in_sums = true; }
        else if (trc != -2) return e->fail(FU_ERR_CUDA, "fused residual dgrad launch failed: %s", tc_last_error());
        else if (e->prof) { cudaEventDestroy(e->prof_recs.back().a); cudaEventDestroy(e->prof_recs.back().b); e->prof_recs.pop_back(); }
      }
      if (!fused) {
        if (d_in->plane) return e->fail(FU_ERR_CUDA, "planar skip gradient needs the fused tensor-core data gradient (set FU_DCAT_PLANAR=0)");
        if ((rc = conv_dgrad<T>(e, cw, blk.dy[0], *d_in, B, H, W, 0, blk.has_res ? nullptr : st, in_sums))) return rc;
        if (blk.has_res && (rc = conv_dgrad<T>(e, blk.res, g, *d_in, B, H, W, 1, st, in_sums))) return rc;
      }
    }
  }
  return FU_OK;
}

template <typename T>
int backward_t(fu_engine* e, const float* d_seg, const float* d_heat, float* flat) {
  const fu_config& c = e->cfg;
  Plan& pl = e->plan;
  const int D = c.depth, esz = e->esz;
  const int B = pl.B, H = pl.H, W = pl.W;
  const long long HW = (long long)H * W, P0 = (long long)B * HW;
  const int training = e->saved_training;
  int rc;
  e->deferred_sums.clear();
  e->batch.unpack.clear();
  e->batch.flat = flat;
  struct SinkGuard { SinkGuard(TcBatch* b) { tc_batch() = b; } ~SinkGuard() { tc_batch() = nullptr; } } sink_guard(&e->batch);
  e->cover_now.clear();
  if (e->gaps_valid && e->n_gaps > 0) {
    LAUNCH(e, zero_ranges_kernel, (unsigned)(e->num_sms * 2), 256, flat, (const long long*)e->gap_tbl, e->n_gaps);
  } else if (!e->gaps_valid) {
    CUDA_TRY(e, cudaMemsetAsync(flat, 0, (size_t)e->grad_numel * sizeof(float), e->stream));
  }
  CUDA_TRY(e, cudaMemsetAsync(e->dscr_bwd, 0, e->dscr_bwd_bytes, e->stream));
  if (e->cfg.precision != FU_PRECISION_FP32 && e->wgrad_scr_bytes > 512 && !e->wgrad_scr_clean)
    CUDA_TRY(e, cudaMemsetAsync(e->wgrad_scr, 0, e->wgrad_scr_bytes, e->stream));
  e->wgrad_scr_clean = false;
  // ---- heads ----
  e->set_tag(0, 0, "heads_bwd");
  View feat = slice(pl.hcat, 0, e->Cf, esz);
  View lg = slice(pl.hcat, e->Cf, c.n_classes, esz);
  View d_feat = slice(pl.d_hcat, 0, e->Cf, esz);
  View d_lg = slice(pl.d_hcat, e->Cf, c.n_classes, esz);
  const bool fused_heads = e->Cf == 32 && c.n_classes == 7 &&
                           (c.num_lands == 0 || (c.num_lands == 14 && e->lands.size() == 2 && e->lands[0].Cout == 21));
  if (fused_heads) {
    const size_t gbytes = ((size_t)(c.num_lands + c.n_classes) * (e->Cf + c.n_classes) + 64) * sizeof(float);
    CUDA_TRY(e, cudaMemsetAsync(e->heads_gacc, 0, gbytes, e->stream));
    // resident blocks per SM: 3 with the landmark head (168 registers, 64.5 KB), 2 without (255 registers)
    const unsigned gridh = (unsigned)std::min<long long>((P0 + 255) / 256, (long long)e->num_sms * (c.num_lands == 14 ? 3 : 2));
    bool heads_done = false;
    if constexpr (sizeof(T) == 2) {
      // bf16 storage: every product of the heads' backward as warp-level tensor-core MMAs (heads_bwd_mma_kernel)
      const bool use_mma = tc_env_int("FU_HEADS_MMA", 1) != 0;
      if (use_mma && P0 < (1ll << 30) && feat.ld % 8 == 0 && d_feat.ld % 8 == 0 && reinterpret_cast<uintptr_t>(feat.p) % 16 == 0 &&
          reinterpret_cast<uintptr_t>(d_feat.p) % 16 == 0) {
        const unsigned gridm = (unsigned)std::min<long long>((P0 + 63) / 64, (long long)e->num_sms * 4);      // 4 resident blocks per SM
        HeadsLoss hl;
        memset(&hl, 0, sizeof(hl));
        if (e->lossf) {
          hl = *e->lossf;
          const LossArgs& la = e->lossf_args;
          LAUNCH(e, loss_coef_kernel, (unsigned)((B * (la.NC + la.NL) + 127) / 128), 128, (const double*)hl.sums, e->lossf_dloss,
                 e->loss_coef, B, la.NC, la.NL, la.Ht, la.Wt, la.skip_bg, la.dice_wgt, la.heat_wgt);
          hl.coef = e->loss_coef;
          if (c.num_lands == 14)
            LAUNCH(e, (heads_bwd_mma_kernel<32, 7, 21, 14, true>), gridm, 128, reinterpret_cast<const bf16*>(feat.p), feat.ld,
                   tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx), (const float*)nullptr,
                   (const float*)nullptr, reinterpret_cast<bf16*>(d_feat.p), d_feat.ld, e->heads_gacc, (int)P0, (int)HW,
                   FastDiv((int)HW), c.do_soft_max, hl);
          else
            LAUNCH(e, (heads_bwd_mma_kernel<32, 7, 1, 0, true>), gridm, 128, reinterpret_cast<const bf16*>(feat.p), feat.ld,
                   tdata(e, e->seg.w_idx), (const float*)nullptr, (const float*)nullptr, (const float*)nullptr,
                   (const float*)nullptr, reinterpret_cast<bf16*>(d_feat.p), d_feat.ld, e->heads_gacc, (int)P0, (int)HW,
                   FastDiv((int)HW), c.do_soft_max, hl);
          if (c.num_lands == 14)
            { SideScope side(e);      /* produces weight gradients only: off the data-gradient chain */ LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx),
                   gptr(e, flat, e->seg.w_idx), gptr(e, flat, e->lands[0].w_idx), gptr(e, flat, e->lands[1].w_idx), 32, 7, 21, 14); }
          else
            { SideScope side(e);      /* produces weight gradients only: off the data-gradient chain */ LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, (const float*)nullptr, (const float*)nullptr,
                   gptr(e, flat, e->seg.w_idx), (float*)nullptr, (float*)nullptr, 32, 7, 1, 0); }
        } else if (c.num_lands == 14) {
          LAUNCH(e, (heads_bwd_mma_kernel<32, 7, 21, 14>), gridm, 128, reinterpret_cast<const bf16*>(feat.p), feat.ld,
                 tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx), d_seg, d_heat,
                 reinterpret_cast<bf16*>(d_feat.p), d_feat.ld, e->heads_gacc, (int)P0, (int)HW, FastDiv((int)HW), c.do_soft_max, hl);
          { SideScope side(e);      /* produces weight gradients only: off the data-gradient chain */ LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx),
                 gptr(e, flat, e->seg.w_idx), gptr(e, flat, e->lands[0].w_idx), gptr(e, flat, e->lands[1].w_idx), 32, 7, 21, 14); }
        } else {
          LAUNCH(e, (heads_bwd_mma_kernel<32, 7, 1, 0>), gridm, 128, reinterpret_cast<const bf16*>(feat.p), feat.ld,
                 tdata(e, e->seg.w_idx), (const float*)nullptr, (const float*)nullptr, d_seg, (const float*)nullptr,
                 reinterpret_cast<bf16*>(d_feat.p), d_feat.ld, e->heads_gacc, (int)P0, (int)HW, FastDiv((int)HW), c.do_soft_max, hl);
          { SideScope side(e);      /* produces weight gradients only: off the data-gradient chain */ LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, (const float*)nullptr, (const float*)nullptr,
                 gptr(e, flat, e->seg.w_idx), (float*)nullptr, (float*)nullptr, 32, 7, 1, 0); }
        }
        heads_done = true;
      }
    }
    if (!heads_done && e->lossf) return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_backward_loss: the fused loss needs bf16 storage and the tensor-core heads");
    if (heads_done) {
    } else if (c.num_lands == 14) {
      LAUNCH_SMEM(e, (heads_bwd_fused_kernel<T, 32, 7, 21, 14>), gridh, 128, (sizeof(T) == 2 ? heads_bwd_smem_bytes_mma<32, 7, 21, 14>() : heads_bwd_smem_bytes<32, 7, 21, 14>()), reinterpret_cast<const T*>(feat.p), feat.ld,
             tdata(e, e->seg.w_idx), tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx), d_seg, d_heat,
             reinterpret_cast<T*>(d_feat.p), d_feat.ld, e->heads_gacc, B, HW, c.do_soft_max);
      LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, tdata(e, e->lands[0].w_idx), tdata(e, e->lands[1].w_idx),
             gptr(e, flat, e->seg.w_idx), gptr(e, flat, e->lands[0].w_idx), gptr(e, flat, e->lands[1].w_idx), 32, 7, 21, 14);
    } else {
      LAUNCH_SMEM(e, (heads_bwd_fused_kernel<T, 32, 7, 1, 0>), gridh, 128, (sizeof(T) == 2 ? heads_bwd_smem_bytes_mma<32, 7, 1, 0>() : heads_bwd_smem_bytes<32, 7, 1, 0>()), reinterpret_cast<const T*>(feat.p), feat.ld,
             tdata(e, e->seg.w_idx), (const float*)nullptr, (const float*)nullptr, d_seg, (const float*)nullptr,
             reinterpret_cast<T*>(d_feat.p), d_feat.ld, e->heads_gacc, B, HW, c.do_soft_max);
      LAUNCH(e, heads_bwd_finalize_kernel, 8, 256, e->heads_gacc, (const float*)nullptr, (const float*)nullptr,
             gptr(e, flat, e->seg.w_idx), (float*)nullptr, (float*)nullptr, 32, 7, 1, 0);
    }
  } else {
  if (c.num_lands > 0 && d_heat) {
    LAUNCH(e, (nchw_to_nhwc_kernel<T>), grid1d(P0, 256, e->num_sms), 256, d_heat, reinterpret_cast<T*>(pl.dheat.p),
           pl.dheat.ld, B, c.num_lands, HW);
    View dcur = pl.dheat;
    for (int k = (int)e->lands.size() - 1; k >= 0; --k) {
      ConvW& cw = e->lands[k];
      View in_k = (k == 0) ? slice(pl.hcat, 0, e->Cf + c.n_classes, esz) : pl.hmid[k - 1];
      View din_k = (k == 0) ? slice(pl.d_hcat, 0, e->Cf + c.n_classes, esz) : pl.d_hmid[k - 1];
      if ((rc = conv_wgrad<T>(e, cw, in_k, dcur, B, H, W, gptr(e, flat, cw.w_idx)))) return rc;
      if ((rc = conv_dgrad<T>(e, cw, dcur, din_k, B, H, W, 0))) return rc;
      dcur = din_k;
    }
  } else {
    CUDA_TRY(e, cudaMemsetAsync(pl.d_hcat.p, 0, (size_t)P0 * pl.d_hcat.ld * esz, e->stream));
  }
  if (d_seg) {
    LAUNCH(e, (softmax_bwd_kernel<T>), grid1d(P0, 128, e->num_sms), 128, reinterpret_cast<const T*>(lg.p), lg.ld,
           d_seg, reinterpret_cast<T*>(d_lg.p), d_lg.ld, B, c.n_classes, HW, c.do_soft_max, 1);
  }
  if ((rc = conv_wgrad<T>(e, e->seg, feat, d_lg, B, H, W, gptr(e, flat, e->seg.w_idx)))) return rc;
  if ((rc = conv_dgrad<T>(e, e->seg, d_lg, d_feat, B, H, W, 1))) return rc;
  }

  // ---- decoder ----
  View g = d_feat;
  for (int j = D - 2; j >= 0; --j) {
    const int l = D - 2 - j;
    const int h = H >> l, w = W >> l;
    bool in_sums = false;
    if ((rc = block_backward<T>(e, e->dec[j], pl.cat[l], g, &pl.d_cat[l], B, h, w, training, flat, &in_sums))) return rc;
    ConvW& up = e->upc[j];
    View d_up = slice(pl.d_cat[l], 0, e->chans[l], esz);
    View u = (l == D - 2) ? pl.bott : pl.decout[l + 1];
    View d_u = (l == D - 2) ? pl.d_bott : pl.d_decout[l + 1];
    e->set_tag(4.0 * B * (h / 2) * (w / 2) * 4.0 * up.Cin * up.Cout, 0, "up_bwd %dx%d %d->%d", h, w, up.Cin, up.Cout);
    if (in_sums) e->deferred_sums.push_back({e->dec[j].dstat, gptr(e, flat, up.b_idx), e->chans[l], 1});   // channels [0,C) of d_cat
    else if ((rc = channel_sum_to<T>(e, d_up, (long long)B * h * w, up.bsum, gptr(e, flat, up.b_idx)))) return rc;
    const Opnd uo = opnd(e, u), dupo = opnd(e, d_up);
    if (tc_up_eligible(up.tc, uo.p, uo.ld, d_u.p, d_u.ld) && tc_ptr_ok(dupo.p, dupo.ld)) {
      if ((rc = ensure_split(e, u, (long long)B * (h / 2) * (w / 2))) || (rc = ensure_split(e, d_up, (long long)B * h * w))) return rc;
      if (e->prof) e->prof_begin("tc_wgrad_kernel");
      int trc;
      {
        SideScope side(e);
        trc = tc_up_wgrad(up.tc, uo.p, uo.ld, dupo.p, dupo.ld, B, h / 2, w / 2, gptr(e, flat, up.w_idx), e->stream, &e->cnt);
      }
      if (e->prof) { e->prof_end(); e->prof_begin("tc_conv_kernel"); }
      if (!trc) trc = tc_up_dgrad(up.tc, dupo.p, dupo.ld, d_u.p, d_u.ld, B, h / 2, w / 2, e->stream, &e->cnt);
      if (e->prof) e->prof_end();
      if (trc) return e->fail(FU_ERR_CUDA, "tensor-core upconv backward failed: %s", tc_last_error());
      g = d_u;
      continue;
    }
    {
      WgradCall wc;  // dW[ci][co][ab] = sum x[n,i,j,ci] * dY[n,2i+a,2j+b,co]
      wc.big = d_up; wc.Hb = h; wc.Wb = w; wc.small = u; wc.Hs = h / 2; wc.Ws = w / 2; wc.B = B;
      wc.KH = 2; wc.stride = 2; wc.pad = 0;
      wc.dw = gptr(e, flat, up.w_idx); wc.s_tap = 1; wc.s_big = 4; wc.s_small = (long long)up.Cout * 4;
      if ((rc = run_wgrad<T>(e, wc))) return rc;
    }
    {
      ConvCall cc;  // dX[n,i,j,ci] = sum_{ab,co} dY[n,2i+a,2j+b,co] W[ci][co][ab]
      cc.x = d_up; cc.B = B; cc.Hi = h; cc.Wi = w; cc.y = d_u; cc.Ho = h / 2; cc.Wo = w / 2;
      cc.w = up.wp_dgrad; cc.N = up.n_dgrad; cc.Npad = up.npad_dgrad; cc.KH = 2; cc.stride = 2; cc.pad = 0;
      if ((rc = run_igemm<T>(e, cc))) return rc;
    }
    g = d_u;
  }
  // ---- encoder ----
  for (int l = D - 1; l >= 0; --l) {
    const int h = H >> l, w = W >> l;
    View gl = (D == 1) ? g : (l == D - 1 ? g : slice(pl.d_cat[l], e->chans[l], e->chans[l], esz));
    View x_in = (l == 0) ? pl.xin : pl.down[l];
    if (l == e->split_level() - 1) {
      // every layer at the deep levels (and the whole decoder) has its weight gradient on the side stream by now: 97 %
      // of the parameters.  Their unpack goes out behind them on that stream, beside the shallow encoder levels.
      {
        SideScope side(e);
        if ((rc = flush_unpack(e, flat, 0))) return rc;
      }
      if (e->bucket_cb && e->early_numel > 0) {
        // the early gradient bucket is final once the bias sums gathered so far are written (main stream) and the
        // unpack above has run (side stream): hand it to the caller's all-reduce on the communication stream
        if ((rc = flush_deferred_sums(e))) return rc;
        CUDA_TRY(e, cudaEventRecord(e->ev_mid_main, e->stream));
        CUDA_TRY(e, cudaStreamWaitEvent(e->comm_stream, e->ev_mid_main, 0));
        if (e->side_used) {
          CUDA_TRY(e, cudaEventRecord(e->ev_mid_side, e->side));
          CUDA_TRY(e, cudaStreamWaitEvent(e->comm_stream, e->ev_mid_side, 0));
        }
        e->bucket_cb(e->bucket_user, 0, 0, e->early_numel);
      }
    }
    bool in_sums = false;
    if ((rc = block_backward<T>(e, e->enc[l], x_in, gl, l == 0 ? nullptr : &pl.d_down[l], B, h, w, training, flat,
                                (l > 0 && !c.max_pool) ? &in_sums : nullptr)))
      return rc;
    if (l > 0) {
      View src = slice(pl.cat[l - 1], e->chans[l - 1], e->chans[l - 1], esz);     // encoder output of level l-1
      View d_src = slice(pl.d_cat[l - 1], e->chans[l - 1], e->chans[l - 1], esz);  // already holds the skip gradient
      e->set_tag(4.0 * B * h * w * 4.0 * src.C * src.C, 0, "down_bwd %dx%d C%d", 2 * h, 2 * w, src.C);
      if (c.max_pool) {
        LAUNCH(e, (maxpool_bwd_kernel<T>), grid1d((long long)B * h * w * (src.C / 4), 256, e->num_sms), 256,
               reinterpret_cast<const T*>(src.p), src.ld, reinterpret_cast<const T*>(pl.d_down[l].p),
               pl.d_down[l].ld, reinterpret_cast<T*>(d_src.p), d_src.ld, B, h, w, src.C, 1);
      } else {
        ConvW& cw = e->downc[l - 1];
        if (in_sums) e->deferred_sums.push_back({e->enc[l].dstat, gptr(e, flat, cw.b_idx), cw.Cout, 1});
        else if ((rc = channel_sum_to<T>(e, pl.d_down[l], (long long)B * h * w, cw.bsum, gptr(e, flat, cw.b_idx)))) return rc;
        const Opnd so = opnd(e, src), ddo = opnd(e, pl.d_down[l]);
        if (tc_down_eligible(cw.tc, so.p, so.ld, d_src.p, d_src.ld) && tc_ptr_ok(ddo.p, ddo.ld)) {
          if ((rc = ensure_split(e, src, 4ll * B * h * w)) || (rc = ensure_split(e, pl.d_down[l], (long long)B * h * w))) return rc;
          if (e->prof) e->prof_begin("tc_wgrad_kernel");
          int trc;
          {
            SideScope side(e);
            trc = tc_down_wgrad(cw.tc, so.p, so.ld, ddo.p, ddo.ld, B, 2 * h, 2 * w,
                                gptr(e, flat, cw.w_idx), e->stream, &e->cnt);
          }
          if (e->prof) { e->prof_end(); e->prof_begin("tc_conv_kernel"); }
          if (!trc) trc = tc_down_dgrad(cw.tc, ddo.p, ddo.ld, d_src.p, d_src.ld, B, 2 * h, 2 * w, 1,
                                        e->stream, &e->cnt);
          if (e->prof) e->prof_end();
          if (trc) return e->fail(FU_ERR_CUDA, "tensor-core downsample backward failed: %s", tc_last_error());
          continue;
        }
        WgradCall wc;  // dW[co][ci][ab] = sum x[n,2i+a,2j+b,ci] * dY[n,i,j,co]
        wc.big = src; wc.Hb = 2 * h; wc.Wb = 2 * w; wc.small = pl.d_down[l]; wc.Hs = h; wc.Ws = w; wc.B = B;
        wc.KH = 2; wc.stride = 2; wc.pad = 0;
        wc.dw = gptr(e, flat, cw.w_idx); wc.s_tap = 1; wc.s_big = 4; wc.s_small = (long long)cw.Cin * 4;
        if ((rc = run_wgrad<T>(e, wc))) return rc;
        ConvCall cc;  // dX[n,2i+a,2j+b,ci] += sum_co dY[n,i,j,co] W[co][ci][ab]
        cc.x = pl.d_down[l]; cc.B = B; cc.Hi = h; cc.Wi = w; cc.y = d_src; cc.Ho = h; cc.Wo = w;
        cc.w = cw.wp_dgrad; cc.N = cw.n_dgrad; cc.Npad = cw.npad_dgrad; cc.KH = 1; cc.stride = 1; cc.pad = 0;
        cc.shuffle = 1; cc.accumulate = 1;
        if ((rc = run_igemm<T>(e, cc))) return rc;
      }
    }
  }
  side_join(e);                  // every weight gradient has been accumulated before anything reads it
  if ((rc = flush_unpack(e, flat, 1))) return rc;     // the shallow layers' gradients (the deep ones went out mid-way)
  if ((rc = flush_deferred_sums(e))) return rc;
  e->wgrad_scr_clean = true;     // every accumulator that was written this step has been read and cleared
  return update_flat_gaps(e);
}

int validate(const fu_config* c, std::string& why) {
  char buf[256];
  auto bad = [&](const char* m) { why = m; return FU_ERR_INVALID_CONFIG; };
  if (!c->padding) return bad("padding=False is not supported (it also crashes the reference with do_res=True, unet.py:227-231); use padding=True");
  if (!c->pad_mode_zeros) return bad("pad_mode must be 'zeros'");
  if (!c->up_mode_upconv) return bad("up_mode='upsample' is not supported; use 'upconv'");
  if (c->lands_block_depth != 0) return bad("lands_block_depth > 0 is not supported");
  if (c->depth < 1 || c->depth > 8) return bad("depth must be in [1,8]");
  if (c->wf < 2 || c->wf + c->depth - 1 > 12) return bad("wf must be >= 2 and 2^(wf+depth-1) <= 4096");
  if (c->precision == FU_PRECISION_BF16 && c->wf < 3) return bad("throughput mode (bf16) needs wf >= 3 (16-byte channel vectors); use precision='fp32'");
  if (c->in_channels < 1 || c->in_channels > 64) return bad("in_channels must be in [1,64]");
  if (c->n_classes < 1 || c->n_classes > kMaxClasses) return bad("n_classes must be in [1,32]");
  if (c->num_lands < 0 || c->num_lands > 256) return bad("num_lands must be in [0,256]");
  if (c->block_depth < 1 || c->block_depth > 8) return bad("block_depth must be in [1,8]");
  if (c->num_lands > 0 && (c->lands_num_1x1 < 1 || c->lands_num_1x1 > 8)) return bad("lands_num_1x1 must be in [1,8]");
  if (c->precision != FU_PRECISION_FP32 && c->precision != FU_PRECISION_BF16 && c->precision != FU_PRECISION_FP32_TC)
    return bad("precision must be FU_PRECISION_FP32, FU_PRECISION_BF16 or FU_PRECISION_FP32_TC");
  (void)buf;
  return FU_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int fu_engine_create(const fu_config* cfg, int device, fu_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return FU_ERR_ARG; }
  *out = nullptr;
  std::string why;
  int rc = validate(cfg, why);
  if (rc) { g_create_error = why; return rc; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || device < 0 || device >= ndev) {
    g_create_error = std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                     (ce != cudaSuccess ? cudaGetErrorString(ce) : "ordinal out of range") +
                     " (this engine has no CPU fallback)";
    return FU_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library is built for sm_100a (B200) only";
    return FU_ERR_CUDA;
  }
  fu_engine* e = new fu_engine();
  e->cfg = *cfg;
  e->device = device;
  e->esz = cfg->precision == FU_PRECISION_BF16 ? 2 : 4;
  e->split = cfg->precision == FU_PRECISION_FP32_TC;
  e->num_sms = prop.multiProcessorCount;
  memset(&e->cnt, 0, sizeof(e->cnt));
  DeviceScope dev_scope(device);
  build_schema(e);
  rc = alloc_persistent(e);
  if (rc) { g_create_error = e->err; delete e; return rc; }
  *out = e;
  return FU_OK;
}

void fu_engine_destroy(fu_engine* e) {
  if (!e) return;
  DeviceScope dev_scope(e->device);
  cudaDeviceSynchronize();
  if (e->plan.arena) cudaFree(e->plan.arena);
  if (e->plan.twin) cudaFree(e->plan.twin);
  if (e->wmem) cudaFree(e->wmem);
  if (e->dscr_fwd) cudaFree(e->dscr_fwd);
  if (e->dscr_bwd) cudaFree(e->dscr_bwd);
  if (e->wgrad_scr) cudaFree(e->wgrad_scr);
  if (e->gap_tbl) cudaFree(e->gap_tbl);
  if (e->loss_coef) cudaFree(e->loss_coef);
  if (e->pack_tbl) cudaFree(e->pack_tbl);
  if (e->unpack_tbl) cudaFree(e->unpack_tbl);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_mid_main) cudaEventDestroy(e->ev_mid_main);
  if (e->ev_mid_side) cudaEventDestroy(e->ev_mid_side);
  if (e->pack_pin) cudaFreeHost(e->pack_pin);
  if (e->unpack_pin) cudaFreeHost(e->unpack_pin);
  delete e;
}

const char* fu_last_error(const fu_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int fu_num_tensors(const fu_engine* e) { return e ? (int)e->tensors.size() : FU_ERR_ARG; }

int fu_tensor_get_info(const fu_engine* e, int index, fu_tensor_info* out) {
  if (!e || !out || index < 0 || index >= (int)e->tensors.size()) return FU_ERR_ARG;
  *out = e->tensors[index].info;
  return FU_OK;
}

int64_t fu_grad_numel(const fu_engine* e) { return e ? e->grad_numel : FU_ERR_ARG; }

int fu_bind_tensors(fu_engine* e, void* const* data_ptrs, int n) {
  if (!e || !data_ptrs) return FU_ERR_ARG;
  if (n != (int)e->tensors.size()) return e->fail(FU_ERR_ARG, "fu_bind_tensors: expected %d pointers, got %d", (int)e->tensors.size(), n);
  bool changed = false;
  for (int i = 0; i < n; ++i) {
    if (!data_ptrs[i]) return e->fail(FU_ERR_ARG, "fu_bind_tensors: null pointer for %s", e->tensors[i].info.name);
    if (!aligned(data_ptrs[i], e->tensors[i].info.dtype == FU_DTYPE_I64 ? 8 : 4))
      return e->fail(FU_ERR_ARG, "fu_bind_tensors: misaligned pointer for %s", e->tensors[i].info.name);
    if (e->tensors[i].data != data_ptrs[i]) changed = true;
    e->tensors[i].data = data_ptrs[i];
  }
  if (changed) e->packed_once = false;
  e->bound = true;
  return FU_OK;
}

int fu_forward(fu_engine* e, const float* x, int B, int H, int W, int training, int save, int64_t weights_version,
               float* seg, float* logits, float* heat, void* stream) {
  if (!e) return FU_ERR_ARG;
  if (!e->bound) return e->fail(FU_ERR_NOT_BOUND, "fu_forward before fu_bind_tensors");
  if (!x || !seg) return e->fail(FU_ERR_ARG, "fu_forward: x and seg must be non-null");
  if (e->cfg.num_lands > 0 && !heat) return e->fail(FU_ERR_ARG, "fu_forward: heat must be non-null when num_lands > 0");
  DeviceScope dev_scope(e->device);
  e->stream = reinterpret_cast<cudaStream_t>(stream);
  const int64_t l0 = e->cnt.kernel_launches;
  int rc = ensure_plan(e, B, H, W);
  if (rc) return rc;
  if (!e->packed_once || e->packed_version != weights_version) {
    if ((rc = pack_all(e, training != 0))) { side_join(e); return rc; }
    e->packed_once = true;
    e->packed_version = weights_version;
  }
  e->saved = false;
  e->fresh_fwd.clear(); e->fresh_bwd.clear(); e->in_backward = false;
  // inference fast path: eval-mode statistics and no backward will follow (FU_EVAL_FAST=0 keeps the two-pass form)
  static const bool eval_fast_on = tc_env_int("FU_EVAL_FAST", 1) != 0;
  const bool eval_fast = eval_fast_on && !training && !save && e->cfg.precision != FU_PRECISION_FP32;
  if (e->cfg.precision == FU_PRECISION_BF16) rc = forward_t<bf16>(e, x, B, H, W, training, seg, logits, heat, eval_fast);
  else rc = forward_t<float>(e, x, B, H, W, training, seg, logits, heat, eval_fast);
  side_join(e);
  if (rc) return rc;
  e->saved = save != 0;
  e->saved_training = training;
  e->cnt.forward_calls++;
  e->cnt.last_fwd_launches = e->cnt.kernel_launches - l0;
  return FU_OK;
}

int fu_backward(fu_engine* e, const float* d_seg, const float* d_heat, float* flat_grads, void* stream) {
  if (!e) return FU_ERR_ARG;
  if (!e->saved) return e->fail(FU_ERR_STATE, "fu_backward without a saved forward (call fu_forward with save=1 first)");
  if (!flat_grads) return e->fail(FU_ERR_ARG, "fu_backward: flat_grads must be non-null");
  if (!aligned(flat_grads, 16)) return e->fail(FU_ERR_ARG, "fu_backward: flat_grads must be 16-byte aligned");
  DeviceScope dev_scope(e->device);
  e->stream = reinterpret_cast<cudaStream_t>(stream);
  const int64_t l0 = e->cnt.kernel_launches;
  int rc;
  e->fresh_bwd.clear(); e->in_backward = true;
  if (e->cfg.precision == FU_PRECISION_BF16) rc = backward_t<bf16>(e, d_seg, d_heat, flat_grads);
  else rc = backward_t<float>(e, d_seg, d_heat, flat_grads);
  e->in_backward = false;
  side_join(e);                  // (already joined on the normal path; an error return may have left work on the side stream)
  if (rc) return rc;
  e->cnt.backward_calls++;
  e->cnt.last_bwd_launches = e->cnt.kernel_launches - l0;
  return FU_OK;
}

int fu_set_bucket_callback(fu_engine* e, fu_bucket_callback cb, void* user, void* comm_stream) {
  if (!e) return FU_ERR_ARG;
  if (cb && !comm_stream) return e->fail(FU_ERR_ARG, "fu_set_bucket_callback: a communication stream is required");
  DeviceScope dev_scope(e->device);
  if (cb && !e->ev_mid_main) {
    CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_mid_main, cudaEventDisableTiming));
    CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_mid_side, cudaEventDisableTiming));
  }
  e->bucket_cb = cb; e->bucket_user = user; e->comm_stream = reinterpret_cast<cudaStream_t>(comm_stream);
  return FU_OK;
}

int64_t fu_early_grad_numel(const fu_engine* e) { return e ? e->early_numel : FU_ERR_ARG; }

int fu_get_counters(const fu_engine* e, fu_counters* out) {
  if (!e || !out) return FU_ERR_ARG;
  *out = e->cnt;
  return FU_OK;
}

int fu_debug_copy(fu_engine* e, const char* name, float* dst, int64_t capacity, int32_t* shape4) {
  if (!e || !name || !shape4) return FU_ERR_ARG;
  if (!e->plan.valid) return e->fail(FU_ERR_STATE, "fu_debug_copy: no forward has run yet");
  Plan& pl = e->plan;
  const int D = e->cfg.depth;
  View v; int lvl = -1;
  std::string nm(name);
  auto level_of_dec = [&](int j) { return D - 2 - j; };
  int idx = -1, sub = -1;
  char kind[16] = {0};
  if (sscanf(name, "enc%d.%15[a-z]%d", &idx, kind, &sub) == 3 && idx >= 0 && idx < D) {
    Block& b = e->enc[idx]; lvl = idx;
    std::string k(kind);
    if (sub < 0 || sub >= (int)b.convs.size()) return e->fail(FU_ERR_ARG, "fu_debug_copy: bad index in %s", name);
    v = k == "r" ? b.r[sub] : k == "z" ? b.z[sub] : k == "dy" ? b.dy[sub] : k == "dz" ? b.dz[sub] : View();
  } else if (sscanf(name, "dec%d.%15[a-z]%d", &idx, kind, &sub) == 3 && idx >= 0 && idx < D - 1) {
    Block& b = e->dec[idx]; lvl = level_of_dec(idx);
    std::string k(kind);
    if (sub < 0 || sub >= (int)b.convs.size()) return e->fail(FU_ERR_ARG, "fu_debug_copy: bad index in %s", name);
    v = k == "r" ? b.r[sub] : k == "z" ? b.z[sub] : k == "dy" ? b.dy[sub] : k == "dz" ? b.dz[sub] : View();
  } else if (sscanf(name, "d_cat%d", &idx) == 1 && idx >= 0 && idx < D - 1) { v = pl.d_cat[idx]; lvl = idx; }
  else if (sscanf(name, "cat%d", &idx) == 1 && idx >= 0 && idx < D - 1) { v = pl.cat[idx]; lvl = idx; }
  else if (sscanf(name, "d_down%d", &idx) == 1 && idx >= 1 && idx < D) { v = pl.d_down[idx]; lvl = idx; }
  else if (sscanf(name, "down%d", &idx) == 1 && idx >= 1 && idx < D) { v = pl.down[idx]; lvl = idx; }
  else if (sscanf(name, "d_decout%d", &idx) == 1 && idx >= 0 && idx < D - 1) { v = pl.d_decout[idx]; lvl = idx; }
  else if (sscanf(name, "decout%d", &idx) == 1 && idx >= 0 && idx < D - 1) { v = pl.decout[idx]; lvl = idx; }
  else if (nm == "bott") { v = pl.bott; lvl = D - 1; }
  else if (nm == "d_bott") { v = pl.d_bott; lvl = D - 1; }
  else if (nm == "hcat") { v = pl.hcat; lvl = 0; }
  else if (nm == "d_hcat") { v = pl.d_hcat; lvl = 0; }
  if (!v.p || lvl < 0) return e->fail(FU_ERR_ARG, "fu_debug_copy: unknown or unmaterialised tensor '%s'", name);
  if (v.plane) return e->fail(FU_ERR_ARG, "fu_debug_copy: '%s' is stored as planes (FU_DCAT_PLANAR=0 for the interleaved layout)", name);
  const int h = pl.H >> lvl, w = pl.W >> lvl;
  shape4[0] = pl.B; shape4[1] = v.C; shape4[2] = h; shape4[3] = w;
  const int64_t need = (int64_t)pl.B * v.C * h * w;
  if (!dst) return FU_OK;
  if (capacity < need) return e->fail(FU_ERR_ARG, "fu_debug_copy: buffer too small");
  const long long HW = (long long)h * w;
  if (e->esz == 2)
    LAUNCH(e, (nhwc_to_nchw_kernel<bf16>), grid1d(pl.B * HW, 256, e->num_sms), 256, reinterpret_cast<const bf16*>(v.p), v.ld, dst, pl.B, v.C, HW);
  else
    LAUNCH(e, (nhwc_to_nchw_kernel<float>), grid1d(pl.B * HW, 256, e->num_sms), 256, reinterpret_cast<const float*>(v.p), v.ld, dst, pl.B, v.C, HW);
  return FU_OK;
}

int fu_profile_enable(fu_engine* e, int on) {
  if (!e) return FU_ERR_ARG;
  for (auto& r : e->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof_recs.clear();
  e->prof = on != 0;
  e->tag = "other"; e->tag_flops = 0; e->tag_bytes = 0;
  return FU_OK;
}

int64_t fu_profile_report(fu_engine* e, char* buf, int64_t cap) {
  if (!e) return FU_ERR_ARG;
  DeviceScope dev_scope(e->device);
  if (!e->prof_recs.empty()) cudaEventSynchronize(e->prof_recs.back().b);
  struct Agg { std::string tag, kern; int n; double ms, flops, bytes; };
  std::vector<Agg> aggs;
  for (auto& r : e->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
    std::string k = r.kern;
    const size_t lt = k.find('<');
    if (lt != std::string::npos) k = k.substr(0, lt);
    while (!k.empty() && (k[0] == '(' || k[0] == ' ')) k.erase(0, 1);
    Agg* hit = nullptr;
    for (auto& a : aggs) if (a.tag == r.tag && a.kern == k) { hit = &a; break; }
    if (!hit) { aggs.push_back({r.tag, k, 0, 0, 0, 0}); hit = &aggs.back(); }
    hit->n++; hit->ms += ms; hit->flops += r.flops; hit->bytes += r.bytes;
  }
  std::string out;
  char line[512];
  for (auto& a : aggs) {
    snprintf(line, sizeof(line), "{\"tag\": \"%s\", \"kernel\": \"%s\", \"launches\": %d, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}\n",
             a.tag.c_str(), a.kern.c_str(), a.n, a.ms, a.flops, a.bytes);
    out += line;
  }
  if (buf && cap > 0) {
    const int64_t n = (int64_t)out.size() < cap - 1 ? (int64_t)out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

// ---- fused loss (kernels_loss.cuh) ----
static int loss_args(const fu_loss_desc* d, LossArgs& a) {
  if (!d || !d->seg || !d->mask || d->B < 1 || d->n_classes < 1 || d->Ht < 1 || d->Wt < 1 || d->num_lands < 0) {
    g_create_error = "fu_loss: bad descriptor";
    return FU_ERR_ARG;
  }
  if ((d->heat == nullptr) != (d->num_lands == 0) || (d->heat && !d->heat_t)) {
    g_create_error = "fu_loss: heat / heat_t / num_lands disagree";
    return FU_ERR_ARG;
  }
  if ((long long)d->Ht * d->Wt < 2) { g_create_error = "fu_loss: ncc needs at least 2 pixels (ncc.py:14)"; return FU_ERR_ARG; }
  memset(&a, 0, sizeof(a));
  a.seg = d->seg; a.seg_sb = d->seg_stride[0]; a.seg_sc = d->seg_stride[1]; a.seg_sr = (int)d->seg_stride[2];
  a.mask = d->mask; a.mask_sb = d->mask_stride[0]; a.mask_sc = d->mask_stride[1]; a.mask_sr = (int)d->mask_stride[2];
  a.heat = d->heat; a.heat_sb = d->heat_stride[0]; a.heat_sc = d->heat_stride[1]; a.heat_sr = (int)d->heat_stride[2];
  a.heat_t = d->heat_t; a.heat_t_sb = d->heat_t_stride[0]; a.heat_t_sc = d->heat_t_stride[1]; a.heat_t_sr = (int)d->heat_t_stride[2];
  a.B = d->B; a.NC = d->n_classes; a.NL = d->num_lands; a.Ht = d->Ht; a.Wt = d->Wt;
  a.skip_bg = d->skip_bg ? 1 : 0; a.dice_wgt = d->dice_wgt; a.heat_wgt = d->heat_wgt;
  if (a.skip_bg && a.NC < 2) { g_create_error = "fu_loss: skip_bg needs at least 2 classes"; return FU_ERR_ARG; }
  return FU_OK;
}

int64_t fu_loss_workspace_doubles(int B, int n_classes, int num_lands) {
  return (int64_t)B * (3 * (int64_t)n_classes + 5 * (int64_t)num_lands);
}

int fu_loss_forward(const fu_loss_desc* d, double* sums, float* loss_out, void* stream) {
  LossArgs a;
  int rc = loss_args(d, a);
  if (rc) return rc;
  if (!sums || !loss_out) { g_create_error = "fu_loss_forward: null workspace / output"; return FU_ERR_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  a.sums = sums;
  const size_t ws = (size_t)fu_loss_workspace_doubles(a.B, a.NC, a.NL) * sizeof(double);
  if (cudaMemsetAsync(sums, 0, ws, st) != cudaSuccess) { g_create_error = "fu_loss_forward: memset failed"; return FU_ERR_CUDA; }
  a.rows = loss_rows_per_block(a.Ht, (long long)a.B * (a.NC + a.NL));
  const dim3 grid((unsigned)((a.Ht + a.rows - 1) / a.rows), (unsigned)(a.B * (a.NC + a.NL)));
  {
    auto al8 = [](const void* p_) { return reinterpret_cast<uintptr_t>(p_) % 8 == 0; };
    auto even = [](long long v) { return (v & 1) == 0; };
    const bool vec2 = even(a.Wt) && al8(a.seg) && al8(a.mask) && even(a.seg_sb) && even(a.seg_sc) && even(a.seg_sr) && even(a.mask_sb) &&
                      even(a.mask_sc) && even(a.mask_sr) &&
                      (!a.heat || (al8(a.heat) && al8(a.heat_t) && even(a.heat_sb) && even(a.heat_sc) && even(a.heat_sr) &&
                                   even(a.heat_t_sb) && even(a.heat_t_sc) && even(a.heat_t_sr))) && tc_env_int("FU_LOSS_VEC", 1) != 0;
    if (vec2) loss_sums_kernel<true><<<grid, 256, 0, st>>>(a);
    else loss_sums_kernel<false><<<grid, 256, 0, st>>>(a);
  }
  loss_finalize_kernel<<<1, 256, 0, st>>>(a, loss_out);
  cudaError_t ce = cudaPeekAtLastError();
  if (ce != cudaSuccess) { g_create_error = std::string("fu_loss_forward: ") + cudaGetErrorString(ce); return FU_ERR_CUDA; }
  return FU_OK;
}

int fu_loss_backward(const fu_loss_desc* d, const double* sums, const float* dloss, int H, int W, int r0, int c0,
                     float* d_seg, float* d_heat, void* stream) {
  LossBwdArgs q;
  int rc = loss_args(d, q.a);
  if (rc) return rc;
  if (!sums || !dloss || !d_seg || (q.a.NL > 0 && !d_heat)) { g_create_error = "fu_loss_backward: null argument"; return FU_ERR_ARG; }
  if (r0 < 0 || c0 < 0 || r0 + q.a.Ht > H || c0 + q.a.Wt > W) { g_create_error = "fu_loss_backward: window outside the output"; return FU_ERR_ARG; }
  q.a.sums = const_cast<double*>(sums);
  q.dloss = dloss; q.d_seg = d_seg; q.d_heat = d_heat; q.H = H; q.W = W; q.r0 = r0; q.c0 = c0;
  q.a.rows = loss_rows_per_block(H, (long long)q.a.B * (q.a.NC + q.a.NL));
  const dim3 grid((unsigned)((H + q.a.rows - 1) / q.a.rows), (unsigned)(q.a.B * (q.a.NC + q.a.NL)));
  // 16-byte path: full rows of 4-column groups, aligned plane origins (the prediction pointers address the window origin)
  auto al16 = [](const void* p_) { return reinterpret_cast<uintptr_t>(p_) % 16 == 0; };
  const float* seg0 = q.a.seg - ((long long)r0 * q.a.seg_sr + c0);
  const float* heat0 = q.a.heat ? q.a.heat - ((long long)r0 * q.a.heat_sr + c0) : nullptr;
  const bool vec4 = W % 4 == 0 && q.a.seg_sr % 4 == 0 && q.a.seg_sb % 4 == 0 && q.a.seg_sc % 4 == 0 && al16(seg0) && al16(d_seg) &&
                    (!heat0 || (q.a.heat_sr % 4 == 0 && q.a.heat_sb % 4 == 0 && q.a.heat_sc % 4 == 0 && al16(heat0) && al16(d_heat))) &&
                    tc_env_int("FU_LOSS_VEC", 1) != 0;
  if (vec4) loss_backward_kernel<true><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q);
  else loss_backward_kernel<false><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q);
  cudaError_t ce = cudaPeekAtLastError();
  if (ce != cudaSuccess) { g_create_error = std::string("fu_loss_backward: ") + cudaGetErrorString(ce); return FU_ERR_CUDA; }
  return FU_OK;
}

// ---- the loss inside the heads kernels (SURVEY 8f row 1 as specified) ----
static int heads_loss_setup(fu_engine* e, const fu_loss_desc* d, int B, int H, int W, int r0, int c0, HeadsLoss& hl, const char* who) {
  LossArgs& a = e->lossf_args;
  fu_loss_desc dd = *d;
  // the prediction pointers of the descriptor are not used here (the predictions never leave the head kernels); loss_args
  // only validates their presence
  static const float dummy = 0.f;
  dd.seg = &dummy; if (dd.num_lands > 0) dd.heat = &dummy; else dd.heat = nullptr;
  int rc = loss_args(&dd, a);
  if (rc) return e->fail(rc, "%s: %s", who, g_create_error.c_str());
  const fu_config& c = e->cfg;
  if (c.precision != FU_PRECISION_BF16 || e->Cf != 32 || c.n_classes != 7 ||
      !(c.num_lands == 0 || (c.num_lands == 14 && e->lands.size() == 2 && e->lands[0].Cout == 21)))
    return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "%s: bf16 storage and the paper heads (32 features, 7 classes, 0 or 14 landmarks) only", who);
  if (a.B != B || a.NC != c.n_classes || a.NL != c.num_lands)
    return e->fail(FU_ERR_ARG, "%s: the loss descriptor's batch / class / landmark counts do not match the network", who);
  if (((long long)H * W) % 16 != 0)
    return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "%s: H * W must be a multiple of 16", who);
  if (r0 < 0 || c0 < 0 || r0 + a.Ht > H || c0 + a.Wt > W) return e->fail(FU_ERR_ARG, "%s: window outside the output", who);
  memset(&hl, 0, sizeof(hl));
  hl.mask = a.mask; hl.mask_sb = a.mask_sb; hl.mask_sc = a.mask_sc; hl.mask_sr = a.mask_sr;
  hl.heat_t = a.heat_t; hl.heat_t_sb = a.heat_t_sb; hl.heat_t_sc = a.heat_t_sc; hl.heat_t_sr = a.heat_t_sr;
  hl.Ht = a.Ht; hl.Wt = a.Wt; hl.r0 = r0; hl.c0 = c0; hl.W = W; hl.fd_w = FastDiv(W);
  return FU_OK;
}

int fu_forward_loss(fu_engine* e, const float* x, int B, int H, int W, int64_t weights_version, const fu_loss_desc* d, int r0, int c0,
                    double* sums, float* loss_out, float* seg, float* heat, void* stream) {
  if (!e) return FU_ERR_ARG;
  if (!d || !sums || !loss_out) return e->fail(FU_ERR_ARG, "fu_forward_loss: null descriptor / workspace / output");
  if (e->cfg.num_lands > 0 && !heat) return e->fail(FU_ERR_ARG, "fu_forward_loss: heat must be non-null when num_lands > 0 (fu_backward_loss reads it)");
  if (!tc_env_int("FU_HEADS_MMA", 1)) return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_forward_loss: needs the tensor-core heads (FU_HEADS_MMA=0 is set)");
  HeadsLoss hl;
  int rc = heads_loss_setup(e, d, B, H, W, r0, c0, hl, "fu_forward_loss");
  if (rc) return rc;
  hl.sums = sums;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    DeviceScope dev_scope(e->device);
    const size_t ws = (size_t)fu_loss_workspace_doubles(B, e->cfg.n_classes, e->cfg.num_lands) * sizeof(double);
    CUDA_TRY(e, cudaMemsetAsync(sums, 0, ws, st));
  }
  e->lossf = &hl; e->lossf_noseg = seg == nullptr;
  // (seg may be NULL: the class probabilities then never leave the head kernel; fu_forward itself insists on a pointer,
  //  which the head launch drops again -- lossf_noseg)
  rc = fu_forward(e, x, B, H, W, 1, 1, weights_version, seg ? seg : loss_out, nullptr, heat, stream);
  e->lossf = nullptr; e->lossf_noseg = false;
  if (rc) return rc;
  {
    DeviceScope dev_scope(e->device);
    LossArgs a = e->lossf_args;
    a.sums = sums;
    loss_finalize_kernel<<<1, 256, 0, st>>>(a, loss_out);
    e->cnt.kernel_launches++;
    cudaError_t ce = cudaPeekAtLastError();
    if (ce != cudaSuccess) return e->fail(FU_ERR_CUDA, "fu_forward_loss: %s", cudaGetErrorString(ce));
  }
  return FU_OK;
}

int fu_backward_loss(fu_engine* e, const fu_loss_desc* d, int r0, int c0, const double* sums, const float* dloss, const float* heat,
                     float* flat_grads, void* stream) {
  if (!e) return FU_ERR_ARG;
  if (!e->saved) return e->fail(FU_ERR_STATE, "fu_backward_loss without a saved forward (call fu_forward_loss first)");
  if (!d || !sums || !dloss) return e->fail(FU_ERR_ARG, "fu_backward_loss: null descriptor / workspace / upstream gradient");
  if (e->cfg.num_lands > 0 && !heat) return e->fail(FU_ERR_ARG, "fu_backward_loss: heat (the forward's heat-map output) must be non-null");
  if (!tc_env_int("FU_HEADS_MMA", 1)) return e->fail(FU_ERR_UNSUPPORTED_SHAPE, "fu_backward_loss: needs the tensor-core heads (FU_HEADS_MMA=0 is set)");
  const int B = e->plan.B, H = e->plan.H, W = e->plan.W;
  HeadsLoss hl;
  int rc = heads_loss_setup(e, d, B, H, W, r0, c0, hl, "fu_backward_loss");
  if (rc) return rc;
  hl.sums = const_cast<double*>(sums);
  hl.heat = heat;
  const size_t need = (size_t)B * (2 * e->cfg.n_classes + 3 * e->cfg.num_lands);
  if (need > e->loss_coef_floats) {
    DeviceScope dev_scope(e->device);
    // (grows only when the batch size does: never inside a captured step that was warmed up at its own shapes)
    if (e->loss_coef) cudaFree(e->loss_coef);
    e->loss_coef = nullptr; e->loss_coef_floats = 0;
    CUDA_TRY(e, cudaMalloc(&e->loss_coef, need * sizeof(float)));
    e->loss_coef_floats = need;
  }
  e->lossf = &hl; e->lossf_dloss = dloss;
  rc = fu_backward(e, nullptr, nullptr, flat_grads, stream);
  e->lossf = nullptr; e->lossf_dloss = nullptr;
  return rc;
}

const char* fu_build_info(void) {
  return "fluoro_unet;arch=sm_100a;tcgen05=" FU_TC_BUILD ";cuda=" FU_STR(CUDART_VERSION);
}

}  // extern "C"

#include "io_api.inl"
#include "test_hooks.inl"
