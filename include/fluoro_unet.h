/*
 * fluoro_unet.h -- C ABI of the B200-native U-Net forward/backward engine.
 *
 * The reference (rg2/DeepFluoroLabeling-IPCAI2020) has no FFI of its own: its hot
 * path is the Python nn.Module `UNet` in train_test_code/unet.py.  This header
 * is the boundary a binding of that module talks to; every entry point names
 * the reference interface it stands in for (file:line into the reference).
 *
 * Conventions
 *   - plain C types only; all tensor pointers are DEVICE pointers on the
 *     engine's device; `stream` is a cudaStream_t passed as void*.
 *   - every call returns 0 (FU_OK) or a negative status; fu_last_error() gives
 *     the message.  Nothing here falls back to the CPU.
 *   - boundary tensors are fp32 NCHW exactly as the reference passes them
 *     (unet.py:161, :190-193); the engine's internal layout (NHWC, fp32 or
 *     bf16) never shows through.
 *   - ownership: the caller (PyTorch) owns parameters, BN buffers, inputs,
 *     outputs and the flat gradient buffer; the engine owns only its
 *     activation/workspace arena and its packed weight copies.
 *   - threading: one caller thread per engine (train.py:376-443 drives the net
 *     from one thread); all work is enqueued on the caller's stream, no host
 *     synchronisation inside fu_forward / fu_backward.
 */
#ifndef FLUORO_UNET_H_
#define FLUORO_UNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FU_OK 0
#define FU_ERR_INVALID_CONFIG (-1)   /* constructor argument the engine rejects */
#define FU_ERR_UNSUPPORTED_SHAPE (-2)
#define FU_ERR_CUDA (-3)
#define FU_ERR_NOT_BOUND (-4)        /* parameters / gradients not bound yet */
#define FU_ERR_STATE (-5)            /* e.g. backward without a saved training forward */
#define FU_ERR_ARG (-6)

#define FU_PRECISION_FP32 0  /* parity mode: fp32 storage + fp32 FMA, matches unet.py to ~1e-5 */
#define FU_PRECISION_BF16 1  /* throughput mode: bf16 NHWC storage, tcgen05 bf16 MMA, fp32 accumulate */
#define FU_PRECISION_FP32_TC 2 /* parity mode on the tensor cores: fp32 NHWC storage; every MMA operand is read from a
                                  split-bf16 twin (hi = bf16(v), lo = bf16(v - hi)) and each contraction runs the three
                                  passes hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator (~2^-16 per product);
                                  matches unet.py to ~1e-5 like FU_PRECISION_FP32, at tensor-core speed */

#define FU_KIND_PARAM 0
#define FU_KIND_BUFFER 1
#define FU_DTYPE_F32 0
#define FU_DTYPE_I64 1

typedef struct fu_engine fu_engine;

/* Mirrors the keyword arguments of UNet.__init__ (unet.py:41-45), one field each. */
typedef struct fu_config {
  int32_t in_channels;
  int32_t n_classes;
  int32_t depth;
  int32_t wf;
  int32_t padding;            /* must be 1 (unet.py:42; padding=False crashes the reference with do_res) */
  int32_t pad_mode_zeros;     /* must be 1 ('zeros', unet.py:42) */
  int32_t batch_norm;
  int32_t up_mode_upconv;     /* must be 1 ('upconv', unet.py:43) */
  int32_t max_pool;
  int32_t num_lands;
  int32_t do_res;
  int32_t block_depth;
  int32_t lands_block_depth;  /* must be 0 (unet.py:44; dead in every reference script) */
  int32_t lands_num_1x1;
  int32_t do_soft_max;
  int32_t precision;          /* FU_PRECISION_* (not a reference argument) */
} fu_config;

typedef struct fu_tensor_info {
  char name[96];     /* state_dict key, e.g. "down_path.0.block.0.weight" */
  int32_t ndim;
  int64_t shape[4];
  int32_t kind;      /* FU_KIND_* */
  int32_t dtype;     /* FU_DTYPE_* */
  int64_t grad_offset; /* element offset into the flat fp32 gradient buffer, -1 if the
                          tensor receives no gradient (buffers; downsample_convs.{depth-1},
                          which unet.py:165-171 never calls) */
} fu_tensor_info;

typedef struct fu_counters {
  int64_t kernel_launches;      /* engine kernels launched since creation */
  int64_t tc_kernel_launches;   /* of which tcgen05 tensor-core kernels */
  int64_t forward_calls;
  int64_t backward_calls;
  int64_t arena_bytes;          /* current activation/workspace arena */
  int64_t last_fwd_launches;    /* launches of the most recent fu_forward */
  int64_t last_bwd_launches;    /* launches of the most recent fu_backward */
} fu_counters;

/* UNet.__init__ (unet.py:41-159): validates the configuration, builds the layer
 * table.  Rejects what no reference script selects with
 * FU_ERR_INVALID_CONFIG.  `device` is the CUDA ordinal (util.py:28-29 hard-wires 0). */
int fu_engine_create(const fu_config* cfg, int device, fu_engine** out);
/* nn.Module destruction. */
void fu_engine_destroy(fu_engine* e);
/* Message of the most recent failure on `e` (or of fu_engine_create when e == NULL). */
const char* fu_last_error(const fu_engine* e);

/* nn.Module.state_dict() schema (train.py:476): number of entries and
 * the i-th entry, in the reference's state_dict order. */
int fu_num_tensors(const fu_engine* e);
int fu_tensor_get_info(const fu_engine* e, int index, fu_tensor_info* out);
/* Number of fp32 elements of the flat gradient buffer fu_backward fills. */
int64_t fu_grad_numel(const fu_engine* e);

/* nn.Module.to(dev) / load_state_dict (train.py:316-319): hand the engine the device
 * address of every state_dict tensor, in schema order (fp32, or int64 for
 * num_batches_tracked).  May be called again whenever an address changes. */
int fu_bind_tensors(fu_engine* e, void* const* data_ptrs, int n);

/* UNet.forward (unet.py:161-193), called at train.py:407, util.py:144,202,271,331.
 *   x        (B,in_channels,H,W) fp32 NCHW
 *   training nn.Module.train()/eval() state: batch statistics + running-stat
 *            update (1) or running statistics (0)
 *   save     keep activations for fu_backward (0 under torch.no_grad())
 *   weights_version  any integer that changes whenever a parameter value changed
 *            since the last call (the engine re-packs its weight copies then)
 *   seg      (B,n_classes,H,W) softmax probabilities (logits if do_soft_max=0)
 *   logits   optional (may be NULL): seg_x of unet.py:176, same shape
 *   heat     (B,num_lands,H,W); NULL iff num_lands == 0
 */
int fu_forward(fu_engine* e, const float* x, int B, int H, int W, int training, int save,
               int64_t weights_version, float* seg, float* logits, float* heat, void* stream);

/* autograd backward of the above, entered at train.py:422.  d_seg / d_heat are
 * dL/dseg, dL/dheat (same shapes as the outputs, fp32 NCHW); either may be NULL
 * (= zero).  flat_grads (fu_grad_numel() floats) is overwritten with the gradient
 * of every reachable parameter at its fu_tensor_info.grad_offset.  d_x is not
 * produced: the input never requires grad in the reference (train.py:395-407). */
int fu_backward(fu_engine* e, const float* d_seg, const float* d_heat, float* flat_grads,
                void* stream);

/* Data parallelism (no reference counterpart: util.py:28-29 hard-wires one GPU).  The flat gradient buffer
 * is laid out so that elements [0, fu_early_grad_numel) are FINAL once the backward pass has left the deep encoder
 * levels -- heads, decoder, deep encoder: 97 % of the paper network's parameters.  With a callback registered,
 * fu_backward calls cb(user, 0, 0, fu_early_grad_numel) at that point, after making `comm_stream` wait (events, no
 * host synchronisation) for everything that produced those elements; the caller enqueues its all-reduce of
 * flat_grads[offset, offset + numel) on comm_stream, where it overlaps the rest of the backward pass, reduces the
 * remaining tail [fu_early_grad_numel, fu_grad_numel) after fu_backward returns, and makes its own stream wait for
 * comm_stream before the optimizer step.  cb == NULL removes the callback. */
typedef void (*fu_bucket_callback)(void* user, int bucket, int64_t offset, int64_t numel);
int fu_set_bucket_callback(fu_engine* e, fu_bucket_callback cb, void* user, void* comm_stream);
int64_t fu_early_grad_numel(const fu_engine* e);

int fu_get_counters(const fu_engine* e, fu_counters* out);

/* Per-launch CUDA-event profiling (the reference's only tracing is time.time(), train.py:377,
 * util.py:321).  While enabled every kernel launch is bracketed by events on the caller's stream;
 * fu_profile_report() synchronises and writes one JSON line per (layer tag, kernel) with the summed
 * device time, algorithmic FLOPs and bytes.  Returns the buffer size needed. */
int fu_profile_enable(fu_engine* e, int on);
int64_t fu_profile_report(fu_engine* e, char* buf, int64_t cap);

/* Read back one internal NHWC tensor of the last forward/backward as fp32 NCHW (per-layer parity
 * tests).  Names: "enc<l>.r<i>" / ".z<i>" (post-ReLU / post-BN activations of conv i of
 * encoder block l), ".dy<i>" / ".dz<i>" (their gradients), "dec<j>.*", "cat<l>", "d_cat<l>", "down<l>",
 * "d_down<l>", "decout<l>", "d_decout<l>", "bott", "d_bott", "hcat", "d_hcat".  shape4 receives
 * (B,C,H,W); with dst == NULL only the shape is returned.  Enqueued on the stream of the last call. */
int fu_debug_copy(fu_engine* e, const char* name, float* dst, int64_t capacity, int32_t* shape4);

/* ---- fused training loss -------------------------------------------------
 * DiceLoss2D / DiceAndHeatMapLoss2D of dice.py:14-86 over ncc_2d (ncc.py:12-38), with the centre
 * crop of the network outputs (util.py:92-114 as called at train.py:414-417) folded in: the
 * prediction pointers address the first element of the crop WINDOW inside the full (B,C,H,W) output
 * and the strides are those of the full tensor.  All tensors fp32, column stride 1.
 *   loss = dice_wgt * mean_b( mean_c( (-2 sum(t p) + 1e-4) / (sum(t^2) + sum(p^2) + 1e-4) ) )
 *        + heat_wgt * mean_{b,l}( -(ncc(heat, heat_t) + 1) / 2 )          [second term iff heat != NULL]
 * DiceLoss2D (train.py:327) is heat == NULL, dice_wgt = 1.  Stateless: no engine handle. */
typedef struct fu_loss_desc {
  const float* seg;    int64_t seg_stride[3];     /* batch, channel, row strides in elements */
  const float* mask;   int64_t mask_stride[3];    /* targets (B,n_classes,Ht,Wt) */
  const float* heat;   int64_t heat_stride[3];    /* NULL: Dice only */
  const float* heat_t; int64_t heat_t_stride[3];  /* (B,num_lands,Ht,Wt) */
  int32_t B, n_classes, num_lands;                /* num_lands = 0 when heat == NULL */
  int32_t Ht, Wt;                                 /* window size = target size */
  int32_t skip_bg;                                /* dice.py:16: leave class 0 out of the Dice mean */
  float dice_wgt, heat_wgt;                       /* dice.py:65-66: dice_wgt = 1 - heatmap_wgt */
} fu_loss_desc;
/* doubles of workspace both calls share: per image 3 sums per class + 5 per landmark */
int64_t fu_loss_workspace_doubles(int B, int n_classes, int num_lands);
/* sums (workspace) is zeroed and filled; loss_out is one device float. */
int fu_loss_forward(const fu_loss_desc* d, double* sums, float* loss_out, void* stream);
/* Gradients of the loss w.r.t. the FULL outputs: d_seg (B,n_classes,H,W), d_heat (B,num_lands,H,W;
 * NULL iff heat == NULL), contiguous, zero outside the window whose origin is (r0,c0).  dloss is the
 * upstream gradient (one device float).  `sums` is the workspace fu_loss_forward filled. */
int fu_loss_backward(const fu_loss_desc* d, const double* sums, const float* dloss, int H, int W,
                     int r0, int c0, float* d_seg, float* d_heat, void* stream);

/* The same loss INSIDE the network's head kernels (SURVEY 8f row 1 as specified; replaces the sequence
 * net(x) -> center_crop -> DiceAndHeatMapLoss2D -> loss.backward() of train.py:407-422 for one training step):
 *   fu_forward_loss  = fu_forward(training = 1, save = 1) whose head kernel also reduces the Dice / NCC sums of its own
 *                      outputs against the targets of `d` (d->seg / d->heat are ignored: the predictions need not leave
 *                      the kernel), then the one-block finalisation -> loss_out (one device float).  `seg` may be NULL
 *                      (class probabilities are then not written at all); `heat` (B,num_lands,H,W) is written and must be
 *                      handed to fu_backward_loss unchanged.  (r0, c0) is the crop-window origin (util.py:99-103).
 *   fu_backward_loss = fu_backward whose head kernel forms d_seg / d_heat per pixel from the targets, the sums and the
 *                      upstream gradient `dloss` (one device float) -- no gradient tensors are written or read.
 * bf16 storage, the paper heads (32 features, 7 classes, 0 or 14 landmarks) and H*W % 16 == 0 only:
 * FU_ERR_UNSUPPORTED_SHAPE otherwise (use fu_forward + fu_loss_forward + fu_loss_backward + fu_backward). */
int fu_forward_loss(fu_engine* e, const float* x, int B, int H, int W, int64_t weights_version, const fu_loss_desc* d,
                    int r0, int c0, double* sums, float* loss_out, float* seg, float* heat, void* stream);
int fu_backward_loss(fu_engine* e, const fu_loss_desc* d, int r0, int c0, const double* sums, const float* dloss,
                     const float* heat, float* flat_grads, void* stream);

/* ---- callers either side of the path ---------------------------------------
 * Sample preparation before the network and inference post-processing after it, as device kernels
 * on the tensors the reference's host code holds (fp32 NCHW, u1 labels).  Stateless (no engine
 * handle); errors through fu_last_error(NULL); everything is enqueued on `stream`, no host sync. */

/* dataset.py:287-293 (RandomDataAugDataSet.__getitem__): reflect-pad every tile by `pad` pixels per
 * side (numpy.pad mode 'reflect'; calc_pad_amount, dataset.py:26-40, gives pad) and, when `normalize`,
 * z-score it with the mean and UNBIASED standard deviation of the padded tile.
 *   tiles (B,h,w) fp32 -> out (B,1,h+2*pad,w+2*pad) fp32;  sums: 2*B doubles of workspace. */
int fu_prep_tiles(const float* tiles, int B, int h, int w, int pad, int normalize, double* sums, float* out,
                  void* stream);

/* dataset.py:295-325: Gaussian heat-map targets exp(-((X-x)^2+(Y-y)^2)/(2 sigma^2)) / (2 pi sigma^2)
 * (sigma 2.5 in the reference), a zero plane for a landmark whose x or y is +-inf (outside the view,
 * dataset.py:316).   lands (B,2,num_lands) fp32, row 0 = column (x), row 1 = row (y) -> out (B,num_lands,H,W). */
int fu_heatmap_targets(const float* lands, int B, int num_lands, int H, int W, float sigma, float* out, void* stream);

/* util.py:293-377 (seg_dataset_ensemble), the arithmetic between the networks' forward calls and the
 * HDF5 write: centre-crop window (r0,c0,h,w) of every network's outputs (util.py:339,347), class
 * probabilities summed in list order and divided by n_nets, arg-max over classes (first maximum wins)
 * -> labels u1 (B,h,w); heat-maps min-max normalised per (network, image) (util.py:348-351), summed
 * and divided by n_nets -> avg_heat (B,num_lands,h,w).
 *   seg / heat: HOST arrays of n_nets (<= 16) DEVICE pointers to (B,n_classes,H,W) / (B,num_lands,H,W);
 *   heat, avg_heat, workspace NULL iff num_lands == 0; workspace: fu_ensemble_workspace_words() uint32. */
int64_t fu_ensemble_workspace_words(int n_nets, int B);
int fu_ensemble_combine(const float* const* seg, const float* const* heat, int n_nets, int B, int n_classes,
                        int num_lands, int H, int W, int r0, int c0, int h, int w, uint32_t* workspace,
                        uint8_t* labels, float* avg_heat, void* stream);

/* est_lands_csv.py:87-134 (rule_3): per (projection, landmark) the arg-max pixel of the heat-map,
 * restricted to pixels whose label in `segs` equals seg_labels[landmark] when segs != NULL and the
 * label is >= 0 (est_lands_csv.py:54-71 holds the label table); kept only if the NCC (ncc.py:12-38) of
 * the tmpl_dim x tmpl_dim Gaussian template (util.py:36-48; 25 and sigma 2.5 in the reference) with the
 * window of the reflect-padded heat-map centred there is >= min_ncc (0.9).
 *   heats (P,num_lands,h,w) fp32 device; segs (P,h,w) u1 device or NULL; seg_labels HOST array of
 *   num_lands (<= 64) ints or NULL;  out_rc (P,num_lands,2) int32 row,col, (-1,-1) = not found
 *   (est_lands_csv.py:126-128);  out_ncc optional (P,num_lands) scores. */
int fu_extract_landmarks(const float* heats, const uint8_t* segs, const int32_t* seg_labels, int P, int num_lands,
                         int h, int w, int tmpl_dim, float sigma, float min_ncc, int32_t* out_rc, float* out_ncc,
                         void* stream);

/* Build information: "sm_100a;tcgen05=1;..." */
const char* fu_build_info(void);

/* ---- kernel-level test hooks (used by tests/ only; no reference counterpart) ----
 * Run one convolution layer through the engine's kernels on caller-provided NHWC
 * tensors so each kernel can be checked against the oracle in isolation.
 *   mode 0: forward 3x3/1x1/2x2s2 conv   y = conv(x, w) + bias [+relu]
 *   mode 1: data gradient                dx = conv^T(dy, w)
 *   mode 2: weight gradient              dw = x (*) dy
 * `impl` 0 = fp32/bf16 CUDA-core path, 1 = tcgen05 path (bf16 tensors), 2 = tcgen05 split-bf16 x3 parity path
 * (fp32 tensors, FU_PRECISION_FP32_TC).
 * Tensors: x (B,H,W,Cin), y/dy (B,Ho,Wo,Cout) in the engine precision's storage
 * type (fp32 or bf16 bits); w, bias, dw fp32 in the torch layout (Cout,Cin,k,k). */
int fu_test_conv(int precision, int impl, int mode, int B, int H, int W, int Cin, int Cout,
                 int ksize, int stride, int pad, int relu,
                 const void* x, const float* w, const float* bias, void* y_or_dx,
                 const void* dy, float* dw, double* stats_or_null, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUORO_UNET_H_ */
