"""Drop-in for the reference's ``train_test_code/unet.py``.

``UNet`` has the constructor signature (unet.py:41-45), the return convention
(unet.py:190-193), the ``state_dict`` keys/shapes and the default
initialisation of the reference module, but ``forward`` and its autograd
backward run on the B200 engine behind ``include/fluoro_unet.h``.

The torch sub-modules created here (``nn.Conv2d`` ...) are parameter containers
only: they give the same names, shapes and RNG-identical default init as the
reference; their own ``forward`` is never called.  There is no CPU / cuDNN
fallback: a non-CUDA input or a missing ``libfluorounet.so`` raises.
"""
import ctypes as C
import os

import torch
from torch import nn

from . import _capi


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the B200 engine; call UNet.forward instead")


class UNetConvBlock(_ParamsOnly):
    """Parameters of unet.py:196-224 (res_conv1x1 first, then block.{0,2,3,5,...})."""

    def __init__(self, in_size, out_size, padding, batch_norm, pad_mode, do_res, block_depth):
        super().__init__()
        assert block_depth > 0
        self.do_res = do_res
        if do_res:
            self.res_conv1x1 = nn.Conv2d(in_size, out_size, kernel_size=1, padding=0)
        block = [nn.Conv2d(in_size, out_size, kernel_size=3, padding=int(padding), padding_mode=pad_mode),
                 nn.ReLU()]
        if batch_norm:
            block.append(nn.BatchNorm2d(out_size))
        for _ in range(block_depth - 1):
            block.append(nn.Conv2d(out_size, out_size, kernel_size=3, padding=int(padding), padding_mode=pad_mode))
            block.append(nn.ReLU())
            if batch_norm:
                block.append(nn.BatchNorm2d(out_size))
        self.block = nn.Sequential(*block)


class UNetUpBlock(_ParamsOnly):
    """Parameters of unet.py:236-246."""

    def __init__(self, in_size, out_size, up_mode, padding, batch_norm, pad_mode, do_res, block_depth):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_size, out_size, kernel_size=2, stride=2)
        self.conv_block = UNetConvBlock(in_size, out_size, padding, batch_norm, pad_mode, do_res=do_res,
                                        block_depth=block_depth)


class _UNetFunction(torch.autograd.Function):
    """One autograd node for the whole network (train.py:407 forward, :422 backward)."""

    @staticmethod
    def forward(ctx, net, save, x, *params):
        seg, heat = net._engine_forward(x, save)
        ctx.net = net
        ctx.generation = net._generation
        ctx.n_params = len(params)
        if heat is None:
            return seg
        return seg, heat

    @staticmethod
    def backward(ctx, d_seg, d_heat=None):
        net = ctx.net
        if ctx.generation != net._generation:
            raise RuntimeError("UNet.backward: the engine's saved activations were overwritten by a later "
                               "forward; run backward before the next forward (as train.py:407-422 does)")
        grads = net._engine_backward(d_seg, d_heat)
        return (None, None, None) + tuple(grads)


class _UNetLossFunction(torch.autograd.Function):
    """Network forward + training loss as ONE autograd node (train.py:407-422 without the output tensors in between): the
    head kernel reduces the Dice / NCC sums of its own outputs, the backward head kernel forms the loss gradient per pixel
    (include/fluoro_unet.h: fu_forward_loss / fu_backward_loss)."""

    @staticmethod
    def forward(ctx, net, x, tgt_seg, tgt_heat, skip_bg, dice_wgt, heat_wgt, *params):
        L = _capi.lib()
        B, _, H, W = x.shape
        nc, nl = net._cfg["n_classes"], net._cfg["num_lands"]
        Ht, Wt = tgt_seg.shape[-2:]
        r0, c0 = int((H - Ht) / 2), int((W - Wt) / 2)              # util.py:99-103
        d = _capi.FuLossDesc()
        d.mask = tgt_seg.data_ptr()
        d.mask_stride[:] = [tgt_seg.stride(0), tgt_seg.stride(1), tgt_seg.stride(2)]
        if nl > 0:
            d.heat_t = tgt_heat.data_ptr()
            d.heat_t_stride[:] = [tgt_heat.stride(0), tgt_heat.stride(1), tgt_heat.stride(2)]
        d.B, d.n_classes, d.num_lands, d.Ht, d.Wt = B, nc, nl, Ht, Wt
        d.skip_bg, d.dice_wgt, d.heat_wgt = int(skip_bg), float(dice_wgt), float(heat_wgt)
        sums = torch.empty(int(L.fu_loss_workspace_doubles(B, nc, nl)), device=x.device, dtype=torch.float64)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        heat = torch.empty((B, nl, H, W), device=x.device, dtype=torch.float32) if nl > 0 else None
        version = 0
        for p in net._state_cache[3]:
            version += p._version
        version = (version & 0xFFFFFFFF) | (net._pack_epoch << 32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        net._generation += 1
        rc = L.fu_forward_loss(net._handle, x.data_ptr(), B, H, W, version, C.byref(d), r0, c0, sums.data_ptr(), loss.data_ptr(),
                               None, heat.data_ptr() if heat is not None else None, stream)
        if rc != 0:
            msg = _capi.last_error(net._handle)
            if rc == _capi.FU_ERR_UNSUPPORTED_SHAPE:
                raise ValueError(msg)
            raise RuntimeError(f"fu_forward_loss failed ({rc}): {msg}")
        ctx.net, ctx.generation = net, net._generation
        ctx.desc, ctx.geom = d, (r0, c0)
        ctx.keep = (tgt_seg, tgt_heat, sums, heat)                   # the descriptor holds raw addresses of these
        return loss

    @staticmethod
    def backward(ctx, dloss):
        net = ctx.net
        if ctx.generation != net._generation:
            raise RuntimeError("UNet.backward: the engine's saved activations were overwritten by a later "
                               "forward; run backward before the next forward (as train.py:407-422 does)")
        L = _capi.lib()
        _, _, sums, heat = ctx.keep
        dev = net._handle_device
        flat = torch.empty(net._grad_numel, device=dev, dtype=torch.float32)
        dloss = dloss.contiguous().float()
        stream = torch.cuda.current_stream(dev).cuda_stream
        net._install_bucket_callback(flat)
        rc = L.fu_backward_loss(net._handle, C.byref(ctx.desc), ctx.geom[0], ctx.geom[1], sums.data_ptr(), dloss.data_ptr(),
                                heat.data_ptr() if heat is not None else None, flat.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"fu_backward_loss failed ({rc}): {_capi.last_error(net._handle)}")
        net._pack_epoch = (net._pack_epoch + 1) & 0x3FFFFFFF
        if net.grad_hook is not None:
            net.grad_hook(flat)
        net.last_flat_grad = flat
        grads = [flat[off:off + numel].view(shape) for name, shape, numel, off in net._grad_params]
        return (None,) * 7 + tuple(grads)


class UNet(nn.Module):
    def __init__(self, in_channels=1, n_classes=2, depth=5, wf=6,
                 padding=False, pad_mode='zeros',
                 batch_norm=False, up_mode='upconv', max_pool=True, num_lands=0,
                 do_res=True, block_depth=2, lands_block_depth=0, lands_num_1x1=2,
                 do_soft_max=True, precision=None):
        """Same arguments as the reference (unet.py:41-45) plus ``precision``:
        ``'fp32'`` (parity mode on the CUDA cores, default), ``'parity_tc'`` (parity mode on the tensor cores:
        fp32 storage, every contraction as three split-bf16 tcgen05 passes, ~1e-5 from the reference) or
        ``'bf16'`` (throughput mode).  The default can be overridden with the environment variable
        FLUORO_UNET_PRECISION."""
        super().__init__()
        if up_mode not in ('upconv', 'upsample'):
            raise ValueError("up_mode must be 'upconv' or 'upsample'")
        # configurations no reference script selects are rejected, never emulated
        if not padding:
            raise ValueError("padding=False is not supported by the B200 engine (it also crashes the reference "
                             "with do_res=True, unet.py:227-231); pass padding=True")
        if pad_mode != 'zeros':
            raise ValueError("pad_mode must be 'zeros'")
        if up_mode != 'upconv':
            raise ValueError("up_mode='upsample' is not supported; use 'upconv'")
        if lands_block_depth != 0:
            raise ValueError("lands_block_depth > 0 is not supported")
        if block_depth < 1:
            raise ValueError("block_depth must be >= 1")
        if wf < 2:
            raise ValueError("wf must be >= 2")
        if num_lands > 0 and lands_num_1x1 < 1:
            raise ValueError("lands_num_1x1 must be >= 1")
        if precision is None:
            precision = os.environ.get("FLUORO_UNET_PRECISION", "fp32")
        if precision not in _capi.PRECISION:
            raise ValueError("precision must be 'fp32', 'parity_tc' or 'bf16'")
        if precision == "bf16" and wf < 3:
            raise ValueError("throughput mode (bf16) needs wf >= 3 (16-byte channel vectors); use precision='fp32'")
        self.padding = padding
        self.pad_mode = pad_mode
        self.depth = depth
        self.do_max_pool = max_pool
        self.num_lands = num_lands
        self.do_soft_max = do_soft_max
        self.precision = precision
        self._cfg = dict(in_channels=in_channels, n_classes=n_classes, depth=depth, wf=wf, padding=1,
                         pad_mode_zeros=1, batch_norm=int(batch_norm), up_mode_upconv=1, max_pool=int(max_pool),
                         num_lands=num_lands, do_res=int(do_res), block_depth=block_depth, lands_block_depth=0,
                         lands_num_1x1=lands_num_1x1, do_soft_max=int(do_soft_max),
                         precision=_capi.PRECISION[precision])

        # ---- parameter containers, created in the reference's order (unet.py:78-159) ----
        self.downsample_convs = None
        if not self.do_max_pool:
            self.downsample_convs = nn.ModuleList()
        prev_channels = in_channels
        self.down_path = nn.ModuleList()
        for i in range(depth):
            self.down_path.append(UNetConvBlock(prev_channels, 2 ** (wf + i), padding, batch_norm, pad_mode,
                                                do_res=do_res, block_depth=block_depth))
            prev_channels = 2 ** (wf + i)
            if not self.do_max_pool:
                self.downsample_convs.append(nn.Conv2d(prev_channels, prev_channels, kernel_size=2, stride=2))
        self.up_path = nn.ModuleList()
        for i in reversed(range(depth - 1)):
            self.up_path.append(UNetUpBlock(prev_channels, 2 ** (wf + i), up_mode, padding, batch_norm, pad_mode,
                                            do_res=do_res, block_depth=block_depth))
            prev_channels = 2 ** (wf + i)
        self.seg_conv = nn.Conv2d(prev_channels, n_classes, kernel_size=1, bias=False)
        if do_soft_max:
            self.soft_max = nn.Softmax2d()
        if self.num_lands > 0:
            self.lands_block = None
            lands_1x1 = []
            nf = num_lands + n_classes if lands_num_1x1 > 1 else num_lands
            lands_1x1.append(nn.Conv2d(prev_channels + n_classes, nf, kernel_size=1, bias=False))
            for _ in range(lands_num_1x1 - 1):
                lands_1x1.append(nn.Conv2d(nf, num_lands, kernel_size=1, bias=False))
                nf = num_lands
            self.lands_1x1 = nn.Sequential(*lands_1x1)

        # ---- engine state (created lazily on the first CUDA forward) ----
        self._handle = None
        self._handle_device = None
        self._schema = None
        self._bound_ptrs = None
        self._state_cache = None
        self._pack_epoch = 0
        self._generation = 0
        self._grad_numel = 0
        self._last_logits = None
        self.keep_logits = False      # when True, forward also stores seg_x (unet.py:176) in .last_logits
        self.grad_hook = None         # callable(flat_fp32_grads) run before gradients are handed to autograd
        # callable(flat, offset, numel) run MID-backward, on `bucket_stream`, as soon as flat[offset:offset+numel] (the
        # early gradient bucket: heads, decoder, deep encoder levels) is final; see parallel.data_parallel(overlap=True)
        self.grad_bucket_hook = None
        self.bucket_stream = None
        self._bucket_cb = None        # keeps the ctypes trampoline alive
        self._bucket_cb_key = None

    # ------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------
    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def _destroy_engine(self):
        if getattr(self, "_handle", None) is not None:
            _capi.lib().fu_engine_destroy(self._handle)
            self._handle = None
            self._bound_ptrs = None

    def _ensure_engine(self, device):
        if self._handle is not None and self._handle_device == device:
            return
        self._destroy_engine()
        L = _capi.lib()
        cfg = _capi.FuConfig(**self._cfg)
        handle = C.c_void_p()
        rc = L.fu_engine_create(C.byref(cfg), device.index if device.index is not None else torch.cuda.current_device(),
                                C.byref(handle))
        if rc != 0:
            msg = _capi.last_error(None)
            if rc == _capi.FU_ERR_INVALID_CONFIG:
                raise ValueError(msg)
            raise RuntimeError(f"fu_engine_create failed ({rc}): {msg}")
        self._handle = handle
        self._handle_device = device
        n = L.fu_num_tensors(handle)
        schema = []
        info = _capi.FuTensorInfo()
        for i in range(n):
            L.fu_tensor_get_info(handle, i, C.byref(info))
            schema.append((info.name.decode(), tuple(info.shape[:info.ndim]), info.kind, info.dtype,
                           info.grad_offset))
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        names = [s[0] for s in schema]
        if set(names) != set(sd.keys()):
            raise RuntimeError("engine schema and module state_dict disagree: "
                               f"{sorted(set(names) ^ set(sd.keys()))[:6]}")
        for name, shape, _, _, _ in schema:
            if tuple(sd[name].shape) != shape:
                raise RuntimeError(f"shape mismatch for {name}: module {tuple(sd[name].shape)} engine {shape}")
        self._schema = schema
        self._grad_numel = int(L.fu_grad_numel(handle))
        self._bound_ptrs = None
        self._state_cache = None
        self._bucket_cb_key = None

    def _state_tensors(self):
        """Parameters and buffers in engine-schema order.  Walking the module tree costs ~0.4 ms of Python per
        call, which sits on the critical path of every step whose loss is read back (train.py:430), so the
        list is cached; `_apply` (.to/.cuda) and `invalidate_cache()` drop it."""
        if self._state_cache is None:
            sd = dict(self.named_parameters())
            sd.update(dict(self.named_buffers()))
            tensors = [sd[name] for name, *_ in self._schema]
            gp, params = [], []
            for (name, shape, kind, _, off), t in zip(self._schema, tensors):
                if kind == 0 and off >= 0:
                    params.append(t)
                    n = 1
                    for d in shape:
                        n *= d
                    gp.append((name, shape, n, off))
            self._state_cache = (tensors, params, gp, list(self.parameters()))
        return self._state_cache[0]

    def invalidate_cache(self):
        """Call after replacing a Parameter/buffer OBJECT of the module by hand, or after changing parameter
        VALUES outside autograd's version tracking (e.g. a fused optimizer step that was not preceded by this
        module's backward).  In-place updates, load_state_dict and .to() need nothing."""
        self._state_cache = None
        self._bound_ptrs = None
        self._pack_epoch = (self._pack_epoch + 1) & 0x3FFFFFFF

    def mark_weights_changed(self):
        """Parameter VALUES changed outside autograd's version tracking (a CUDA-graph replay of a training step, a
        `.data` write such as dist.broadcast(p.data)): the next forward re-packs the engine's weight copies."""
        self._pack_epoch = (self._pack_epoch + 1) & 0x3FFFFFFF

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._state_cache = None
        self._bound_ptrs = None
        return out

    def _bind(self, device):
        tensors = self._state_tensors()
        ptrs = [t.data_ptr() for t in tensors]
        if ptrs == self._bound_ptrs:      # same storages as the last call: already validated and bound
            return tensors
        ptrs = []
        for (name, _, _, dtype, _), t in zip(self._schema, tensors):
            want = torch.int64 if dtype == 1 else torch.float32
            if t.device != device:
                raise RuntimeError(f"{name} is on {t.device} but the input is on {device}; call net.to(device)")
            if t.dtype != want:
                raise TypeError(f"{name} has dtype {t.dtype}; the engine keeps master weights in {want} "
                                "(do not call .half()/.double() on this module)")
            if not t.is_contiguous():
                raise RuntimeError(f"{name} is not contiguous")
            ptrs.append(t.data_ptr())
        if ptrs != self._bound_ptrs:
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            rc = _capi.lib().fu_bind_tensors(self._handle, arr, len(ptrs))
            if rc != 0:
                raise RuntimeError(f"fu_bind_tensors failed ({rc}): {_capi.last_error(self._handle)}")
            self._bound_ptrs = ptrs
        return tensors

    def _engine_forward(self, x, save):
        L = _capi.lib()
        B, Cin, H, W = x.shape
        nc, nl = self._cfg["n_classes"], self._cfg["num_lands"]
        seg = torch.empty((B, nc, H, W), device=x.device, dtype=torch.float32)
        heat = torch.empty((B, nl, H, W), device=x.device, dtype=torch.float32) if nl > 0 else None
        logits = torch.empty_like(seg) if self.keep_logits else None
        # weights_version: tensor version counters catch ordinary in-place updates (optimizer.step, copy_,
        # load_state_dict); fused / foreach optimizers (torch.optim.SGD(fused=True)) update parameters WITHOUT
        # bumping them, so every backward also advances an epoch: the forward after a backward always re-packs.
        version = 0
        for p in self._state_cache[3]:
            version += p._version
        version = (version & 0xFFFFFFFF) | (self._pack_epoch << 32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._generation += 1
        # (the C entry points make the engine's device current for their launches and restore the caller's)
        rc = L.fu_forward(self._handle, x.data_ptr(), B, H, W, int(self.training), int(save), version,
                          seg.data_ptr(), logits.data_ptr() if logits is not None else None,
                          heat.data_ptr() if heat is not None else None, stream)
        if rc != 0:
            msg = _capi.last_error(self._handle)
            if rc == _capi.FU_ERR_UNSUPPORTED_SHAPE:
                raise ValueError(msg)
            raise RuntimeError(f"fu_forward failed ({rc}): {msg}")
        self._last_logits = logits
        return seg, heat

    def _engine_backward(self, d_seg, d_heat):
        L = _capi.lib()
        dev = self._handle_device
        flat = torch.empty(self._grad_numel, device=dev, dtype=torch.float32)

        def prep(g):
            if g is None:
                return None
            return g.contiguous().float()
        d_seg, d_heat = prep(d_seg), prep(d_heat)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._install_bucket_callback(flat)
        rc = L.fu_backward(self._handle, d_seg.data_ptr() if d_seg is not None else None,
                           d_heat.data_ptr() if d_heat is not None else None, flat.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"fu_backward failed ({rc}): {_capi.last_error(self._handle)}")
        self._pack_epoch = (self._pack_epoch + 1) & 0x3FFFFFFF   # an optimizer step usually follows
        if self.grad_hook is not None:
            self.grad_hook(flat)
        self.last_flat_grad = flat
        grads = []
        for name, shape, numel, off in self._grad_params:
            grads.append(flat[off:off + numel].view(shape))
        return grads

    def _install_bucket_callback(self, flat):
        """Register (or remove) the engine's mid-backward callback for the early gradient bucket."""
        L = _capi.lib()
        key = (self.grad_bucket_hook, self.bucket_stream)
        self._cur_flat = flat
        if key == self._bucket_cb_key:
            return
        if self.grad_bucket_hook is None:
            L.fu_set_bucket_callback(self._handle, _capi.BUCKET_CALLBACK(), None, None)
            self._bucket_cb = None
        else:
            if self.bucket_stream is None:
                raise RuntimeError("grad_bucket_hook needs bucket_stream (a torch.cuda.Stream)")
            hook, comm = self.grad_bucket_hook, self.bucket_stream

            def tramp(_user, _bucket, offset, numel):
                with torch.cuda.stream(comm):
                    hook(self._cur_flat, int(offset), int(numel))
            self._bucket_cb = _capi.BUCKET_CALLBACK(tramp)
            rc = L.fu_set_bucket_callback(self._handle, self._bucket_cb, None, comm.cuda_stream)
            if rc != 0:
                raise RuntimeError(f"fu_set_bucket_callback failed ({rc}): {_capi.last_error(self._handle)}")
        self._bucket_cb_key = key

    @property
    def early_grad_numel(self):
        """Elements of the flat gradient buffer that form the early bucket (include/fluoro_unet.h)."""
        return int(_capi.lib().fu_early_grad_numel(self._handle)) if self._handle is not None else 0

    @property
    def last_logits(self):
        """seg_x of unet.py:176 from the most recent forward (only when keep_logits=True)."""
        return self._last_logits

    def engine_counters(self):
        cnt = _capi.FuCounters()
        if self._handle is None:
            return {}
        _capi.lib().fu_get_counters(self._handle, C.byref(cnt))
        return {n: int(getattr(cnt, n)) for n, _ in cnt._fields_}

    def debug_tensor(self, name):
        """fp32 NCHW copy of an internal engine tensor of the last forward/backward (see fu_debug_copy)."""
        L = _capi.lib()
        shape = (C.c_int32 * 4)()
        rc = L.fu_debug_copy(self._handle, name.encode(), None, 0, shape)
        if rc != 0:
            raise KeyError(_capi.last_error(self._handle))
        out = torch.empty(tuple(shape), device=self._handle_device, dtype=torch.float32)
        rc = L.fu_debug_copy(self._handle, name.encode(), out.data_ptr(), out.numel(), shape)
        if rc != 0:
            raise RuntimeError(_capi.last_error(self._handle))
        return out

    def profile(self, on=True):
        """Start/stop per-launch CUDA-event profiling inside the engine."""
        if self._handle is None:
            raise RuntimeError("run one forward first")
        _capi.lib().fu_profile_enable(self._handle, int(on))

    def profile_report(self):
        """List of dicts {tag, kernel, launches, ms, flops, bytes} since profile(True)."""
        import json
        L = _capi.lib()
        n = L.fu_profile_report(self._handle, None, 0)
        buf = C.create_string_buffer(int(n) + 16)
        L.fu_profile_report(self._handle, buf, n + 16)
        return [json.loads(l) for l in buf.value.decode().splitlines() if l.strip()]

    def forward_loss(self, x, target, criterion):
        """loss = criterion(center_crop(net(x)), target) of train.py:407-420 as ONE engine call pair: the Dice / NCC sums are
        reduced by the head kernel itself and `loss.backward()` forms the loss gradient inside the backward head kernel, so
        neither the (B, 7 + 14, H, W) outputs nor their gradients make a round trip through HBM.

        criterion: FusedDiceLoss2D (target = mask) or FusedDiceAndHeatMapLoss2D (target = (mask, heat-maps)); targets are
        fp32 CUDA tensors of the cropped size.  Needs precision='bf16', the paper heads and H*W % 16 == 0 -- anything else
        raises ValueError (use `criterion(net(x), target)`, which runs the same arithmetic as separate kernels).  The module
        must be in train() mode; gradients reach every parameter exactly as through forward()."""
        from . import losses
        if isinstance(criterion, losses.FusedDiceAndHeatMapLoss2D):
            tgt_seg, tgt_heat = target
            skip_bg, dice_wgt, heat_wgt = criterion.skip_bg, criterion.dice_wgt, criterion.heatmap_wgt
        elif isinstance(criterion, losses.FusedDiceLoss2D):
            tgt_seg, tgt_heat = target, None
            skip_bg, dice_wgt, heat_wgt = criterion.skip_bg, 1.0, 0.0
        else:
            raise TypeError("forward_loss: criterion must be FusedDiceLoss2D or FusedDiceAndHeatMapLoss2D")
        if not self.training:
            raise RuntimeError("forward_loss is a training-step call: put the module in train() mode")
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or not x.is_cuda or x.dtype != torch.float32:
            raise ValueError("forward_loss expects a (B,C,H,W) float32 CUDA tensor")
        if x.shape[1] != self._cfg["in_channels"]:
            raise ValueError(f"expected {self._cfg['in_channels']} input channels, got {x.shape[1]}")
        nl = self._cfg["num_lands"]
        if (nl > 0) != (tgt_heat is not None):
            raise ValueError("forward_loss: heat-map targets must be given exactly when the network has a landmark head")
        for name, t in (("mask", tgt_seg), ("heat-map target", tgt_heat)):
            if t is None:
                continue
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.stride(-1) == 1):
                raise ValueError(f"forward_loss: {name} must be a 4-D float32 CUDA tensor with unit column stride")
        B, _, H, W = x.shape
        if tuple(tgt_seg.shape[:2]) != (B, self._cfg["n_classes"]) or tgt_seg.shape[-2] > H or tgt_seg.shape[-1] > W:
            raise ValueError(f"forward_loss: mask {tuple(tgt_seg.shape)} does not fit the network output")
        if tgt_heat is not None and tuple(tgt_heat.shape) != (B, nl) + tuple(tgt_seg.shape[-2:]):
            raise ValueError("forward_loss: heat-map target shape disagrees with the mask")
        x = x.contiguous()
        self._ensure_engine(x.device)
        self._bind(x.device)
        params, self._grad_params = self._state_cache[1], self._state_cache[2]
        return _UNetLossFunction.apply(self, x, tgt_seg, tgt_heat, skip_bg, dice_wgt, heat_wgt, *params)

    # ------------------------------------------------------------------
    # nn.Module interface
    # ------------------------------------------------------------------
    def forward(self, x):
        """unet.py:161-193.  x: (B, in_channels, H, W) fp32 on a CUDA device."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("UNet.forward expects a (B,C,H,W) tensor")
        if not x.is_cuda:
            raise RuntimeError("the B200 U-Net engine runs on CUDA devices only; there is no CPU fallback "
                               "(move the module and the input to a cuda device)")
        if x.shape[1] != self._cfg["in_channels"]:
            raise ValueError(f"expected {self._cfg['in_channels']} input channels, got {x.shape[1]}")
        if x.dtype != torch.float32:
            raise TypeError("UNet.forward expects a float32 input (as the reference's dataset produces)")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("UNet.forward: the engine does not produce a gradient for the INPUT (the reference never asks "
                               "for one: train.py:395-407 feeds data tensors); detach the input")
        x = x.contiguous()
        self._ensure_engine(x.device)
        self._bind(x.device)
        params, self._grad_params = self._state_cache[1], self._state_cache[2]
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if save:
            return _UNetFunction.apply(self, True, x, *params)
        seg, heat = self._engine_forward(x, False)
        return seg if heat is None else (seg, heat)
