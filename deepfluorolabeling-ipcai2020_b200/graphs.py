"""CUDA-graph replay of a whole training step.

The reference's loop (train.py:395-430) reads the loss back every step (`loss.item()`), so the host cannot run
ahead of the device: the ~215 kernel launches of one step, plus the Python around them, are re-issued from an
idle pipeline every time.  `GraphedStep` captures ONE step -- zero_grad, UNet forward, loss, backward (incl. the
engine's weight re-pack and gradient unpack), optimizer.step -- into a CUDA graph and replays it per batch, so a
step costs one launch.  Streams and graphs instead of a tracing compiler: nothing is re-written, the captured
kernels are exactly the ones the eager path launches.

Requirements (checked or documented, never silently worked around):
  * fixed input shapes (the reference trains on fixed-size tiles, dataset.py:26-40);
  * `step_fn` must not synchronise or read device values on the host (return the loss tensor, read it outside);
  * optimisers whose step is capture-safe (torch.optim.SGD, also fused=True; Adam needs capturable=True);
  * under torch.distributed the NCCL gradient all-reduce of `parallel.data_parallel` is captured too
    (allow_distributed=True): every rank has to capture and replay in lock step.
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, example_inputs, warmup=3, allow_distributed=False, modules=()):
        """step_fn(*inputs) -> loss tensor (or tuple of tensors); example_inputs: CUDA tensors of the step's shapes.
        allow_distributed: also capture under torch.distributed (the NCCL gradient all-reduce of
        parallel.data_parallel becomes a node of the graph; every rank must capture and replay in lock step).
        modules: the UNet modules the step trains.  A replay changes their parameters on the device without running
        any Python, so neither the tensors' version counters nor the modules' pack epochs move; every replay
        therefore tells them (`UNet.mark_weights_changed`) so that the next EAGER forward -- e.g. the per-epoch
        validation of train.py:446-454 -- re-packs the engine's weight copies instead of using the ones the last
        replayed step packed before its optimizer update."""
        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise ValueError("GraphedStep: example_inputs must be CUDA tensors")
        if (not allow_distributed and torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1):
            raise RuntimeError("GraphedStep: pass allow_distributed=True to capture the data-parallel step "
                               "(all ranks must then capture and replay together)")
        self.step_fn = step_fn
        self.modules = [m for m in modules if hasattr(m, "mark_weights_changed")]
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for d, s in zip(self.static_in, example_inputs):
            d.copy_(s)
        dev = example_inputs[0].device
        # warm-up on a side stream: lazy allocations (engine arena, tensor maps, job tables, optimizer state)
        # must all have happened before the capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)
        torch.cuda.synchronize(dev)

    def __call__(self, *inputs):
        """Copies `inputs` into the graph's static buffers (device-to-device, on the current stream), replays the
        step and returns the static output tensor(s); their values are overwritten by the next call."""
        if len(inputs) != len(self.static_in):
            raise ValueError("GraphedStep: wrong number of inputs")
        for d, s in zip(self.static_in, inputs):
            if s.shape != d.shape or s.dtype != d.dtype:
                raise ValueError(f"GraphedStep: input {tuple(s.shape)}/{s.dtype} does not match the captured {tuple(d.shape)}/{d.dtype}")
            if s.data_ptr() != d.data_ptr():
                d.copy_(s, non_blocking=True)
        self.graph.replay()
        for m in self.modules:
            m.mark_weights_changed()
        return self.static_out


class GraphedForward:
    """CUDA-graph replay of the inference forward of one or several networks on a fixed input shape.

    The ensemble loop of util.py:321-366 runs every network on every image batch; eagerly that is ~60 kernel launches
    per network per batch issued from Python through ctypes, and the host cannot keep three B200-sized forward passes
    fed.  Capturing the eval-mode, no-grad forwards of all `nets` on one static input -- the networks run back to back
    inside ONE graph, programmatic dependent launches included -- makes a batch one launch.  The networks must be in
    eval() mode and their weights frozen: the engine packs its weight copies when it sees a new parameter version, which a
    replay cannot do, so re-create the object after load_state_dict / optimizer steps."""

    def __init__(self, nets, example_x, warmup=2):
        if not isinstance(example_x, torch.Tensor) or not example_x.is_cuda:
            raise ValueError("GraphedForward: example_x must be a CUDA tensor")
        nets = list(nets)
        if not nets or any(n.training for n in nets):
            raise ValueError("GraphedForward: pass networks in eval() mode")
        self.nets = nets
        self.static_x = example_x.clone()
        dev = example_x.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                for n in nets:
                    n(self.static_x)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = [n(self.static_x) for n in nets]
        torch.cuda.synchronize(dev)

    def __call__(self, x):
        """Returns the list of the networks' outputs (static tensors, overwritten by the next call)."""
        if x.shape != self.static_x.shape or x.dtype != self.static_x.dtype:
            raise ValueError(f"GraphedForward: input {tuple(x.shape)}/{x.dtype} does not match the captured "
                             f"{tuple(self.static_x.shape)}/{self.static_x.dtype}")
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
