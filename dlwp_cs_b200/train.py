"""
Data-parallel training step for the cubed-sphere U-Net: what `DLWPFunctional.fit` / `DLWPTorchNN.fit` do per batch
(reference DLWP/model/models.py:380-406, models_torch.py:253-263: forward, 'mse' loss, backward, Adam), with the
multi-GPU scheme of SURVEY.md section 8(e) in place of Keras `multi_gpu_model` (models.py:105-110): every rank holds the
whole 0.68 M-parameter replica and a shard of the batch; all parameters and all gradients live in ONE flat float32
buffer each, so the only exchange step of the path is a single `all_reduce(sum)` over NCCL, followed by one fused
kernel that applies the 1/world scale and the Adam update (`dlwpcs_adam_step`).

  * ``FlatBuffers``          -- re-homes a module's parameters / gradients into two flat buffers (device agnostic: the
    bucket logic is exercised on CPU with gloo in tests/test_train_cpu.py).
  * ``DataParallelTrainer``  -- forward through the differentiable cubed-sphere operators, MSE through
    `dlwpcs_mse_loss_grad`, backward (dgrad with halo scatter-add, wgrad), all-reduce, fused Adam.
"""
import torch
import torch.distributed as dist

from . import _lib
from . import functional as F_cs


class FlatBuffers(object):
    """One contiguous float32 buffer for all parameters and one for all gradients; the module's ``p.data`` / ``p.grad``
    become views into them (same order as ``module.parameters()``, i.e. the reference's ``add_weight`` order per layer,
    custom.py:882-914)."""

    def __init__(self, module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError('module has no trainable parameters')
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError('FlatBuffers needs float32 parameters on one device')
        self.sizes = [p.numel() for p in params]
        self.count = sum(self.sizes)
        self.param = torch.empty(self.count, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.count, dtype=torch.float32, device=dev)
        off = 0
        self.grad_views = []
        for p in params:
            n = p.numel()
            self.param[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.param[off:off + n].view(p.shape)
            self.grad_views.append(self.grad[off:off + n].view(p.shape))
            p.grad = self.grad_views[-1]
            off += n
        self.params = params

    def collect_grads(self, stop=None):
        """Gather freshly produced per-parameter gradients (``p.grad`` tensors that are not views of the flat buffer) into
        the flat buffer with one multi-tensor copy and re-point ``p.grad`` at the views.  Used by the trainer, which clears
        ``p.grad`` before backward so that autograd hands the gradients over instead of launching one accumulation kernel
        per parameter.  Parameters from index `stop` on are only re-pointed (their gradients were already copied -- and
        may be in flight in a collective)."""
        src, dst = [], []
        for i, (p, v) in enumerate(zip(self.params, self.grad_views)):
            if stop is not None and i >= stop:
                p.grad = v
                continue
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce(self, group=None, lo=0, hi=None):
        """Sum the flat gradient buffer (or its element range [lo, hi)) over the ranks.  Returns the world size."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            g = self.grad if (lo == 0 and hi is None) else self.grad[lo:hi]
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def bucket_split(self, min_fraction=0.3):
        """(parameter index, element offset) where the tail of the flat buffer -- the parameters whose gradients a backward
        pass produces FIRST (the last layers) -- holds at least `min_fraction` of all elements, cut at a parameter
        boundary that is a multiple of 4 elements (16-byte aligned float32 ranges for the collective)."""
        cuts = self.bucket_cuts((min_fraction,))
        return cuts[0] if cuts else (0, 0)

    def bucket_cuts(self, fractions=(0.3, 0.85)):
        """Cut points [(parameter index, element offset), ...] walking the flat buffer from its END (reverse layer order =
        the order in which backward completes the gradients): a cut is made when the elements behind it reach the next
        cumulative fraction.  n cuts give n + 1 buckets; the head bucket (the first layers) is what is left."""
        cuts, tail, want = [], 0, list(fractions)
        for i in range(len(self.sizes) - 1, 0, -1):
            tail += self.sizes[i]
            if want and tail >= want[0] * self.count and (self.count - tail) % 4 == 0:
                cuts.append((i, self.count - tail))
                want.pop(0)
        return cuts

    def broadcast_params(self, src=0, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.param, src=src, group=group)


def shard_batch(global_batch, rank, world):
    """Contiguous shard [lo, hi) of a global batch for one rank (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def repack_reference(y, solar, x_first, t_in, n_var):
    """Input of forecast step s >= 1 of the reference's multi-step model (`complete_model`, Azure/train_cs.py:391-409): per
    input time step the n_var variables just predicted, then that time step's insolation; the constants are the trailing
    channels of the first input.  y (B,6,N,N,t_in*n_var), solar (B,t_in,6,N,N,1), x_first (B,6,N,N,t_in*(n_var+1)+n_const)."""
    parts = []
    for t in range(t_in):
        parts.append(y[..., t * n_var:(t + 1) * n_var])
        parts.append(solar[:, t].to(y.dtype))
    parts.append(x_first[..., t_in * (n_var + 1):])
    return torch.cat(parts, dim=-1)


class DataParallelTrainer(object):
    """
        trainer = DataParallelTrainer(model, lr=1e-3)
        loss = trainer.step(x_shard, target_shard)          # one optimizer step; loss is a device scalar (mean over shard)

    Keras Adam defaults (lr 1e-3, beta 0.9 / 0.999, eps 1e-7: train_cs.py:424 passes 'adam').  The loss of a rank is the
    mean over its shard; summing gradients and scaling by 1/world gives the gradient of the global-batch mean when the
    shards are equal (the reference scales the batch by the GPU count: Azure/train_tf.py:166).
    """

    def __init__(self, model, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, group=None, use_graph=True, distributed=True,
                 overlap=True):
        self.model = model
        self.distributed = bool(distributed)      # False: a single-rank step even inside an initialised process group
        self.use_graph = use_graph
        self.prepack = True            # bf16: pack the weight images of all layers in one launch per step
        self._packed_state = None
        self._graphs = {}
        self.flat = FlatBuffers(model)
        if self.flat.param.device.type != 'cuda':
            raise _lib.DlwpcsError('DataParallelTrainer needs the model on a CUDA device (no CPU path)')
        self.m = torch.zeros_like(self.flat.param)
        self.v = torch.zeros_like(self.flat.param)
        self._hyper = dict(lr=float(lr), beta1=float(beta1), beta2=float(beta2), eps=float(eps))
        self.group = group
        self.t = 0
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=self.flat.param.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.flat.param.device)
        if self.distributed:
            self.flat.broadcast_params(group=group)
        # Gradient sink: the layers' backward writes the weight gradients straight into the flat buffer (functional.py) and
        # reports every delivered layer.  With world > 1 the flat buffer is cut in two buckets in reverse layer order: when
        # the tail bucket (the decoder layers, whose wgrad runs first) is complete its all-reduce is forked onto a side
        # stream -- inside the captured graph -- while the encoder's dgrad / wgrad kernels still run; the head bucket
        # follows on the main stream after backward and Adam waits for both.
        self._view_of = {p.data_ptr(): i for i, p in enumerate(self.flat.params)}
        self._overlap = False
        world = dist.get_world_size(group) if (self.distributed and dist.is_available() and dist.is_initialized()) else 1
        self._cuts = []
        if overlap and world > 1:
            self._cuts = self.flat.bucket_cuts()         # [(k_A, off_A), (k_B, off_B)]: tail bucket first
            if self._cuts:
                self._overlap = True
                self._side = torch.cuda.Stream(device=self.flat.param.device)
        import inspect
        self._keep_pad = 'keep_pad' in inspect.signature(model.forward).parameters
        self._delivered = set()
        self._fired = 0            # buckets (from the tail) whose all-reduce has been forked in this backward
        self._main = None

    # ---- gradient sink protocol (dlwp_cs_b200.functional.set_grad_sink) --------------------------------------------------
    def views(self, params):
        idx = [None if p is None else self._view_of.get(p.data_ptr()) for p in params]
        if any(p is not None and i is None for p, i in zip(params, idx)):
            return None                              # not (all) parameters of this trainer's model
        return tuple(None if i is None else self.flat.grad_views[i] for i in idx)

    def delivered(self, params):
        """Called from the layer's backward (autograd thread) once its weight gradients sit in the flat buffer."""
        for p in params:
            if p is not None:
                self._delivered.add(self._view_of[p.data_ptr()])
        while self._overlap and self._fired < len(self._cuts):
            k, off = self._cuts[self._fired]
            k_hi, off_hi = (len(self.flat.params), self.flat.count) if self._fired == 0 else self._cuts[self._fired - 1]
            if not all(i in self._delivered for i in range(k, k_hi)):
                break
            self._fired += 1
            main = self._main          # the stream the step is enqueued on (the autograd thread may have another current)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.flat.all_reduce(self.group, lo=off, hi=off_hi)

    # lr / beta1 / beta2 / eps are baked into the captured launches as scalars: changing one (a learning-rate schedule,
    # models_torch.py:297-298) drops the cached graphs so that the next step re-captures with the new value
    def _set_hyper(self, name, value):
        if self._hyper[name] != float(value):
            self._hyper[name] = float(value)
            self._graphs.clear()

    lr = property(lambda self: self._hyper['lr'], lambda self, v: self._set_hyper('lr', v))
    beta1 = property(lambda self: self._hyper['beta1'], lambda self, v: self._set_hyper('beta1', v))
    beta2 = property(lambda self: self._hyper['beta2'], lambda self, v: self._set_hyper('beta2', v))
    eps = property(lambda self: self._hyper['eps'], lambda self, v: self._set_hyper('eps', v))

    def close(self):
        """Release the captured CUDA graphs (they hold NCCL kernels of the communicator): call before
        ``torch.distributed.destroy_process_group()``."""
        torch.cuda.synchronize(self.flat.param.device)
        self._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize(self.flat.param.device)

    def _backward_into_flat(self, outputs, grads):
        """Backward with this trainer as the gradient sink; whatever autograd still hands over as ``p.grad`` (layers outside
        the sink protocol) is gathered afterwards.  Returns the number of buckets whose all-reduce is already in flight."""
        from . import functional as F_cs
        for p in self.flat.params:          # every parameter is used by one layer: take the gradients, do not accumulate
            p.grad = None
        self._delivered.clear()
        self._fired = 0
        self._main = torch.cuda.current_stream(self.flat.param.device)
        prev = F_cs.set_grad_sink(self)
        try:
            torch.autograd.backward(outputs, grads)
        finally:
            F_cs.set_grad_sink(prev)
        src, dst = [], []
        for i, (p, v) in enumerate(zip(self.flat.params, self.flat.grad_views)):
            if i not in self._delivered:
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    if self._fired and i >= self._cuts[self._fired - 1][0]:
                        raise _lib.DlwpcsError('gradient of a parameter arrived after the all-reduce of its bucket started')
                    src.append(p.grad)
                    dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)
        return self._fired

    def forward_backward(self, x, target):
        """Gradients of mean((model(x) - target)^2) in the flat buffer; returns the loss (device scalar)."""
        self.loss.zero_()
        y, target, scale = self._forward_padded(x, target)
        dy = _lib.mse_loss_grad(y.detach(), target, self.loss, scale=scale)
        self._tail_in_flight = self._backward_into_flat([y], [dy])
        return self.loss

    def _forward_padded(self, x, target):
        """Forward pass; for a model that supports it the output keeps its zero pad channels (bf16: channel counts are
        multiples of 8) and the target is zero-padded to match -- the pad channels contribute exactly zero to the squared
        error, the mean is taken over the real element count through `scale`, and the slice / re-pad copies around the
        loss disappear.  -> (y, target, loss scale)"""
        if not self._keep_pad:
            return self.model(x), target, 1.0
        y = self.model(x, keep_pad=True)
        extra = y.shape[-1] - target.shape[-1]
        if extra == 0:
            return y, target, 1.0
        target = torch.nn.functional.pad(target, (0, extra))
        return y, target, float(y.shape[-1]) / float(y.shape[-1] - extra)

    def _reduce_and_update(self):
        world = 1
        if self.distributed:
            if self._tail_in_flight:           # the buckets behind this offset are being reduced on the side stream
                world = self.flat.all_reduce(self.group, lo=0, hi=self._cuts[self._tail_in_flight - 1][1])
                self._main.wait_stream(self._side)
            else:
                world = self.flat.all_reduce(self.group)
        # the step counter lives on the device (incremented by the call), so the same launches serve every step
        _lib.adam_step_dev(self.flat.param, self.flat.grad, self.m, self.v, self.lr, self.beta1, self.beta2, self.eps,
                           self.step_counter, 1.0 / world)

    def _prepack(self, x):
        """bf16: the weight images of every layer in ONE launch at the start of the step (instead of one or two launches
        per layer inside the layers' forward)."""
        if x.dtype == torch.bfloat16 and self.prepack:
            self._packed_state = F_cs.prepack_model(self.model, torch.bfloat16, state=self._packed_state)

    def _step_body(self, x, target):
        self._prepack(x)
        try:
            loss = self.forward_backward(x, target)
        finally:
            F_cs.clear_prepacked()
        self._reduce_and_update()
        return loss

    def _graphed(self, tag, tensors, body):
        """Run body(*tensors) -- one whole optimizer step -- from a CUDA graph captured once per (tag, shapes, dtypes); the
        tensors are copied into the graph's static input buffers before every replay."""
        if not self.use_graph:
            return body(*tensors)
        key = (tag,) + tuple((tuple(t.shape), t.dtype) for t in tensors)
        if key not in self._graphs:
            dev = tensors[0].device
            statics = [torch.empty_like(t) for t in tensors]
            for s_, t in zip(statics, tensors):
                s_.copy_(t)
            keep = (self.flat.param, self.m, self.v, self.step_counter)
            state = [t.clone() for t in keep]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                   # warm-up off the capture: lazy tables, autograd buffers
                for _ in range(2):
                    body(*statics)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for dst, src in zip(keep, state):
                dst.copy_(src)                              # the warm-up steps do not count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body(*statics)
            for dst, src in zip(keep, state):
                dst.copy_(src)                              # neither does the capture (nothing ran, but be explicit)
            self._graphs[key] = (g, statics)
        g, statics = self._graphs[key]
        for s_, t in zip(statics, tensors):
            s_.copy_(t, non_blocking=True)
        g.replay()
        return self.loss

    def step(self, x, target):
        """One optimizer step.  With ``use_graph`` the whole step (forward, loss, backward, all-reduce, Adam: ~150
        launches) is captured once per input shape into a CUDA graph and replayed.

        Returns the trainer's persistent 1-element loss buffer (no allocation, no synchronisation): it is overwritten by
        the next step -- ``.clone()`` or ``.item()`` it to keep a history.  Every rank must hold the same number of
        samples (the gradient of the global-batch mean is the 1/world-scaled sum of the shard means, train_tf.py:166);
        ``shard_batch`` gives equal shards whenever the global batch is a multiple of the world size."""
        self.t += 1
        return self._graphed('step', [x, target], self._step_body)

    # ---- multi-step ("integration_steps") training of the reference: Azure/train_cs.py:101, 391-426 --------------------
    def _sequence_body(self, t_in, n_var, n_steps, x, *rest):
        solars, targets = rest[:n_steps - 1], rest[n_steps - 1:]
        self.loss.zero_()
        ys, xin = [], x
        self._prepack(x)
        try:
            for s in range(n_steps):                 # the same layer objects (shared weights) at every step
                y = self.model(xin)
                ys.append(y)
                if s + 1 < n_steps:
                    xin = repack_reference(y, solars[s], x, t_in, n_var)
        finally:
            F_cs.clear_prepacked()
        # keras multi-output loss: sum_s w_s * mse(y_s, t_s) with w_s = 1/S (train_cs.py:424-426)
        dys = [_lib.mse_loss_grad(y.detach(), t, self.loss, scale=1.0 / n_steps) for y, t in zip(ys, targets)]
        # shared layers: every parameter receives one gradient per model step, so the sink (which overwrites) is bypassed
        # and autograd accumulates them (train_cs.py:391-409 applies the same layer objects S times)
        for p in self.flat.params:
            p.grad = None
        torch.autograd.backward(ys, dys)
        self.flat.collect_grads()
        self._tail_in_flight = 0
        self._reduce_and_update()
        return self.loss

    def step_sequence(self, x, solars, targets, t_in, n_var):
        """One optimizer step of the S-step model: x = input of step 0 in the reference's packing (per input time step the
        n_var variables then its insolation, constants last), solars[s-1] (B,t_in,6,N,N,1) = insolation of step s, targets[s]
        = truth of step s -- exactly what ``DeviceDataFeed(..., sequence=S).generate`` returns.  Needs t_out == t_in."""
        n_steps = len(targets)
        if len(solars) != n_steps - 1:
            raise ValueError('%d targets need %d insolation tensors, got %d' % (n_steps, n_steps - 1, len(solars)))
        self.t += 1
        body = lambda *ts: self._sequence_body(t_in, n_var, n_steps, *ts)
        return self._graphed(('sequence', t_in, n_var, n_steps), [x] + list(solars) + list(targets), body)
