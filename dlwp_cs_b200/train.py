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

    def collect_grads(self):
        """Gather freshly produced per-parameter gradients (``p.grad`` tensors that are not views of the flat buffer) into
        the flat buffer with one multi-tensor copy and re-point ``p.grad`` at the views.  Used by the trainer, which clears
        ``p.grad`` before backward so that autograd hands the gradients over instead of launching one accumulation kernel
        per parameter."""
        src, dst = [], []
        for p, v in zip(self.params, self.grad_views):
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

    def all_reduce(self, group=None):
        """Sum the flat gradient buffer over the ranks: the path's single collective.  Returns the world size."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def broadcast_params(self, src=0, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.param, src=src, group=group)


def shard_batch(global_batch, rank, world):
    """Contiguous shard [lo, hi) of a global batch for one rank (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallelTrainer(object):
    """
        trainer = DataParallelTrainer(model, lr=1e-3)
        loss = trainer.step(x_shard, target_shard)          # one optimizer step; loss is a device scalar (mean over shard)

    Keras Adam defaults (lr 1e-3, beta 0.9 / 0.999, eps 1e-7: train_cs.py:424 passes 'adam').  The loss of a rank is the
    mean over its shard; summing gradients and scaling by 1/world gives the gradient of the global-batch mean when the
    shards are equal (the reference scales the batch by the GPU count: Azure/train_tf.py:166).
    """

    def __init__(self, model, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, group=None, use_graph=True):
        self.model = model
        self.use_graph = use_graph
        self._graph = None
        self._static = None
        self.flat = FlatBuffers(model)
        if self.flat.param.device.type != 'cuda':
            raise _lib.DlwpcsError('DataParallelTrainer needs the model on a CUDA device (no CPU path)')
        self.m = torch.zeros_like(self.flat.param)
        self.v = torch.zeros_like(self.flat.param)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.group = group
        self.t = 0
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=self.flat.param.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.flat.param.device)
        self.flat.broadcast_params(group=group)

    def forward_backward(self, x, target):
        """Gradients of mean((model(x) - target)^2) accumulated into the flat buffer; returns the loss (device scalar)."""
        self.loss.zero_()
        for p in self.flat.params:          # every parameter is used exactly once: take the gradients, do not accumulate
            p.grad = None
        y = self.model(x)
        dy = _lib.mse_loss_grad(y.detach(), target, self.loss)
        y.backward(dy)
        self.flat.collect_grads()
        return self.loss

    def _step_body(self, x, target):
        loss = self.forward_backward(x, target)
        world = self.flat.all_reduce(self.group)
        # the step counter lives on the device (incremented by the call), so the same launches serve every step
        _lib.adam_step_dev(self.flat.param, self.flat.grad, self.m, self.v, self.lr, self.beta1, self.beta2, self.eps,
                           self.step_counter, 1.0 / world)
        return loss

    def step(self, x, target):
        """One optimizer step.  With ``use_graph`` the whole step (forward, loss, backward, all-reduce, Adam: ~150
        launches) is captured once per input shape into a CUDA graph and replayed; x / target are copied into the
        graph's static input buffers."""
        self.t += 1
        if not self.use_graph:
            return self._step_body(x, target)
        key = (tuple(x.shape), x.dtype, tuple(target.shape), target.dtype)
        if self._graph is None or self._static[0] != key:
            sx, st = torch.empty_like(x), torch.empty_like(target)
            sx.copy_(x)
            st.copy_(target)
            state = [t.clone() for t in (self.flat.param, self.m, self.v, self.step_counter)]
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):                   # warm-up off the capture: lazy tables, autograd buffers
                for _ in range(2):
                    self._step_body(sx, st)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            for dst, src in zip((self.flat.param, self.m, self.v, self.step_counter), state):
                dst.copy_(src)                              # the warm-up steps do not count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body(sx, st)
            for dst, src in zip((self.flat.param, self.m, self.v, self.step_counter), state):
                dst.copy_(src)                              # neither does the capture (nothing ran, but be explicit)
            self._graph, self._static = g, (key, sx, st)
        _, sx, st = self._static
        sx.copy_(x, non_blocking=True)
        st.copy_(target, non_blocking=True)
        self._graph.replay()
        return self.loss
