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
        for p in params:
            n = p.numel()
            self.param[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.param[off:off + n].view(p.shape)
            p.grad = self.grad[off:off + n].view(p.shape)
            off += n
        self.params = params

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

    def __init__(self, model, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, group=None):
        self.model = model
        self.flat = FlatBuffers(model)
        if self.flat.param.device.type != 'cuda':
            raise _lib.DlwpcsError('DataParallelTrainer needs the model on a CUDA device (no CPU path)')
        self.m = torch.zeros_like(self.flat.param)
        self.v = torch.zeros_like(self.flat.param)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.group = group
        self.t = 0
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.flat.param.device)
        self.flat.broadcast_params(group=group)

    def forward_backward(self, x, target):
        """Gradients of mean((model(x) - target)^2) accumulated into the flat buffer; returns the loss (device scalar)."""
        self.flat.zero_grad()
        self.loss.zero_()
        y = self.model(x)
        dy = _lib.mse_loss_grad(y.detach(), target, self.loss)
        y.backward(dy)
        return self.loss

    def step(self, x, target):
        loss = self.forward_backward(x, target)
        world = self.flat.all_reduce(self.group)
        self.t += 1
        _lib.adam_step(self.flat.param, self.flat.grad, self.m, self.v, self.lr, self.beta1, self.beta2, self.eps, self.t,
                       1.0 / world)
        return loss
