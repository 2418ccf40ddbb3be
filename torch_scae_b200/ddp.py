"""Data-parallel training plumbing: one process per GPU, gradients averaged with ONE NCCL all-reduce per step.

Both likelihood kernels are independent per image, so the batch shards across ranks with no data-path collective
(SURVEY.md section 8e); the only exchange is the parameter-gradient all-reduce.  All ``.grad`` tensors are views
into one flat fp32 bucket (3.5 M floats = 14 MB for the MNIST model), so the collective needs no packing copies and
NCCL moves it over NVLink/NVSwitch in a single call.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameter gradients of ``module`` in ONE flat fp32 buffer (what the all-reduce and the optimizer work on).

    ``assign=False`` (default): every ``p.grad`` is a view into the buffer; ``zero()`` clears it and autograd
    accumulates into the views in place -- one small add kernel per parameter (276 per step for the MNIST model).

    ``assign=True``: ``zero()`` only drops the ``.grad`` references (no kernel), so autograd hands every parameter its
    freshly computed gradient tensor without an add; ``collect()`` then copies them into the flat buffer with one
    multi-tensor launch and re-points ``p.grad`` at the views.  Parameters that received no gradient keep zeros.

    ``flat_params=True`` additionally re-homes every parameter as a view into one flat buffer (``flat_param``), which is
    what ``FlatRMSprop`` updates in a single kernel.  Build the bucket AFTER moving the module to its device.
    """

    def __init__(self, module, process_group=None, assign=False, flat_params=False):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        self.assign = assign
        # every parameter starts on a 16-byte boundary of the flat buffers (sizes rounded up to 4 floats): kernels that
        # take parameters as raw pointers (bulk copies, float4 loads) rely on that alignment
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, device=ref.device, dtype=ref.dtype)
        self.flat_param = None
        if flat_params:
            self.flat_param = torch.zeros(n, device=ref.device, dtype=ref.dtype)
        self.views = []
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(view)
            if flat_params:
                with torch.no_grad():
                    home = self.flat_param[off:off + p.numel()].view_as(p)
                    home.copy_(p)
                    p.data = home
            p.grad = None if assign else view      # autograd accumulates into the views in place
            off += (p.numel() + 3) // 4 * 4
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        if self.assign:
            for p in self.params:
                p.grad = None
        else:
            self.flat.zero_()

    def collect(self):
        """assign mode: gradients -> flat buffer (one multi-tensor copy); no-op otherwise."""
        if not self.assign:
            return
        src, dst, stale = [], [], []
        for p, view in zip(self.params, self.views):
            if p.grad is None:
                stale.append(view)                 # unused this step: its slot must not keep the previous gradient
            elif p.grad is not view:
                src.append(p.grad)
                dst.append(view)
        if src:
            torch._foreach_copy_(dst, src)
        if stale:
            # torch.optim skips parameters without a gradient; the flat optimizer and the all-reduce see the whole
            # buffer, so such slots must not keep the previous step's gradient.  Single process: they are marked NaN, which
            # FlatRMSprop treats as "skip" (parameter, square average and momentum untouched, exactly like torch.optim);
            # several ranks: zero, the value DDP reduces for an unused parameter.  One multi-tensor launch, graph-safe.
            torch._foreach_zero_(stale)
            if self.world == 1:
                torch._foreach_add_(stale, float('nan'))
        for p, view in zip(self.params, self.views):
            if p.grad is not None:
                p.grad = view                      # readers of .grad see the (reduced) bucket contents

    def all_reduce_mean(self):
        """Averages gradients across ranks (what Lightning's DDP does for the reference, train.py:40)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)

    def check_views(self):
        """True iff every existing .grad aliases the bucket (accumulate mode: autograd kept accumulating in place;
        assign mode: ``collect`` ran after the last backward)."""
        base = self.flat.untyped_storage().data_ptr()
        if self.assign:
            return all(p.grad is None or p.grad.untyped_storage().data_ptr() == base for p in self.params)
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


class FlatRMSprop:
    """``torch.optim.RMSprop(lr, alpha, eps, momentum)`` (centered = False, weight_decay = 0 -- the reference's
    optimizer, base_experiment.py:47-53) over a ``FlatGradBucket(flat_params=True)``: parameters, gradients and both
    state buffers are flat, so a step is ONE kernel (csrc/support.cu::rmsprop_kernel) instead of seven multi-tensor
    launches over 276 tensors.  Graph-capturable (no host-side state).
    """

    def __init__(self, bucket, lr=1e-2, alpha=0.99, eps=1e-8, momentum=0.0):
        if bucket.flat_param is None:
            raise ValueError('FlatRMSprop needs FlatGradBucket(..., flat_params=True)')
        self.bucket = bucket
        self.lr, self.alpha, self.eps, self.momentum = float(lr), float(alpha), float(eps), float(momentum)
        self.square_avg = torch.zeros_like(bucket.flat)
        self.momentum_buffer = torch.zeros_like(bucket.flat) if momentum > 0 else None
        # same attribute torch optimizers expose; GraphedTrainStep snapshots / restores these tensors around warm-up
        self.state = {0: dict(square_avg=self.square_avg, **(
            dict(momentum_buffer=self.momentum_buffer) if self.momentum_buffer is not None else {}))}

    def step(self):
        from . import _lib, ops
        lib = _lib.load()
        b = self.bucket
        if not b.flat.is_cuda:
            # host-side logic tests (gloo): the same arithmetic with torch ops
            live = ~torch.isnan(b.flat)                          # NaN = no gradient this step (FlatGradBucket.collect)
            g = torch.where(live, b.flat, torch.zeros_like(b.flat))
            sq = self.square_avg * self.alpha + (1 - self.alpha) * g * g
            self.square_avg.copy_(torch.where(live, sq, self.square_avg))
            step = g / (self.square_avg.sqrt() + self.eps)
            if self.momentum_buffer is not None:
                buf = self.momentum_buffer * self.momentum + step
                self.momentum_buffer.copy_(torch.where(live, buf, self.momentum_buffer))
                step = self.momentum_buffer
            b.flat_param.add_(torch.where(live, step, torch.zeros_like(step)), alpha=-self.lr)
            return
        _lib.check(ops._timed('scae_rmsprop_step', lib.scae_rmsprop_step, _lib.ptr(b.flat_param), _lib.ptr(b.flat),
                              _lib.ptr(self.square_avg), _lib.ptr(self.momentum_buffer), b.flat.numel(), self.lr,
                              self.alpha, self.eps, self.momentum, ops._stream()), 'scae_rmsprop_step')

    def zero_grad(self, set_to_none=True):
        self.bucket.zero()


def broadcast_parameters(module, src=0, process_group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src, group=process_group)


# ---- batch-global statistics under data parallelism ---------------------------------------------
class _AllReduceSum(torch.autograd.Function):
    """sum over ranks, differentiable: d(total)/d(local) = 1, and every rank holds the same upstream gradient."""

    @staticmethod
    def forward(ctx, local, group):
        total = local.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, factor):
        ctx.factor = factor
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.factor, None


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def global_sum(local, group=None):
    """``local`` summed over all ranks (a no-op on one rank); gradients flow back to the local term."""
    return _AllReduceSum.apply(local, group) if world_size(group) > 1 else local


def global_stat_loss(term, group=None):
    """A loss term computed from ``global_sum`` statistics is the SAME number on every rank; averaging the parameter
    gradients over ranks (FlatGradBucket.all_reduce_mean) would therefore weigh it 1/world.  This keeps the value and
    multiplies its gradient by the world size, so that the averaged gradients equal the single-process gradients of
    the global-batch loss."""
    w = world_size(group)
    return _ScaleGrad.apply(term, float(w)) if w > 1 else term
