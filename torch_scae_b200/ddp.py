"""Data-parallel training plumbing: one process per GPU, gradients averaged with ONE NCCL all-reduce per step.

Both likelihood kernels are independent per image, so the batch shards across ranks with no data-path collective
(SURVEY.md section 8e); the only exchange is the parameter-gradient all-reduce.  All ``.grad`` tensors are views
into one flat fp32 bucket (3.5 M floats = 14 MB for the MNIST model), so the collective needs no packing copies and
NCCL moves it over NVLink/NVSwitch in a single call.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, device=ref.device, dtype=ref.dtype)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)      # autograd accumulates into these views in place
            off += p.numel()
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """Averages gradients across ranks (what Lightning's DDP does for the reference, train.py:40)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)

    def check_views(self):
        """True iff every .grad still aliases the bucket (autograd kept accumulating in place)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


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
