"""Linear layers on tall-skinny activations: split-K weight gradients.

The set transformer applies 16-wide ``nn.Linear`` layers to B*M = 40960 rows.  Autograd's weight gradient for such a
layer is a (out x rows) @ (rows x in) GEMM with a 16x16 .. 16x144 result: cuBLAS runs it as ONE 64x64-tile CTA looping
over 40960 rows (~50 us per layer, 24 layers = 7.5 % of the train step, profiles/r01c).  ``linear`` below keeps
``F.linear``'s forward and input gradient but computes the weight gradient as a batched GEMM over S row-chunks followed
by a sum over the chunks (split-K), which spreads the reduction over S CTAs.  Same fp32 arithmetic, different summation
order.  Pure PyTorch plumbing for the callers of the hot paths; parameters and module structure are untouched.
"""
import torch
import torch.nn.functional as F


def _chunks(rows, target=128, min_rows=128):
    """largest divisor of `rows` that is <= target and leaves at least `min_rows` rows per chunk (1 if none)."""
    best = 1
    for s in range(2, target + 1):
        if rows % s == 0 and rows // s >= min_rows:
            best = s
    return best


class _SplitKLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        g2 = g.reshape(-1, g.shape[-1])
        if ctx.needs_input_grad[0]:
            gx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, x.shape[-1])
            rows = x2.shape[0]
            s = _chunks(rows)
            if s > 1:
                gw = torch.bmm(g2.view(s, rows // s, -1).transpose(1, 2), x2.view(s, rows // s, -1)).sum(0)
            else:
                gw = g2.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import ops                      # column sums at the HBM rate (csrc/api.cu::colsum_kernel)
            gb = ops.colsum(g2)
        return gx, gw, gb


def linear(x, layer, max_weight_elems=256 * 256, min_rows=8192):
    """``layer(x)`` for an nn.Linear; uses the split-K weight gradient when the layer is small and x is tall."""
    rows = x.numel() // x.shape[-1]
    if layer.weight.numel() <= max_weight_elems and rows >= min_rows and torch.is_grad_enabled() and x.is_cuda:
        return _SplitKLinear.apply(x, layer.weight, layer.bias)
    return layer(x)
