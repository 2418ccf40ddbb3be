"""``torch.autograd.Function`` wrappers around the C-ABI kernels (the only callers of libscae_b200.so).

PyTorch is plumbing here: it owns device memory and the stream; the arithmetic of both hot paths happens in
csrc/tmpl_ll.cu and csrc/caps_ll.cu.  All tensors crossing the boundary are made contiguous fp32 first.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (CAPS_OUTPUT_FIELDS, CAPS_UPSTREAM_FIELDS, CapsArgs, CapsExplicitArgs, CapsOutputs, CapsSaved,
                   CapsUpstream, TmplArgs, check, ptr)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelTimer:
    """Optional per-entry-point device timing (CUDA events on the launching stream) and launch counting.

    ``with ops.KernelTimer() as t: ...`` records an event pair around every C-ABI call made inside the block;
    ``t.summary()`` (after a synchronize) returns {entry point: (calls, kernel launches, total ms)}.  bench.py uses
    it for the roofline numbers; it is off (zero overhead) otherwise.
    """
    active = None

    def __init__(self):
        self.records = {}

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        out = {}
        for name, recs in self.records.items():
            out[name] = (len(recs), sum(r[2] for r in recs), sum(r[0].elapsed_time(r[1]) for r in recs))
        return out

    def medians(self):
        """{entry point: median ms per call}: an event pair also spans whatever the host did between the two records, so
        one late launch (a busy host core) inflates the mean of a 40 us kernel; the median does not move."""
        out = {}
        for name, recs in self.records.items():
            ms = sorted(r[0].elapsed_time(r[1]) for r in recs)
            out[name] = ms[len(ms) // 2]
        return out


def _timed(name, fn, *args):
    """Calls a C-ABI entry point, bracketing it with CUDA events when a KernelTimer is active; the number of kernels
    the call launched comes from the library's own counter (scae_launch_count)."""
    t = KernelTimer.active
    if t is None:
        return fn(*args)
    lib = _lib.load()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = lib.scae_launch_count()
    start.record()
    rc = fn(*args)
    end.record()
    t.records.setdefault(name, []).append((start, end, lib.scae_launch_count() - before))
    return rc


def _f32c(t):
    """contiguous fp32 CUDA tensor (no copy when it already is one); None passes through."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.ScaeError('torch_scae_b200 runs on CUDA only (no CPU fallback); got a CPU tensor')
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def colsum(x2d):
    """Sum over the rows of a contiguous fp32 CUDA matrix (rows, cols) -> (cols,) through scae_colsum: the bias gradient
    of the set transformer's tall-skinny linear layers.  Shapes the kernel does not cover use torch.sum (same values up
    to summation order; this is host-side plumbing, not one of the two hot paths)."""
    lib = _lib.load()
    rows, cols = x2d.shape
    ws_bytes = lib.scae_colsum_workspace_bytes(rows, cols) if x2d.is_cuda and x2d.dtype == torch.float32 else 0
    if ws_bytes == 0:
        return x2d.sum(0)
    x2d = x2d.contiguous()
    out = torch.empty(cols, device=x2d.device, dtype=torch.float32)
    ws = torch.empty(ws_bytes, device=x2d.device, dtype=torch.uint8)
    check(_timed('scae_colsum', lib.scae_colsum, ptr(x2d), rows, cols, ptr(out), ptr(ws), ws_bytes, _stream()),
          'scae_colsum')
    return out


# =================================================================================================================
# fused plumbing kernels for the callers of the hot paths (csrc/support.cu)
# =================================================================================================================
# These are NOT likelihood paths: each falls back to the stock PyTorch op when the kernel does not cover the input
# (CPU tensors, other dtypes / shapes), with identical semantics.

def _workspace(n_bytes, device):
    return torch.empty(max(int(n_bytes), 16), device=device, dtype=torch.uint8)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        lib = _lib.load()
        x = x.contiguous()
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        stats = torch.empty(rows, 2, device=x.device, dtype=torch.float32)
        check(_timed('scae_layernorm_fwd', lib.scae_layernorm_fwd, ptr(x), ptr(weight), ptr(bias), eps, rows, d,
                     ptr(y), ptr(stats), _stream()), 'scae_layernorm_fwd')
        ctx.save_for_backward(x, weight, stats)
        ctx.has = (weight is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, weight, stats = ctx.saved_tensors
        g = g.contiguous()
        d = x.shape[-1]
        rows = x.numel() // d
        gx = torch.empty_like(x)
        ggb = torch.empty(2 * d, device=x.device, dtype=torch.float32)
        ws_bytes = lib.scae_layernorm_bwd_workspace_bytes(rows, d)
        ws = _workspace(ws_bytes, x.device)
        check(_timed('scae_layernorm_bwd', lib.scae_layernorm_bwd, ptr(g), ptr(x), ptr(weight), ptr(stats), rows, d,
                     ptr(gx), ptr(ggb), ptr(ws), ws_bytes, _stream()), 'scae_layernorm_bwd')
        return gx, ggb[:d] if ctx.has[0] else None, ggb[d:] if ctx.has[1] else None, None


def layer_norm(x, weight, bias, eps):
    """F.layer_norm over the last dimension."""
    d = x.shape[-1]
    if x.is_cuda and x.dtype == torch.float32 and d in (8, 16, 32, 64) and x.numel() > 0 \
            and (weight is None or weight.dtype == torch.float32):
        return _LayerNorm.apply(x, weight, bias, float(eps))
    return torch.nn.functional.layer_norm(x, (d,), weight, bias, eps)


class _ConvBiasAct(torch.autograd.Function):
    """cuDNN convolution (no padding / dilation / groups) + per-channel bias + optional ReLU."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, relu):
        lib = _lib.load()
        if relu and os.environ.get('SCAE_B200_CUDNN_FUSED_RELU', '0') == '1':
            # experiment (off by default): cuDNN's own fused convolution + bias + ReLU forward, no separate pass
            y = torch.cudnn_convolution_relu(x, weight, bias, stride, (0, 0), (1, 1), 1).contiguous()
        else:
            y = torch.nn.functional.conv2d(x, weight, None, stride).contiguous()
            N, C, H, W = y.shape
            check(_timed('scae_bias_act_fwd', lib.scae_bias_act_fwd, ptr(y), ptr(bias), N, C, H * W, int(relu),
                         _stream()), 'scae_bias_act_fwd')
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.stride, ctx.relu = stride, relu
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, weight, y = ctx.saved_tensors
        g = g.contiguous()
        N, C, H, W = g.shape
        g_bias = torch.empty(C, device=g.device, dtype=torch.float32)
        gx_pre = torch.empty_like(g) if ctx.relu else g
        ws_bytes = lib.scae_bias_act_bwd_workspace_bytes(N, C, H * W)
        ws = _workspace(ws_bytes, g.device)
        check(_timed('scae_bias_act_bwd', lib.scae_bias_act_bwd, ptr(g), ptr(y), ptr(gx_pre) if ctx.relu else None,
                     ptr(g_bias), N, C, H * W, int(ctx.relu), ptr(ws), ws_bytes, _stream()), 'scae_bias_act_bwd')
        gx, gw, _ = torch.ops.aten.convolution_backward(
            gx_pre, x, weight, None, ctx.stride, (0, 0), (1, 1), False, (0, 0), 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        return gx, gw, g_bias if ctx.needs_input_grad[2] else None, None, None


def _transpose_batched(x, batch, R, Cc):
    """x viewed as [batch, R, Cc] (contiguous fp32) -> new tensor [batch, Cc, R] through scae_transpose_batched."""
    lib = _lib.load()
    out = torch.empty(batch, Cc, R, device=x.device, dtype=torch.float32)
    check(_timed('scae_transpose_batched', lib.scae_transpose_batched, ptr(x), ptr(out), batch, R, Cc, _stream()),
          'scae_transpose_batched')
    return out


class _NchwToRows(torch.autograd.Function):
    """(B, C, H, W) -> (B*H*W, C): the rows = positions matrix of a feature map, and back in the backward."""

    @staticmethod
    def forward(ctx, x):
        B, C, H, W = x.shape
        ctx.shape = (B, C, H, W)
        return _transpose_batched(x.contiguous(), B, C, H * W).view(B * H * W, C)

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        return _transpose_batched(g.contiguous(), B, H * W, C).view(B, C, H, W)


def _split_k_wgrad(g2d, cols, cap=32, min_rows=256):
    """g2d^T @ cols for tall operands (rows, C_out) / (rows, K): few output tiles and a very long reduction, so the
    rows are cut into up to ``cap`` chunks reduced by one batched GEMM and summed (60 vs 44 TFLOP/s on B200 for the
    stride-2 layer, tools/conv_gemm_probe.py)."""
    rows = g2d.shape[0]
    s = max((d for d in range(1, cap + 1) if rows % d == 0 and rows // d >= min_rows), default=1)
    if s > 1:
        return torch.bmm(g2d.view(s, rows // s, -1).transpose(1, 2), cols.view(s, rows // s, -1)).sum(0)
    return g2d.t() @ cols


class _Conv3x3Gemm(torch.autograd.Function):
    """``relu?(conv2d(x, weight, bias, stride))`` for an unpadded 3x3 convolution with GEMM-form passes (csrc/conv_cols.cu).

    ``full``: im2col + one cuBLAS SGEMM forward (bias and ReLU in its epilogue), weight gradient as a split-K GEMM on the
    saved columns.  Otherwise forward and weight gradient stay with cuDNN.  The data gradient is a GEMM + col2im either
    way -- cuDNN's fp32 data-gradient engine runs these layers at 24-25 TFLOP/s, the GEMM at 43-49."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, relu, full):
        lib = _lib.load()
        B, C, H, W = x.shape
        Co = weight.shape[0]
        Ho, Wo = (H - 3) // stride + 1, (W - 3) // stride + 1
        cols = None
        if full:
            cols = torch.empty(B * Ho * Wo, C * 9, device=x.device, dtype=torch.float32)
            check(_timed('scae_im2col3x3', lib.scae_im2col3x3, ptr(x), ptr(cols), B, C, H, W, stride, _stream()),
                  'scae_im2col3x3')
            w2d_t = weight.reshape(Co, C * 9).t()
            y2d = torch._addmm_activation(bias, cols, w2d_t) if relu else torch.addmm(bias, cols, w2d_t)
            y = _transpose_batched(y2d, B, Ho * Wo, Co).view(B, Co, Ho, Wo)
        elif relu and os.environ.get('SCAE_B200_CUDNN_FUSED_RELU', '0') == '1':
            y = torch.cudnn_convolution_relu(x, weight, bias, (stride, stride), (0, 0), (1, 1), 1).contiguous()
        else:
            y = torch.nn.functional.conv2d(x, weight, None, stride).contiguous()
            check(_timed('scae_bias_act_fwd', lib.scae_bias_act_fwd, ptr(y), ptr(bias), B, Co, Ho * Wo, int(relu),
                         _stream()), 'scae_bias_act_fwd')
        ctx.save_for_backward(cols if full else x, weight, y if relu else None)
        ctx.cfg = (stride, relu, full, (B, C, H, W))
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        saved, weight, y = ctx.saved_tensors
        stride, relu, full, (B, C, H, W) = ctx.cfg
        Co = weight.shape[0]
        g = g.contiguous()
        _, _, Ho, Wo = g.shape
        g_bias = torch.empty(Co, device=g.device, dtype=torch.float32)
        gx_pre = torch.empty_like(g) if relu else g
        ws_bytes = lib.scae_bias_act_bwd_workspace_bytes(B, Co, Ho * Wo)
        ws = _workspace(ws_bytes, g.device)
        check(_timed('scae_bias_act_bwd', lib.scae_bias_act_bwd, ptr(g), ptr(y), ptr(gx_pre) if relu else None,
                     ptr(g_bias), B, Co, Ho * Wo, int(relu), ptr(ws), ws_bytes, _stream()), 'scae_bias_act_bwd')
        g2d = _transpose_batched(gx_pre, B, Co, Ho * Wo).view(B * Ho * Wo, Co)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            dcols = g2d @ weight.reshape(Co, C * 9)
            gx = torch.empty(B, C, H, W, device=g.device, dtype=torch.float32)
            check(_timed('scae_col2im3x3', lib.scae_col2im3x3, ptr(dcols), ptr(gx), B, C, H, W, stride, _stream()),
                  'scae_col2im3x3')
        if ctx.needs_input_grad[1]:
            if full:
                gw = _split_k_wgrad(g2d, saved).view_as(weight)
            else:
                gw = torch.ops.aten.convolution_backward(gx_pre, saved, weight, None, (stride, stride), (0, 0), (1, 1),
                                                         False, (0, 0), 1, [False, True, False])[1]
        return gx, gw, g_bias if ctx.needs_input_grad[2] else None, None, None, None


def _conv_gemm_mode(x, conv):
    """None (cuDNN for everything), 'dgrad' or 'full' for an nn.Conv2d applied to x; SCAE_B200_CONV_GEMM=0|dgrad|full
    overrides the rule (A/B timing).  Rule from tools/conv_gemm_probe.py on B200: the data gradient as GEMM + col2im
    whenever the layer is wide enough for the GEMM to pay (C_in >= 32); forward and weight gradient too for strided
    layers, where cuDNN's fp32 kernels are ~30 TFLOP/s (its stride-1 Winograd forward / weight gradient are faster
    than any SIMT GEMM and stay)."""
    want = os.environ.get('SCAE_B200_CONV_GEMM', 'auto')
    if want == '0' or conv.kernel_size != (3, 3) or conv.stride[0] != conv.stride[1] or conv.stride[0] not in (1, 2):
        return None
    B, C, H, W = x.shape
    if C < 32 or conv.out_channels < 32 or not _lib.load().scae_conv_cols_supported(B, C, H, W, conv.stride[0]):
        return None
    if want in ('dgrad', 'full'):
        return want
    return 'full' if conv.stride[0] == 2 else 'dgrad'


def conv_bias_act(x, conv, relu):
    """``relu(conv(x))`` / ``conv(x)`` for an nn.Conv2d; the bias add, the ReLU and (backward) the ReLU mask + bias
    gradient run as one pass each instead of four."""
    ok = (x.is_cuda and x.dtype == torch.float32 and conv.bias is not None and conv.padding == (0, 0)
          and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == 'zeros' and x.dim() == 4
          and conv.weight.dtype == torch.float32)
    if not ok:
        y = conv(x)
        return torch.relu(y) if relu else y
    mode = _conv_gemm_mode(x, conv)
    if mode is not None:
        return _Conv3x3Gemm.apply(x.contiguous(), conv.weight, conv.bias, conv.stride[0], bool(relu), mode == 'full')
    return _ConvBiasAct.apply(x.contiguous(), conv.weight, conv.bias, tuple(conv.stride), bool(relu))


class _AttentionPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, groups, D, S):
        lib = _lib.load()
        h = h.contiguous()
        out = torch.empty(groups, D, device=h.device, dtype=torch.float32)
        check(_timed('scae_attnpool_fwd', lib.scae_attnpool_fwd, ptr(h), ptr(out), groups, D, S, _stream()),
              'scae_attnpool_fwd')
        ctx.save_for_backward(h)
        ctx.dims = (groups, D, S)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (h,) = ctx.saved_tensors
        groups, D, S = ctx.dims
        gh = torch.empty_like(h)
        g = g.contiguous()                   # (a named reference: the buffer must outlive the call)
        check(_timed('scae_attnpool_bwd', lib.scae_attnpool_bwd, ptr(h), ptr(g), ptr(gh), groups, D, S,
                     _stream()), 'scae_attnpool_bwd')
        return gh, None, None, None


def attention_pool(feature_map, n_attention_map):
    """nn_ext.multiple_attention_pooling_2d on a CUDA fp32 map with at most 64 positions; None when not covered."""
    B, C, H, W = feature_map.shape
    S, D = H * W, C // n_attention_map - 1
    if not (feature_map.is_cuda and feature_map.dtype == torch.float32 and S <= 64 and 0 < D <= 1024):
        return None
    return _AttentionPool.apply(feature_map, B * n_attention_map, D, S).view(B, C - n_attention_map, 1, 1)


class _AttentionPoolCL(torch.autograd.Function):
    """Attention pooling of a channels-last map y (B, S, n*(D+1)) -> (B*n, D) (csrc/attnpool_cl.cu)."""

    @staticmethod
    def forward(ctx, y, n, D):
        lib = _lib.load()
        y = y.contiguous()
        B, S, _ = y.shape
        out = torch.empty(B * n, D, device=y.device, dtype=torch.float32)
        check(_timed('scae_attnpool_cl_fwd', lib.scae_attnpool_cl_fwd, ptr(y), ptr(out), B * n, n, D, S, _stream()),
              'scae_attnpool_cl_fwd')
        ctx.save_for_backward(y)
        ctx.dims = (n, D)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        n, D = ctx.dims
        B, S, _ = y.shape
        gy = torch.empty_like(y)
        g = g.contiguous()
        check(_timed('scae_attnpool_cl_bwd', lib.scae_attnpool_cl_bwd, ptr(y), ptr(g), ptr(gy), B * n, n,
                     D, S, _stream()), 'scae_attnpool_cl_bwd')
        return gy, None, None


class _PositionsGemm(torch.autograd.Function):
    """y = x @ w^T for x (rows, in) tall and w (out, in): plain cuBLAS SGEMMs, the weight gradient split over row chunks
    (its (out x rows) @ (rows x in) product has few output tiles and a very long reduction)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = g @ w
        if ctx.needs_input_grad[1]:
            rows = x.shape[0]
            cap = int(os.environ.get('SCAE_B200_ATT_SPLITK', '32'))          # 1: leave the split to cuBLAS
            s = max((d for d in range(1, cap + 1) if rows % d == 0 and rows // d >= 256), default=1)
            if s > 1:
                gw = torch.bmm(g.view(s, rows // s, -1).transpose(1, 2), x.view(s, rows // s, -1)).sum(0)
            else:
                gw = g.t() @ x
        return gx, gw


def attention_conv_pool(feature_map, conv, n_caps):
    """``multiple_attention_pooling_2d(conv(feature_map), n_caps)`` for the part encoder's 1x1 attention convolution
    (part_encoder.py:95-101) as GEMM + channels-last pooling kernel + bias on the pooled result; (B, n*D, 1, 1), or None
    when the layer / shape is not covered (the caller then runs convolution and pooling separately)."""
    if not (feature_map.is_cuda and feature_map.dtype == torch.float32 and feature_map.dim() == 4
            and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is not None
            and conv.weight.dtype == torch.float32 and conv.out_channels % n_caps == 0
            and conv.out_channels // n_caps >= 2):
        return None
    B, Cin, H, W = feature_map.shape
    S, Ctot = H * W, conv.out_channels
    G = Ctot // n_caps
    if B == 0 or not _lib.load().scae_attnpool_cl_supported(B * n_caps, n_caps, G - 1, S):
        return None
    x2d = _NchwToRows.apply(feature_map)
    y = _PositionsGemm.apply(x2d, conv.weight.reshape(Ctot, Cin))
    pooled = _AttentionPoolCL.apply(y.view(B, S, Ctot), n_caps, G - 1).view(B, n_caps, G - 1)
    return (pooled + conv.bias.view(n_caps, G)[:, :-1]).reshape(B, n_caps * (G - 1), 1, 1)


class _PoseTransform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, similarity):
        lib = _lib.load()
        t = t.contiguous()
        out = torch.empty_like(t)
        rows = t.numel() // 6
        check(_timed('scae_pose_transform', lib.scae_pose_transform, ptr(t), None, ptr(out), rows, int(similarity),
                     _stream()), 'scae_pose_transform')
        ctx.save_for_backward(t)
        ctx.similarity = similarity
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (t,) = ctx.saved_tensors
        gt = torch.empty_like(t)
        g = g.contiguous()
        check(_timed('scae_pose_transform', lib.scae_pose_transform, ptr(t), ptr(g), ptr(gt),
                     t.numel() // 6, int(ctx.similarity), _stream()), 'scae_pose_transform')
        return gt, None


def pose_transform(pose_tensor, similarity):
    """cv_ops.geometric_transform(nonlinear=True, as_matrix=False) as one kernel per direction; None when not covered."""
    if not (pose_tensor.is_cuda and pose_tensor.dtype == torch.float32 and pose_tensor.shape[-1] == 6
            and pose_tensor.numel() > 0):
        return None
    return _PoseTransform.apply(pose_tensor, bool(similarity))


class _LossHead(torch.autograd.Function):
    """Sparsity losses + classifier cross-entropies of SCAE.loss on the outputs of hot path 2 (csrc/loss_head.cu)."""

    @staticmethod
    def forward(ctx, caps_presence, posterior, label, weight, bias, cfg, world=1):
        """``world`` > 1: the between-example statistics are those of the global batch -- the shard's column sums are
        all-reduced between the two halves of the forward (cfg's between_constant is then the global batch / n_classes)
        and the between terms' gradients are scaled by the world size (ddp.global_stat_loss)."""
        lib = _lib.load()
        caps_presence, posterior = caps_presence.contiguous(), posterior.contiguous()
        label = label.contiguous() if label is not None else None
        weight = weight.contiguous() if label is not None else None
        bias = bias.contiguous() if label is not None else None
        B, O, V = posterior.shape
        K = weight.shape[0] if weight is not None else 0
        dev = posterior.device
        args = _lib.LossHeadArgs(ptr(caps_presence), ptr(posterior), ptr(label), ptr(weight), ptr(bias), B, O, V, K,
                                 *cfg)
        terms = torch.empty(8, device=dev, dtype=torch.float32)
        stats = torch.empty(128, device=dev, dtype=torch.float32)
        probs = torch.empty(2, B, K, device=dev, dtype=torch.float32) if label is not None else None
        ws_bytes = lib.scae_loss_head_workspace_bytes(ctypes.byref(args))
        ws = _workspace(ws_bytes, dev)
        if world > 1:
            import torch.distributed as dist
            colsums = torch.empty(132, device=dev, dtype=torch.float32)
            check(_timed('scae_loss_head_fwd', lib.scae_loss_head_fwd_rows, ctypes.byref(args), ptr(probs), ptr(colsums),
                         ptr(ws), ws_bytes, _stream()), 'scae_loss_head_fwd_rows')
            dist.all_reduce(colsums[:128], op=dist.ReduceOp.SUM)     # both O-float column sums in one collective
            check(lib.scae_loss_head_fwd_finish(ctypes.byref(args), ptr(colsums), ptr(terms), ptr(stats), _stream()),
                  'scae_loss_head_fwd_finish')
            # averaged over ranks, a term every rank computes from global statistics would weigh 1 / world
            cfg = cfg[:4] + (cfg[4] * world,) + cfg[5:6] + (cfg[6] * world,) + cfg[7:]
        else:
            check(_timed('scae_loss_head_fwd', lib.scae_loss_head_fwd, ctypes.byref(args), ptr(terms), ptr(probs),
                         ptr(stats), ptr(ws), ws_bytes, _stream()), 'scae_loss_head_fwd')
        saved = [caps_presence, posterior, stats] + ([label, weight, bias] if label is not None else [])
        ctx.save_for_backward(*saved)
        ctx.cfg = cfg
        total = terms[6].clone()
        if probs is None:
            probs = terms.new_empty(0)
        ctx.mark_non_differentiable(terms, probs)
        return total, terms, probs

    @staticmethod
    def backward(ctx, g_total, _g_terms, _g_probs):
        lib = _lib.load()
        caps_presence, posterior, stats, *cls = ctx.saved_tensors
        label, weight, bias = cls if cls else (None, None, None)
        B, O, V = posterior.shape
        K = weight.shape[0] if weight is not None else 0
        dev = posterior.device
        cfg = ctx.cfg
        args = _lib.LossHeadArgs(ptr(caps_presence), ptr(posterior), ptr(label), ptr(weight), ptr(bias), B, O, V, K,
                                 *cfg)
        sparsity = bool(cfg[0])
        g_cp = torch.empty_like(caps_presence) if sparsity and ctx.needs_input_grad[0] else None
        g_post = torch.empty_like(posterior) if sparsity and ctx.needs_input_grad[1] else None
        g_cls = torch.empty(K * O + K, device=dev, dtype=torch.float32) if label is not None else None
        ws_bytes = lib.scae_loss_head_workspace_bytes(ctypes.byref(args))
        ws = _workspace(ws_bytes, dev)
        g_total = _f32c(g_total).reshape(1)
        check(_timed('scae_loss_head_bwd', lib.scae_loss_head_bwd, ctypes.byref(args), ptr(stats), ptr(g_total),
                     ptr(g_cp), ptr(g_post), ptr(g_cls), ptr(ws), ws_bytes, _stream()), 'scae_loss_head_bwd')
        g_w = g_cls[:K * O].view(K, O) if g_cls is not None else None
        g_b = g_cls[K * O:] if g_cls is not None else None
        return g_cp, g_post, None, g_w, g_b, None, None


LOSS_HEAD_TERMS = ('prior_within_sparsity_loss', 'prior_between_sparsity_loss', 'posterior_within_sparsity_loss',
                   'posterior_between_sparsity_loss', 'prior_cls_xe', 'posterior_cls_xe')


def loss_head(caps_presence, posterior, label, classifier, n_classes, prior_type, posterior_type, weights,
              prior_within_constant=None, sparsity=True, sync_batch_stats=False):
    """The (B,O)-sized tail of SCAE.loss (stacked_capsule_auto_encoder.py:243-285) in one kernel pair per direction:
    the prior sparsity loss on ``caps_presence`` (B,O), the posterior one on ``posterior.sum(-1) / V`` and, with
    ``label``, the cross-entropies of both classifier heads (``classifier`` = nn.Linear shared by both, sic :211; its
    inputs are detached as in the reference).  ``weights`` = (prior within, prior between, posterior within,
    posterior between).  Returns (total, terms (8,), class probabilities (2,B,K) | None) with total = the weighted sum
    SCAE.loss adds, or None when the kernels do not cover the request (the caller then runs the PyTorch ops).

    ``label`` values must lie in [0, n_classes): the kernel does no range check (a host-side check would be a device
    sync inside the captured step).  ``F.cross_entropy`` raises on an out-of-range label and leaves ``ignore_index``
    rows out of the mean; here such a row matches no class and yields a finite but meaningless term -- validate labels
    in the data pipeline."""
    if not (caps_presence.is_cuda and caps_presence.dtype == torch.float32 and posterior.dtype == torch.float32
            and posterior.dim() == 3 and tuple(caps_presence.shape) == tuple(posterior.shape[:2])
            and 0 < posterior.shape[1] <= 64 and posterior.shape[0] > 0 and posterior.shape[2] > 0):
        return None
    if not (sparsity or label is not None):
        return None
    B, O, V = posterior.shape
    if sparsity:
        if prior_type not in _lib.LOSS_TYPES or posterior_type not in _lib.LOSS_TYPES:
            return None
        if 'l2' in (prior_type, posterior_type) and not n_classes:
            return None                       # the reference fails on the division; let the PyTorch path do that
    weight = bias = None
    if label is not None:
        weight, bias = classifier.weight, classifier.bias
        if not (bias is not None and weight.dtype == torch.float32 and 0 < weight.shape[0] <= 16
                and weight.shape[1] == O and label.dtype == torch.int64 and tuple(label.shape) == (B,)):
            return None
    default_c = float(O) / n_classes if n_classes else 0.0
    world = 1
    if sync_batch_stats:
        from . import ddp
        world = ddp.world_size()
    cfg = (int(bool(sparsity)), _lib.LOSS_TYPES.get(prior_type, 0), _lib.LOSS_TYPES.get(posterior_type, 0),
           float(weights[0]), float(weights[1]), float(weights[2]), float(weights[3]),
           float(default_c if prior_within_constant is None else prior_within_constant), float(default_c),
           float(B * world) / n_classes if n_classes else 0.0)
    total, terms, probs = _LossHead.apply(caps_presence, posterior, label, weight, bias, cfg, world)
    return total, terms, probs if label is not None else None


class _SetAttentionBlock(torch.autograd.Function):
    """One SAB of the set transformer as a single kernel per direction (csrc/sab.cu).  ``params`` in the order
    wq bq wk bk wv bv wo bo wf bf ln0_w ln0_b ln1_w ln1_b."""

    @staticmethod
    def forward(ctx, x, presence, eps0, eps1, *params):
        lib = _lib.load()
        x = x.contiguous()
        presence = presence.contiguous() if presence is not None else None
        params = tuple(p.contiguous() for p in params)
        B, N, _ = x.shape
        sp = _lib.SabParams(*[ptr(p) for p in params], eps0, eps1)
        y = torch.empty_like(x)
        check(_timed('scae_sab_fwd', lib.scae_sab_fwd, ptr(x), ptr(presence), ctypes.byref(sp), B, N, ptr(y),
                     _stream()), 'scae_sab_fwd')
        ctx.save_for_backward(x, presence, *params)
        ctx.eps = (eps0, eps1)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, presence, *params = ctx.saved_tensors
        B, N, _ = x.shape
        sp = _lib.SabParams(*[ptr(p) for p in params], *ctx.eps)
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        gp = torch.empty(5 * 256 + 9 * 16, device=x.device, dtype=torch.float32)
        ws_bytes = lib.scae_sab_bwd_workspace_bytes(B, N)
        ws = _workspace(ws_bytes, x.device)
        check(_timed('scae_sab_bwd', lib.scae_sab_bwd, ptr(x), ptr(presence), ctypes.byref(sp), ptr(gy), B, N, ptr(gx),
                     ptr(gp), ptr(ws), ws_bytes, _stream()), 'scae_sab_bwd')
        w = [gp[k * 256:(k + 1) * 256].view(16, 16) for k in range(5)]          # wq wk wv wo wf
        v = [gp[1280 + k * 16:1280 + (k + 1) * 16] for k in range(9)]           # bq bk bv bo bf g0 b0 g1 b1
        grads = (w[0], v[0], w[1], v[1], w[2], v[2], w[3], v[3], w[4], v[4], v[5], v[6], v[7], v[8])
        return (gx, None, None, None) + grads


def set_attention_block(x, presence, mab):
    """``MAB(x, x, presence)`` of a set_transformer.MAB (n_heads = 1, d = 16, layer_norm) through the fused kernels, or
    None when the block / input is not covered (the caller then runs the PyTorch ops)."""
    att = getattr(mab, 'mqkv', None)
    ok = (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[-1] == 16 and 0 < x.shape[1] <= 64
          and att is not None and att.n_heads == 1 and att.d_k == 16 and att.d_v == 16 and mab.layer_norm
          and (presence is None or (not presence.requires_grad and presence.dtype == torch.float32)))
    if not ok:
        return None
    params = (att.q_projector.weight, att.q_projector.bias, att.k_projector.weight, att.k_projector.bias,
              att.v_projector.weight, att.v_projector.bias, att.o_projector.weight, att.o_projector.bias,
              mab.fc.weight, mab.fc.bias, mab.ln0.weight, mab.ln0.bias, mab.ln1.weight, mab.ln1.bias)
    return _SetAttentionBlock.apply(x, presence, float(mab.ln0.eps), float(mab.ln1.eps), *params)


# =================================================================================================================
# hot path 1
# =================================================================================================================

def _tmpl_args(templates, templates_alpha, pose, presence, bg_image, bg_value, bg_mixing_logit, temperature_logit,
               scale, output_size, color=None):
    """``color`` (B,M,C): fused colourisation -- ``templates`` are then the batch-shared raw templates (1|-,M,C,h,w)."""
    M, C, h, w = templates.shape[-4:]
    B = pose.shape[0]
    if color is not None:
        if templates.numel() != M * C * h * w or tuple(color.shape) != (B, M, C):
            raise ValueError(f'fused colourisation needs raw templates (1,M,C,h,w) and colour (B,M,C); got '
                             f'{tuple(templates.shape)} and {tuple(color.shape)}')
    elif templates.dim() != 5 or templates.shape[0] != B:
        raise ValueError(f'templates have shape {tuple(templates.shape)}, expected ({B}, M, C, h, w)')
    H, W = output_size
    alpha_mode = templates_alpha is not None
    a = TmplArgs(ptr(templates), ptr(templates_alpha), ptr(pose), ptr(presence), ptr(bg_image), ptr(bg_value),
                 ptr(bg_mixing_logit) if alpha_mode else None, None if alpha_mode else ptr(temperature_logit),
                 ptr(scale), ptr(color), B, M, C, h, w, H, W,
                 _lib.TMPL_MODE_ALPHA if alpha_mode else _lib.TMPL_MODE_TEMPERATURE)
    return a


class TemplateMixtureLogProb(torch.autograd.Function):
    """(log_prob [B,C,H,W], ll [B]) of x under the template mixture; see include/scae_b200.h scae_tmpl_ll_fwd/bwd.

    Differentiable w.r.t. templates, pose, presence, bg_image and the decoder parameters; ``x`` is data (no grad).
    """

    @staticmethod
    def forward(ctx, templates, pose, presence, bg_image, x, templates_alpha, bg_value, bg_mixing_logit,
                temperature_logit, scale, output_size, color=None):
        lib = _lib.load()
        tensors = [_f32c(t) for t in (templates, pose, presence, bg_image, x, templates_alpha, bg_value,
                                      bg_mixing_logit, temperature_logit, scale, color)]
        templates, pose, presence, bg_image, x, templates_alpha, bg_value, bg_mixing_logit, temperature_logit, \
            scale, color = tensors
        M, C, h, w = templates.shape[-4:]
        B = pose.shape[0]
        H, W = output_size
        if tuple(x.shape) != (B, C, H, W):
            raise ValueError(f'log_prob target has shape {tuple(x.shape)}, expected {(B, C, H, W)}')
        if bg_image is None and bg_value is None:
            # same failure as the reference (part_decoder.py:192 reads self.bg_value)
            raise AttributeError("'TemplateBasedImageDecoder' object has no attribute 'bg_value'")
        args = _tmpl_args(templates, templates_alpha, pose, presence, bg_image, bg_value, bg_mixing_logit,
                          temperature_logit, scale, (H, W), color)
        log_prob = torch.empty(B, C, H, W, device=x.device, dtype=torch.float32)
        ll = torch.empty(B, device=x.device, dtype=torch.float32)
        cache = torch.empty(B, 2, C, H, W, device=x.device, dtype=torch.float32)
        check(_timed('scae_tmpl_ll_fwd', lib.scae_tmpl_ll_fwd, ctypes.byref(args), ptr(x), ptr(log_prob), ptr(ll),
                     ptr(cache), _stream()), 'scae_tmpl_ll_fwd')
        ctx.save_for_backward(*[t for t in tensors if t is not None], cache)
        ctx.present = [t is not None for t in tensors]
        ctx.output_size = (H, W)
        ctx.set_materialize_grads(False)
        return log_prob, ll

    @staticmethod
    def backward(ctx, g_log_prob, g_ll):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        cache = saved.pop()
        it = iter(saved)
        templates, pose, presence, bg_image, x, templates_alpha, bg_value, bg_mixing_logit, temperature_logit, scale, \
            color = [next(it) if p else None for p in ctx.present]
        M, C, h, w = templates.shape[-4:]
        B = pose.shape[0]
        H, W = ctx.output_size
        if g_log_prob is None and g_ll is None:
            return (None,) * 12
        if g_ll is not None:
            g = g_ll.reshape(B, 1, 1, 1).expand(B, C, H, W)
            g = g + g_log_prob if g_log_prob is not None else g
        else:
            g = g_log_prob
        g = _f32c(g)
        args = _tmpl_args(templates, templates_alpha, pose, presence, bg_image, bg_value, bg_mixing_logit,
                          temperature_logit, scale, (H, W), color)
        dev = templates.device
        g_templates = torch.empty_like(templates)            # raw-template gradient (batch-reduced) with colour
        g_color = torch.empty_like(color) if color is not None else None
        g_pose = torch.empty_like(pose)
        g_presence = torch.empty_like(presence) if presence is not None else None
        g_bg_image = torch.empty_like(bg_image) if bg_image is not None else None
        g_alpha = torch.empty_like(templates_alpha) if templates_alpha is not None else None
        g_scalars = torch.empty(4, device=dev, dtype=torch.float32)
        ws_bytes = lib.scae_tmpl_ll_bwd_workspace_bytes(ctypes.byref(args))
        ws = torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8)
        # kernel launches: the backward kernel + the partial-sum reductions (alpha gradient, scalar gradients)
        check(_timed('scae_tmpl_ll_bwd', lib.scae_tmpl_ll_bwd, ctypes.byref(args),
                     ptr(x), ptr(g), ptr(cache), ptr(g_templates), ptr(g_color), ptr(g_pose), ptr(g_presence),
                     ptr(g_bg_image), ptr(g_alpha), ptr(g_scalars), ptr(ws), ws_bytes, _stream()), 'scae_tmpl_ll_bwd')

        def scalar(i, p):
            return g_scalars[i:i + 1].reshape(p.shape) if p is not None else None
        return (g_templates, g_pose, g_presence, g_bg_image, None, g_alpha, scalar(0, bg_value),
                scalar(1, bg_mixing_logit), scalar(2, temperature_logit), scalar(3, scale), None, g_color)


def template_render(templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit,
                    temperature_logit, scale, output_size, want=('transformed_templates', 'mixing_logits'), color=None):
    """No-grad materialisation through scae_tmpl_render; returns a dict with the requested tensors."""
    lib = _lib.load()
    templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit, temperature_logit, scale, color = \
        [_f32c(t.detach()) if t is not None else None for t in (
            templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit, temperature_logit, scale,
            color)]
    if bg_image is None and bg_value is None:
        raise AttributeError("'TemplateBasedImageDecoder' object has no attribute 'bg_value'")
    M, C, h, w = templates.shape[-4:]
    B = pose.shape[0]
    H, W = output_size
    args = _tmpl_args(templates, templates_alpha, pose, presence, bg_image, bg_value, bg_mixing_logit,
                      temperature_logit, scale, (H, W), color)
    dev = templates.device
    shapes = dict(transformed_templates=(B, M + 1, C, H, W),
                  mixing_logits=(B, M + 1, 1 if templates_alpha is not None else C, H, W),
                  mode=(B, C, H, W), mean=(B, C, H, W),
                  mode_component=(B, 1 if templates_alpha is not None else C, H, W))
    out = {k: torch.empty(shapes[k], device=dev, dtype=torch.float32) for k in want}
    check(lib.scae_tmpl_render(ctypes.byref(args), ptr(out.get('transformed_templates')),
                               ptr(out.get('mixing_logits')), ptr(out.get('mode')), ptr(out.get('mean')),
                               ptr(out.get('mode_component')), _stream()), 'scae_tmpl_render')
    return out


class TemplateMixtureMode(torch.autograd.Function):
    """``pdf.mode()`` (distributions.py:50-77 without the straight-through estimator) with its backward: the render kernel
    also records which component every pixel took its value from, and scae_tmpl_mode_bwd scatters the upstream gradient
    into that component's warp (the same segmented-scan scatter as the likelihood backward).  Differentiable w.r.t.
    templates | (raw templates, colours), pose, bg_image, bg_value; the arg-max passes nothing to the mixing logits."""

    @staticmethod
    def forward(ctx, templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit, temperature_logit,
                scale, output_size, color=None):
        tensors = [_f32c(t) for t in (templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit,
                                      temperature_logit, scale, color)]
        r = template_render(*tensors[:9], output_size, ('mode', 'mode_component'), tensors[9])
        ctx.save_for_backward(*[t for t in tensors if t is not None], r['mode_component'])
        ctx.present = [t is not None for t in tensors]
        ctx.output_size = tuple(output_size)
        return r['mode']

    @staticmethod
    def backward(ctx, g_mode):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        comp = saved.pop()
        it = iter(saved)
        templates, pose, presence, bg_image, templates_alpha, bg_value, bg_mixing_logit, temperature_logit, scale, \
            color = [next(it) if p else None for p in ctx.present]
        M, C, h, w = templates.shape[-4:]
        B = pose.shape[0]
        H, W = ctx.output_size
        g = _f32c(g_mode)
        args = _tmpl_args(templates, templates_alpha, pose, presence, bg_image, bg_value, bg_mixing_logit,
                          temperature_logit, scale, (H, W), color)
        dev = templates.device
        # [B,2,C,H,W] in the layout of the likelihood cache: plane 0 = component index per channel, plane 1 unused
        cache = torch.zeros(B, 2, C, H, W, device=dev, dtype=torch.float32)
        cache[:, 0] = comp.expand(B, C, H, W)
        g_templates = torch.empty_like(templates)
        g_color = torch.empty_like(color) if color is not None else None
        g_pose = torch.empty_like(pose)
        g_bg_image = torch.empty_like(bg_image) if bg_image is not None else None
        g_scalars = torch.empty(4, device=dev, dtype=torch.float32)
        ws_bytes = lib.scae_tmpl_ll_bwd_workspace_bytes(ctypes.byref(args))
        ws = torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8)
        check(_timed('scae_tmpl_mode_bwd', lib.scae_tmpl_mode_bwd, ctypes.byref(args), ptr(g), ptr(cache),
                     ptr(g_templates), ptr(g_color), ptr(g_pose), ptr(g_bg_image), ptr(g_scalars), ptr(ws), ws_bytes,
                     _stream()), 'scae_tmpl_mode_bwd')
        g_bg_value = g_scalars[0:1].reshape(bg_value.shape) if bg_value is not None and bg_image is None else None
        return (g_templates, g_pose, None, g_bg_image, None, g_bg_value, None, None, None, None, g_color)


# =================================================================================================================
# hot path 2
# =================================================================================================================

# order of the tensors CapsuleVoteLikelihood returns
CAPS_RETURNS = ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote',
                'caps_presence', 'll_per_example', 'reg_per_example', 'vote_presence_binary', 'winner',
                'winner_presence', 'is_from_capsule', 'soft_winner', 'soft_winner_presence', 'posterior_mixing_prob',
                'mixing_log_prob', 'mixing_logit')
_CAPS_NON_DIFF = ('vote_presence_binary', 'is_from_capsule')


def _caps_args(all_param, cpr_static, biases, noise_caps, noise_vote, x, presence, dummy_vote, flags):
    B, O, A = all_param.shape
    V = x.shape[1]
    return CapsArgs(ptr(all_param), ptr(cpr_static), ptr(biases[0]), ptr(biases[1]), ptr(biases[2]), ptr(biases[3]),
                    ptr(noise_caps), ptr(noise_vote), ptr(x), ptr(presence), ptr(dummy_vote), B, O, V, flags)


def caps_fast_path_count():
    """hot-path-2 calls served by the persistent kernels (csrc/caps_ll3*.cu) so far; bench.py and tests guard against
    silent fall-backs to the older paths with it"""
    return int(_lib.load().scae_caps_persistent_path_count())


EXPLICIT_RETURNS = ('ll_per_example', 'vote_presence_binary', 'winner', 'winner_presence', 'soft_winner',
                    'soft_winner_presence', 'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit', 'is_from_capsule')


class CapsuleExplicitLikelihood(torch.autograd.Function):
    """The standalone ``CapsuleLikelihood(vote, scale, vote_presence, dummy_vote)(x, presence)`` of the reference
    (object_decoder.py:243-372) on explicit vote tensors, one kernel per direction (csrc/caps_explicit.cu).
    ``vote`` (B,O,V,6), ``dummy_vote`` (1,1,V,6) | (V,6), ``x`` (B,V,6); returns EXPLICIT_RETURNS."""

    @staticmethod
    def forward(ctx, vote, scale, vote_presence, dummy_vote, x, presence):
        lib = _lib.load()
        vote, scale, vote_presence, dummy_vote, x, presence = [_f32c(t) for t in (vote, scale, vote_presence, dummy_vote, x,
                                                                                  presence)]
        B, O, V, P = vote.shape
        if P != 6 or tuple(x.shape) != (B, V, 6) or dummy_vote.numel() != V * 6:
            raise ValueError(f'expected vote (B,O,V,6), x (B,V,6), dummy_vote (1,1,V,6); got {tuple(vote.shape)}, '
                             f'{tuple(x.shape)}, {tuple(dummy_vote.shape)}')
        dev = vote.device
        f = dict(device=dev, dtype=torch.float32)
        shapes = dict(log_prob_per_point=(B, V), ll_per_example=(B,), vote_presence_binary=(B, O, V), winner=(B, V, 6),
                      winner_presence=(B, V), soft_winner=(B, V, 6), soft_winner_presence=(B, V),
                      posterior_mixing_prob=(B, O, V), mixing_log_prob=(B, O + 1, V), mixing_logit=(B, O + 1, V))
        out = {k: torch.empty(s, **f) for k, s in shapes.items()}
        out['winner_idx'] = torch.empty(B, V, device=dev, dtype=torch.int64)
        out['is_from_capsule'] = torch.empty(B, V, device=dev, dtype=torch.int64)
        point_ll = torch.empty(B, V, **f)
        args = CapsExplicitArgs(ptr(vote), ptr(scale), ptr(vote_presence), ptr(dummy_vote), ptr(x), ptr(presence),
                                ptr(point_ll), B, O, V)
        outs = CapsOutputs(*[ptr(out.get(k)) for k in CAPS_OUTPUT_FIELDS])
        check(_timed('scae_caps_explicit_fwd', lib.scae_caps_explicit_fwd, ctypes.byref(args), ctypes.byref(outs),
                     _stream()), 'scae_caps_explicit_fwd')
        ctx.has_presence = presence is not None
        ctx.save_for_backward(vote, scale, vote_presence, dummy_vote, x, *([presence] if presence is not None else []),
                              out['posterior_mixing_prob'], out['log_prob_per_point'], out['winner_idx'])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(out['vote_presence_binary'], out['is_from_capsule'])
        return tuple(out[k] for k in EXPLICIT_RETURNS)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        winner_idx, lse, posterior = saved.pop(), saved.pop(), saved.pop()
        vote, scale, vote_presence, dummy_vote, x = saved[:5]
        presence = saved[5] if ctx.has_presence else None
        B, O, V, _ = vote.shape
        g = {k: _f32c(v) for k, v in zip(EXPLICIT_RETURNS, grads) if v is not None}
        up = CapsUpstream(*[ptr(g.get(k[2:])) for k in CAPS_UPSTREAM_FIELDS])
        sv = CapsSaved(ptr(posterior), ptr(lse), None, ptr(winner_idx))
        args = CapsExplicitArgs(ptr(vote), ptr(scale), ptr(vote_presence), ptr(dummy_vote), ptr(x), ptr(presence), None,
                                B, O, V)
        g_vote, g_scale, g_vp = torch.empty_like(vote), torch.empty_like(scale), torch.empty_like(vote_presence)
        g_dummy = torch.empty_like(dummy_vote) if ctx.needs_input_grad[3] else None
        g_x = torch.empty_like(x) if ctx.needs_input_grad[4] else None
        g_presence = torch.empty_like(presence) if presence is not None and ctx.needs_input_grad[5] else None
        ws_bytes = lib.scae_caps_explicit_bwd_workspace_bytes(ctypes.byref(args))
        ws = torch.empty(max(ws_bytes, 16), device=vote.device, dtype=torch.uint8)
        check(_timed('scae_caps_explicit_bwd', lib.scae_caps_explicit_bwd, ctypes.byref(args), ctypes.byref(sv),
                     ctypes.byref(up), ptr(g_vote), ptr(g_scale), ptr(g_vp), ptr(g_dummy), ptr(g_x), ptr(g_presence),
                     ptr(ws), ws_bytes, _stream()), 'scae_caps_explicit_bwd')
        return g_vote, g_scale, g_vp, g_dummy, g_x, g_presence


class CapsuleVoteLikelihood(torch.autograd.Function):
    """Everything CapsuleObjectDecoder.forward computes after the per-capsule MLPs, as one kernel per direction."""

    @staticmethod
    def forward(ctx, all_param, cpr_static, b0, b1, b2, b3, dummy_vote, x, presence, noise_caps, noise_vote, flags):
        lib = _lib.load()
        all_param, cpr_static, b0, b1, b2, b3, dummy_vote, x, presence, noise_caps, noise_vote = \
            [_f32c(t) for t in (all_param, cpr_static, b0, b1, b2, b3, dummy_vote, x, presence, noise_caps, noise_vote)]
        B, O, A = all_param.shape
        V = x.shape[1]
        if A != 8 * V + 7:
            raise ValueError(f'all_param has {A} columns, expected 8*V+7 = {8 * V + 7}')
        dev = all_param.device
        f = dict(device=dev, dtype=torch.float32)
        shapes = dict(vote=(B, O, V, 6), scale=(B, O, V), vote_presence=(B, O, V),
                      presence_logit_per_caps=(B, O, 1), presence_logit_per_vote=(B, O, V), caps_presence=(B, O),
                      log_prob_per_point=(B, V), ll_per_example=(B,), reg_per_example=(B,),
                      vote_presence_binary=(B, O, V), winner=(B, V, 6), winner_presence=(B, V),
                      soft_winner=(B, V, 6), soft_winner_presence=(B, V), posterior_mixing_prob=(B, O, V),
                      mixing_log_prob=(B, O + 1, V), mixing_logit=(B, O + 1, V))
        out = {k: torch.empty(s, **f) for k, s in shapes.items()}
        out['caps_presence_arg'] = torch.empty(B, O, device=dev, dtype=torch.int32)
        out['winner_idx'] = torch.empty(B, V, device=dev, dtype=torch.int64)
        out['is_from_capsule'] = torch.empty(B, V, device=dev, dtype=torch.int64)
        args = _caps_args(all_param, cpr_static, (b0, b1, b2, b3), noise_caps, noise_vote, x, presence, dummy_vote,
                          flags)
        outs = CapsOutputs(*[ptr(out[k]) for k in CAPS_OUTPUT_FIELDS])
        check(_timed('scae_caps_ll_fwd', lib.scae_caps_ll_fwd, ctypes.byref(args), ctypes.byref(outs), _stream()),
              'scae_caps_ll_fwd')
        inputs = (all_param, cpr_static, b0, b1, b2, b3, dummy_vote, x, presence, noise_caps, noise_vote)
        ctx.present = [t is not None for t in inputs]
        ctx.save_for_backward(*[t for t in inputs if t is not None], out['posterior_mixing_prob'],
                              out['log_prob_per_point'], out['caps_presence_arg'], out['winner_idx'])
        ctx.flags = flags
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*[out[k] for k in _CAPS_NON_DIFF])
        return tuple(out[k] for k in CAPS_RETURNS)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        winner_idx = saved.pop()
        caps_arg = saved.pop()
        lse = saved.pop()
        posterior = saved.pop()
        it = iter(saved)
        all_param, cpr_static, b0, b1, b2, b3, dummy_vote, x, presence, noise_caps, noise_vote = \
            [next(it) if p else None for p in ctx.present]
        B, O, A = all_param.shape
        V = x.shape[1]
        g = {k: _f32c(v) for k, v in zip(CAPS_RETURNS, grads) if k not in _CAPS_NON_DIFF}
        up = CapsUpstream(*[ptr(g.get(k[2:])) for k in CAPS_UPSTREAM_FIELDS])
        sv = CapsSaved(ptr(posterior), ptr(lse), ptr(caps_arg), ptr(winner_idx))
        args = _caps_args(all_param, cpr_static, (b0, b1, b2, b3), noise_caps, noise_vote, x, presence, dummy_vote,
                          ctx.flags)
        dev = all_param.device
        g_all = torch.empty_like(all_param)
        g_shared = torch.empty(O, A, device=dev, dtype=torch.float32)
        g_dummy = torch.empty(V, 6, device=dev, dtype=torch.float32) if ctx.needs_input_grad[6] else None
        g_x = torch.empty_like(x) if ctx.needs_input_grad[7] else None
        g_presence = torch.empty_like(presence) if presence is not None and ctx.needs_input_grad[8] else None
        ws_bytes = lib.scae_caps_ll_bwd_workspace_bytes(ctypes.byref(args))
        ws = torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8)
        # kernel launches: backward kernel + finalize + reduction of the shared-parameter partials (+ dummy-vote sum)
        check(_timed('scae_caps_ll_bwd', lib.scae_caps_ll_bwd,
                     ctypes.byref(args), ctypes.byref(sv), ctypes.byref(up), ptr(g_all), ptr(g_shared), ptr(g_dummy),
                     ptr(g_x), ptr(g_presence), ptr(ws), ws_bytes, _stream()), 'scae_caps_ll_bwd')
        g_static = g_shared[:, :6 * V].reshape(cpr_static.shape)
        g_b0 = g_shared[:, 6 * V:6 * V + 6].reshape(b0.shape)
        g_b1 = g_shared[:, 6 * V + 6].reshape(b1.shape)
        g_b2 = g_shared[:, 6 * V + 7:7 * V + 7].reshape(b2.shape)
        g_b3 = g_shared[:, 7 * V + 7:].reshape(b3.shape)
        return (g_all, g_static, g_b0, g_b1, g_b2, g_b3, g_dummy.reshape(dummy_vote.shape) if g_dummy is not None else None,
                g_x, g_presence, None, None, None)
