"""Stock-PyTorch restatements of the algebraic RESTRUCTURINGS the CUDA path uses around the hot paths (any device and
dtype), so that the CPU tests can pin each one against the reference formulation in fp64 with autograd:

* the part encoder's capsule head as a GEMM over positions with the bias added after the pooling
  (torch_scae_b200/ops.py::attention_conv_pool vs conv2d -> multiple_attention_pooling_2d, part_encoder.py:95-101);
* the GEMM formulation of the 3x3 convolutions (ops._Conv3x3Gemm vs F.conv2d, nn_ext.py:34-59).
Test infrastructure; nothing under torch_scae_b200/ imports this module.
"""
import torch


def attention_conv_pool_reference(feature_map, weight, bias, n_caps):
    """The algebra of ``ops.attention_conv_pool`` in stock PyTorch ops (any device / dtype): 1x1 convolution as a GEMM over
    the positions, attention pooling on the channels-last result, bias added AFTER the pooling -- exact, because the
    softmax weights of a group sum to one (pooled channels) and a constant added to every position's logit does not
    change the softmax (the group's logit channel).  Used by the tests to pin the restructuring against the reference
    formulation conv -> multiple_attention_pooling_2d (part_encoder.py:95-101, nn_ext.py:76-101)."""
    B, Cin, H, W = feature_map.shape
    Ctot = weight.shape[0]
    G = Ctot // n_caps
    y = feature_map.permute(0, 2, 3, 1).reshape(B * H * W, Cin) @ weight.reshape(Ctot, Cin).t()
    grouped = y.view(B, H * W, n_caps, G)
    pooled = (grouped[..., :-1] * torch.softmax(grouped[..., -1:], 1)).sum(1)            # (B, n, D)
    return (pooled + bias.view(n_caps, G)[:, :-1]).reshape(B, n_caps * (G - 1), 1, 1)


def conv3x3_gemm_reference(x, weight, bias, stride, relu, g):
    """The GEMM formulation of ``relu?(conv2d(x, weight, bias, stride))`` and of its three gradients for an upstream
    gradient ``g``, in stock PyTorch ops (any device / dtype; F.unfold / F.fold stand in for the im2col / col2im
    kernels).  Returns (y, (gx, gw, gb)).  Pins ``ops._Conv3x3Gemm`` against F.conv2d in the CPU tests."""
    B, C, H, W = x.shape
    Co = weight.shape[0]
    Ho, Wo = (H - 3) // stride + 1, (W - 3) // stride + 1
    cols = torch.nn.functional.unfold(x, 3, stride=stride).transpose(1, 2).reshape(B * Ho * Wo, C * 9)
    w2d = weight.reshape(Co, C * 9)
    y2d = cols @ w2d.t() + bias
    if relu:
        y2d = torch.relu(y2d)
    y = y2d.view(B, Ho, Wo, Co).permute(0, 3, 1, 2)
    g2d = (g * (y > 0) if relu else g).permute(0, 2, 3, 1).reshape(B * Ho * Wo, Co)
    dcols = g2d @ w2d
    gx = torch.nn.functional.fold(dcols.view(B, Ho * Wo, C * 9).transpose(1, 2), (H, W), 3, stride=stride)
    return y, (gx, (g2d.t() @ cols).view_as(weight), g2d.sum(0))
