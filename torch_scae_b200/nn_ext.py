"""PyTorch building blocks for the parts of SCAE that stay in PyTorch (dense matmuls / convolutions).

Constructor signatures and the resulting ``state_dict`` layouts follow the reference's nn_ext.py:19-140 and
nn_utils.py:23-66 so that checkpoints are interchangeable.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def MLP(sizes, activation=nn.ReLU, activate_final=True, bias=True):
    """Linear/activation chain as an nn.Sequential (linears at even indices, like nn_ext.py:19-31)."""
    assert len(sizes) >= 2, "There must be at least two sizes"
    layers = []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        layers += [nn.Linear(fan_in, fan_out, bias=bias), activation()]
    return nn.Sequential(*(layers if activate_final else layers[:-1]))


class _ConvStack(nn.Sequential):
    """nn.Sequential of Conv2d / activation modules (same children, same state_dict) whose forward runs every
    Conv2d + ReLU pair through ``ops.conv_bias_act``: bias add and ReLU in one pass after the cuDNN convolution, ReLU
    mask and bias gradient in one pass before its backward."""

    def forward(self, x):
        from . import ops
        layers = list(self)
        i = 0
        while i < len(layers):
            layer = layers[i]
            if isinstance(layer, nn.Conv2d):
                relu = i + 1 < len(layers) and type(layers[i + 1]) is nn.ReLU
                x = ops.conv_bias_act(x, layer, relu)
                i += 2 if relu else 1
            else:
                x = layer(x)
                i += 1
        return x


def Conv2dStack(in_channels, out_channels, kernel_sizes, strides, activation=nn.ReLU, activate_final=True):
    """Unpadded conv/activation chain (convs at even indices, like nn_ext.py:34-59)."""
    assert len(out_channels) == len(kernel_sizes) == len(strides)
    layers = []
    for c_out, k, s in zip(out_channels, kernel_sizes, strides):
        layers += [nn.Conv2d(in_channels, c_out, kernel_size=k, stride=s), activation()]
        in_channels = c_out
    return _ConvStack(*(layers if activate_final else layers[:-1]))


def multiple_attention_pooling_2d(feature_map, n_attention_map):
    """(B, n*(D+1), G, G) -> (B, n*D, 1, 1): every group's last channel is a spatial attention logit map that
    softmax-pools the group's other D channels (nn_ext.py:76-101)."""
    B, C, H, W = feature_map.shape
    assert n_attention_map > 0
    assert C > n_attention_map, "Attention maps cannot be more than feature maps"
    assert C % n_attention_map == 0, "Incompatible attention map count"
    from . import ops
    fused = ops.attention_pool(feature_map, n_attention_map)     # one warp per group on CUDA (csrc/support.cu)
    if fused is not None:
        return fused
    grouped = feature_map.view(B, n_attention_map, C // n_attention_map, H * W)
    pooled = (grouped[:, :, :-1] * F.softmax(grouped[:, :, -1:], -1)).sum(-1)
    return pooled.reshape(B, C - n_attention_map, 1, 1)


def relu1(x):
    return F.relu6(x * 6.) / 6.


def choose_activation(name):
    if name == 'sigmoid':
        return torch.sigmoid
    if name == 'relu1':
        return relu1
    fn = getattr(F, name, None)
    if fn is None:
        raise ValueError('Invalid activation function: "{}".'.format(name))
    return fn


def measure_shape(network, input_shape, input_dtype=torch.float32):
    device = next(iter(network.parameters())).device
    with torch.no_grad():
        return network(torch.rand(1, *input_shape, dtype=input_dtype, device=device)).shape[1:]
