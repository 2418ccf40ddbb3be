"""CNN part-capsule encoder.  Stays in PyTorch/cuDNN (dense convolutions; BASELINE.json north_star).

API and state-dict layout of the reference's part_encoder.py:26-113.
"""
import os
from typing import Tuple

import torch
import torch.nn as nn

from . import cv_ops
from .attrdict import AttrDict
from .nn_ext import Conv2dStack, measure_shape, multiple_attention_pooling_2d


class CNNEncoder(nn.Module):
    def __init__(self, input_shape, out_channels, kernel_sizes, strides, activation=nn.ReLU, activate_final=True):
        super().__init__()
        self.network = Conv2dStack(input_shape[0], out_channels, kernel_sizes, strides, activation, activate_final)
        self.output_shape = measure_shape(self.network, input_shape=input_shape)

    def forward(self, image):
        return self.network(image)


class CapsuleImageEncoder(nn.Module):
    """image -> per-part pose (B,M,6), presence (B,M), special features (B,M,S)."""

    def __init__(self, input_shape: Tuple[int, int, int], encoder: CNNEncoder, n_caps: int, n_poses: int,
                 n_special_features: int = 0, noise_scale: float = 4., similarity_transform: bool = False):
        super().__init__()
        self.input_shape = input_shape
        self.encoder = encoder
        self.n_caps = n_caps
        self.n_poses = n_poses
        self.n_special_features = n_special_features
        self.noise_scale = noise_scale
        self.similarity_transform = similarity_transform
        self.caps_dim_splits = [n_poses, 1, n_special_features]
        self.n_total_caps_dims = sum(self.caps_dim_splits)
        self.img_embedding_bias = nn.Parameter(torch.zeros(tuple(encoder.output_shape), dtype=torch.float32))
        # one extra channel per capsule: its attention-pooling logit map
        self.att_conv = nn.Conv2d(encoder.output_shape[0], n_caps * (self.n_total_caps_dims + 1), kernel_size=1)
        self.output_shapes = AttrDict(pose=(n_caps, n_poses), presence=(n_caps,),
                                      feature=(n_caps, n_special_features))

    def forward(self, image, presence_noise=None):
        """``presence_noise`` (B,M), already scaled, replaces the internally drawn training noise
        (part_encoder.py:105-107); used by parity tests to inject the reference's noise."""
        B = image.shape[0]
        from . import ops
        feature_map = self.encoder(image) + self.img_embedding_bias.unsqueeze(0)
        # 1x1 attention convolution + attention pooling: on CUDA one GEMM over the B*S positions and a channels-last
        # pooling kernel (ops.attention_conv_pool); SCAE_B200_ATT_GEMM=0 keeps the cuDNN convolution (A/B timing)
        h = ops.attention_conv_pool(feature_map, self.att_conv, self.n_caps) \
            if os.environ.get('SCAE_B200_ATT_GEMM', '1') != '0' else None
        if h is None:
            h = multiple_attention_pooling_2d(ops.conv_bias_act(feature_map, self.att_conv, relu=False), self.n_caps)
        h = h.view(B, self.n_caps, self.n_total_caps_dims)
        pose, presence_logit, feature = torch.split(h, self.caps_dim_splits, -1)
        presence_logit = presence_logit.squeeze(-1)
        if presence_noise is not None:
            presence_logit = presence_logit + presence_noise
        elif self.training and self.noise_scale > 0.:
            presence_logit = presence_logit + (torch.rand_like(presence_logit) - .5) * self.noise_scale
        return AttrDict(pose=cv_ops.geometric_transform(pose, self.similarity_transform),
                        presence=torch.sigmoid(presence_logit),
                        feature=feature if self.n_special_features > 0 else None)
