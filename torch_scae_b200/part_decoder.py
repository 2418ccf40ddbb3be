"""Template generator and the fused template-based image decoder (hot path 1).

``TemplateBasedImageDecoder`` keeps the reference's constructor, call signature, parameter names and returned keys
(part_decoder.py:113-243) but returns *lazy* results: the likelihood is evaluated by
``res.pdf.log_prob(x)`` -> csrc/tmpl_ll.cu, and ``transformed_templates`` / ``mixing_logits`` are rendered only when
read.  ``TemplateGenerator`` (part_decoder.py:31-110) stays in PyTorch (tiny elementwise work).
"""
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import math_ops, skinny
from .attrdict import AttrDict, LazyAttrDict
from .distributions import TemplateMixture
from .nn_ext import MLP, choose_activation, relu1


class TemplateGenerator(nn.Module):
    """Learnt templates (1,M,C,h,w), optionally coloured per image by an MLP of the part features."""

    def __init__(self, n_templates, n_channels, template_size, template_nonlin='relu1', dim_feature=None,
                 colorize_templates=False, color_nonlin='relu1'):
        super().__init__()
        self.n_templates = n_templates
        self.template_size = template_size
        self.n_channels = n_channels
        self.template_nonlin = choose_activation(template_nonlin)
        self.dim_feature = dim_feature
        self.colorize_templates = colorize_templates
        self.color_nonlin = choose_activation(color_nonlin)
        # templates start mutually orthogonal: rows of the Q factor of a random matrix, rescaled to [0, 1]
        n_elems = n_channels * template_size[0] * template_size[1]
        n = max(n_templates, n_elems)
        q = np.linalg.qr(np.random.uniform(size=[n, n]))[0][:n_templates, :n_elems]
        q = q.reshape(1, n_templates, n_channels, *template_size).astype(np.float32)
        self.template_logits = nn.Parameter(torch.from_numpy((q - q.min()) / (q.max() - q.min())))
        if colorize_templates:
            self.templates_color_mlp = MLP(sizes=[dim_feature, 32, n_channels])

    def raw_templates(self):
        """(1,M,C,h,w): the learnt templates after their non-linearity (part_decoder.py:91)."""
        return self.template_nonlin(self.template_logits)

    def color(self, feature):
        """(B,M,C) per-image template colours (part_decoder.py:93-103), or None when templates are not coloured."""
        if not (self.colorize_templates and feature is not None):
            return None
        B, M = feature.shape[:2]
        color = feature.reshape(B * M, -1)
        for layer in self.templates_color_mlp:          # Linear/ReLU chain; tall-skinny -> split-K weight grads
            color = skinny.linear(color, layer) if isinstance(layer, nn.Linear) else layer(color)
        if self.color_nonlin == relu1:
            color = color + .99
        return self.color_nonlin(color).view(B, M, -1)

    def forward(self, feature=None, batch_size=None):
        raw_templates = self.raw_templates()
        color = self.color(feature)
        if color is not None:
            templates = raw_templates * color[:, :, :, None, None]
        else:
            templates = raw_templates.repeat(batch_size, 1, 1, 1, 1)
        return AttrDict(raw_templates=raw_templates, templates=templates)


class TemplateBasedImageDecoder(nn.Module):
    """Affine-warps M templates into the image plane and mixes them per pixel (plus a background component)."""

    def __init__(self, n_templates: int, template_size: Tuple[int, int], output_size: Tuple[int, int],
                 learn_output_scale=False, use_alpha_channel=False, background_value=True):
        super().__init__()
        self.n_templates = n_templates
        self.template_size = template_size
        self.output_size = output_size
        self.learn_output_scale = learn_output_scale
        self.use_alpha_channel = use_alpha_channel
        self.background_value = background_value
        if use_alpha_channel:
            self.templates_alpha = nn.Parameter(torch.zeros(1, n_templates, 1, *template_size))
        else:
            self.temperature_logit = nn.Parameter(torch.rand(1))
        if learn_output_scale:
            self.scale = nn.Parameter(torch.rand(1))
        self.bg_mixing_logit = nn.Parameter(torch.tensor([0.0]))
        if background_value:
            self.bg_value = nn.Parameter(torch.tensor([0.0]))

    def output_scale(self):
        """sigma of every mixture component (part_decoder.py:220-223)."""
        if self.learn_output_scale:
            return F.softplus(self.scale) + 1e-4
        return torch.ones(1, device=self.bg_mixing_logit.device)

    def forward(self, templates, pose, presence=None, bg_image=None, template_color=None):
        """templates (B,M,C,h,w), pose (B,M,6), presence (B,M)|None, bg_image (B,C,H,W)|None -> AttrDict with the
        reference's keys ``transformed_templates`` (B,M+1,C,H,W), ``mixing_logits``, ``pdf`` (all lazy).

        Extension: with ``template_color`` (B,M,C), ``templates`` are the batch-shared raw
        templates (1,M,C,h,w) and the coloured per-image templates ``raw * colour`` (TemplateGenerator.forward,
        part_decoder.py:90-105) exist only inside the kernels."""
        M = templates.shape[-4]
        B = pose.shape[0]
        if template_color is None and (templates.dim() != 5 or templates.shape[0] != B):
            raise ValueError(f'templates have shape {tuple(templates.shape)}, expected ({B}, M, C, h, w)')
        if pose.shape[1] != M or pose.shape[-1] != 6:
            raise ValueError(f'pose has shape {tuple(pose.shape)}, expected {(B, M, 6)}')
        pdf = TemplateMixture(self, templates, pose.reshape(B, M, 6), presence, bg_image, template_color)
        res = LazyAttrDict(pdf=pdf)
        res.set_lazy('transformed_templates', lambda: pdf.transformed_templates)
        res.set_lazy('mixing_logits', lambda: pdf.mixing_logits)
        return res

    def differentiable_materialize(self, templates, pose, presence=None, bg_image=None):
        """(transformed_templates, mixing_logits) through differentiable PyTorch CUDA ops; only used when a gradient
        is requested *through the materialised tensors* (never on the training hot path)."""
        B, M, C, h, w = templates.shape
        H, W = self.output_size
        grid = F.affine_grid(pose.reshape(B * M, 2, 3), [B * M, C, H, W], align_corners=False)
        loc = F.grid_sample(templates.reshape(B * M, C, h, w), grid, align_corners=False).view(B, M, C, H, W)
        bg = bg_image.unsqueeze(1) if bg_image is not None else torch.sigmoid(self.bg_value).expand(B, 1, C, H, W)
        loc = torch.cat([loc, bg], 1)
        if self.use_alpha_channel:
            alpha = self.templates_alpha.expand(B, M, 1, h, w).reshape(B * M, 1, h, w)
            logits = F.grid_sample(alpha, grid, align_corners=False).view(B, M, 1, H, W)
            logits = torch.cat([logits, F.softplus(self.bg_mixing_logit).expand(B, 1, 1, H, W)], 1)
        else:
            logits = loc / (F.softplus(self.temperature_logit + .5) + 1e-4)
        if presence is not None:
            full = torch.cat([presence, presence.new_ones(B, 1)], 1)
            logits = logits + math_ops.log_safe(full).view(B, M + 1, 1, 1, 1)
        return loc, logits
