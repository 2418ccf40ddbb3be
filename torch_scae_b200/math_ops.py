"""Small PyTorch helpers used by the loss terms that stay outside the fused kernels.

Same names and semantics as the reference's math_ops.py:18-34 (the (B,O)-sized sparsity losses keep using them;
inside the kernels ``log_safe`` is restated in CUDA, see csrc/common.cuh).
"""
import torch

LOG_SAFE_EPS = 1e-16
LOG_SAFE_FLOOR = -1e8


def log_safe(tensor, eps=LOG_SAFE_EPS):
    tiny = tensor < eps
    logs = torch.log(torch.where(tiny, torch.ones_like(tensor), tensor))
    return torch.where(tiny, torch.full_like(tensor, LOG_SAFE_FLOOR), logs)


def cross_entropy_safe(true_probs, probs, dim=-1):
    return torch.mean(-torch.sum(true_probs * log_safe(probs), dim=dim))


def normalize(tensor, dim):
    return tensor / (torch.sum(tensor, dim, keepdim=True) + 1e-8)


def l2_loss(tensor):
    return torch.sum(tensor ** 2) / 2
