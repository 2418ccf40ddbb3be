"""Pose 6-vector -> affine parameters, PyTorch version for the part encoder (small, (B,M,6)-sized).

Semantics of the reference's ``cv_ops.geometric_transform`` (cv_ops.py:20-76).  The object decoder does not call
this: its vote composition is fused into the ``caps_ll`` CUDA kernels (csrc/caps_ll.cu), which restate the same
arithmetic.  Unlike the reference (cv_ops.py:45) nothing is modified in place, so it works under autograd.
"""
import math

import torch


def geometric_transform(pose_tensor, similarity=False, nonlinear=True, as_matrix=False):
    if nonlinear and not as_matrix and pose_tensor.is_cuda:
        from . import ops
        fused = ops.pose_transform(pose_tensor, similarity)      # ~60 tiny launches fwd + bwd -> one kernel each
        if fused is not None:
            return fused
    sx, sy, theta, shear, tx, ty = pose_tensor.unbind(-1)
    if nonlinear:
        sx, sy = torch.sigmoid(sx) + 1e-2, torch.sigmoid(sy) + 1e-2
        tx, ty, shear = torch.tanh(tx * 5.), torch.tanh(ty * 5.), torch.tanh(shear * 5.)
        theta = theta * (2. * math.pi)
    else:
        sx, sy = abs(sx) + 1e-2, abs(sy) + 1e-2
    c, s = torch.cos(theta), torch.sin(theta)
    if similarity:
        entries = (sx * c, -sx * s, tx, sx * s, sx * c, ty)
    else:
        entries = (sx * c + shear * sy * s, -sx * s + shear * sy * c, tx, sy * s, sy * c, ty)
    pose = torch.stack(entries, -1)
    if as_matrix:
        last = pose.new_zeros(*pose.shape[:-1], 3)
        last[..., 2] = 1
        pose = torch.cat([pose.view(*pose.shape[:-1], 2, 3), last.unsqueeze(-2)], -2)
    return pose
