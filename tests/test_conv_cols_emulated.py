"""The DEVICE code of csrc/conv_cols.cu (im2col / col2im for the part encoder's 3x3 convolutions in GEMM form) executed
on the CPU under the SIMT emulation of tests/emu/simt.h, against F.unfold / F.fold; plus the algebra of the GEMM-form
convolution (forward, data gradient, weight gradient) against F.conv2d with autograd in fp64."""
import struct
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from test_attnpool_cl_emulated import build_emulated


@pytest.fixture(scope='module')
def emu_binary(tmp_path_factory):
    return build_emulated(tmp_path_factory.mktemp('conv_cols_emu'), 'conv_cols.cu', 'conv_cols_harness.cpp',
                          'conv_cols_emu')


@pytest.mark.parametrize('group', [16, 32])
@pytest.mark.parametrize('B,C,H,W,stride', [(2, 40, 19, 19, 2), (1, 32, 9, 9, 1), (2, 5, 7, 6, 1), (1, 33, 8, 11, 2),
                                            (1, 64, 3, 3, 1)])
def test_emulated_im2col_col2im(emu_binary, tmp_path, B, C, H, W, stride, group):
    g = torch.Generator().manual_seed(C + H + stride)
    Ho, Wo = (H - 3) // stride + 1, (W - 3) // stride + 1
    L = Ho * Wo
    x = torch.randn(B, C, H, W, generator=g)
    dcols = torch.randn(B * L, C * 9, generator=g)
    with open(tmp_path / 'in.bin', 'wb') as f:
        f.write(struct.pack('6i', B, C, H, W, stride, group))
        f.write(x.numpy().tobytes())
        f.write(dcols.numpy().tobytes())
    subprocess.run([emu_binary, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True, timeout=600)
    raw = np.fromfile(tmp_path / 'out.bin', dtype=np.float32)
    cols = torch.from_numpy(raw[:dcols.numel()].copy()).view(B * L, C * 9)
    dx = torch.from_numpy(raw[dcols.numel():].copy()).view(B, C, H, W)
    ref_cols = F.unfold(x, 3, stride=stride).transpose(1, 2).reshape(B * L, C * 9)
    assert torch.equal(cols, ref_cols)                                   # pure data movement: exact
    ref_dx = F.fold(dcols.double().view(B, L, C * 9).transpose(1, 2), (H, W), 3, stride=stride)
    assert rel_err(dx, ref_dx) < 1e-6


@pytest.mark.parametrize('stride', [1, 2])
def test_gemm_form_convolution_is_conv2d(stride):
    """formulations.conv3x3_gemm_reference -- y = cols @ W2d^T + b, dx = col2im(g2d @ W2d), dW = g2d^T @ cols, what the
    CUDA path uses -- equals F.conv2d and its autograd gradients (nn_ext.py:34-59 Conv2dStack layers)."""
    import formulations
    torch.manual_seed(stride)
    x = torch.randn(3, 6, 9, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(5, 6, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.randn(5, dtype=torch.float64, requires_grad=True)
    ref = torch.relu(F.conv2d(x, w, b, stride))
    up = torch.randn_like(ref)
    g_ref = torch.autograd.grad((ref * up).sum(), [x, w, b])
    y, (gx, gw, gb) = formulations.conv3x3_gemm_reference(x.detach(), w.detach(), b.detach(), stride, True, up)
    assert rel_err(y, ref) < 1e-12
    for a, r in zip((gx, gw, gb), g_ref):
        assert rel_err(a, r) < 1e-12
