"""The DEVICE code of csrc/attnpool_cl.cu (attention pooling on the channels-last output of the 1x1 attention
convolution run as a GEMM) executed on the CPU under the SIMT emulation of tests/emu/simt.h, against the reference's
formula nn_ext.multiple_attention_pooling_2d (nn_ext.py:76-101) in fp64 with autograd."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err

EMU = os.path.join(ROOT, 'tests', 'emu')
CSRC = os.path.join(ROOT, 'torch_scae_b200', 'csrc')


def build_emulated(build, source, harness, exe_name):
    """Pastes the device part of csrc/<source> (between `namespace scae {` and the `// ---- host side` marker) and the
    pieces of common.cuh it uses into include files and compiles tests/emu/<harness> against them."""
    src = open(os.path.join(CSRC, source)).read()
    body = src.split('namespace scae {', 1)[1].split('// ---- host side', 1)[0]
    open(build / (source.replace('.cu', '') + '_device.inc'), 'w').write(body)
    common = open(os.path.join(CSRC, 'common.cuh')).read()
    consts = re.findall(r'^constexpr float kLogSafe\w+ = [^;]+;', common, re.M)
    warp_sum = re.search(r'__device__ __forceinline__ float warp_sum\(float v\) \{.*?\n\}', common, re.S).group(0)
    assert len(consts) == 2
    open(build / 'common_device.inc', 'w').write('\n'.join(consts) + '\n' + warp_sum + '\n')
    exe = build / exe_name
    subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-I', str(build), '-I', EMU,
                    '-I', os.path.join(ROOT, 'include'), os.path.join(EMU, harness), '-o', str(exe)], check=True)
    return str(exe)


@pytest.fixture(scope='module')
def emu_binary(tmp_path_factory):
    return build_emulated(tmp_path_factory.mktemp('attnpool_cl_emu'), 'attnpool_cl.cu', 'attnpool_cl_harness.cpp',
                          'attnpool_cl_emu')


@pytest.mark.parametrize('B,n,D,S,grid', [(3, 40, 23, 25, 4), (2, 5, 7, 49, 1), (9, 3, 40, 9, 2), (1, 2, 1, 64, 1),
                                          (4, 6, 33, 33, 3)])
def test_emulated_channels_last_attention_pooling(emu_binary, tmp_path, B, n, D, S, grid):
    g = torch.Generator().manual_seed(B * n + D + S)
    G = D + 1
    y = torch.randn(B, S, n * G, generator=g)
    up = torch.randn(B * n, D, generator=g)
    with open(tmp_path / 'in.bin', 'wb') as f:
        f.write(struct.pack('5i', B, n, D, S, grid))
        f.write(y.numpy().tobytes())
        f.write(up.numpy().tobytes())
    subprocess.run([emu_binary, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True, timeout=300)
    raw = np.fromfile(tmp_path / 'out.bin', dtype=np.float32)
    out = torch.from_numpy(raw[:B * n * D].copy()).view(B * n, D)
    gy = torch.from_numpy(raw[B * n * D:].copy()).view(B, S, n * G)
    # reference: NCHW formulation of nn_ext.py:76-101 on the same values
    y64 = y.double().requires_grad_(True)
    grouped = y64.view(B, S, n, G).permute(0, 2, 3, 1)                      # (B, n, G, S)
    pooled = (grouped[:, :, :-1] * torch.softmax(grouped[:, :, -1:], -1)).sum(-1).reshape(B * n, D)
    (ref_gy,) = torch.autograd.grad((pooled * up.double()).sum(), [y64])
    assert rel_err(out, pooled) < 1e-5
    assert rel_err(gy, ref_gy) < 1e-5
