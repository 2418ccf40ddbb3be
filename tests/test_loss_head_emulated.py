"""The DEVICE code of csrc/loss_head.cu, executed on the CPU under a SIMT emulation (tests/emu/simt.h: a thread per
CUDA thread, real barriers behind __syncthreads and the warp shuffles), against the analytic oracle
(oracle/manual_backward.py::loss_head_forward_backward, itself checked against autograd).  This is how the kernel text
-- indexing, lane masks, the partial-sum layout, the launch sequence -- is verified where no GPU is present; the GPU
parity tests (tests/test_gpu_plumbing.py) then only have to confirm it."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import manual_backward as mb

EMU = os.path.join(ROOT, 'tests', 'emu')
CSRC = os.path.join(ROOT, 'torch_scae_b200', 'csrc')
TYPES = {'l2': 0, 'entropy': 1, 'kl': 2}


@pytest.fixture(scope='module')
def emu_binary(tmp_path_factory):
    from test_attnpool_cl_emulated import build_emulated
    return build_emulated(tmp_path_factory.mktemp('loss_head_emu'), 'loss_head.cu', 'loss_head_harness.cpp',
                          'loss_head_emu')


def run_emulated(exe, tmp_path, cp, post, label, weight, bias, cfg, grid, g_total):
    B, O, V = post.shape
    K = weight.shape[0] if label is not None else 0
    sparsity, prior_type, posterior_type, ws, consts = cfg
    with open(tmp_path / 'in.bin', 'wb') as f:
        f.write(struct.pack('9i', B, O, V, K, int(label is not None), int(sparsity), TYPES[prior_type],
                            TYPES[posterior_type], grid))
        f.write(struct.pack('8f', *ws, *consts, g_total))
        f.write(cp.numpy().astype(np.float32).tobytes())
        f.write(post.numpy().astype(np.float32).tobytes())
        if label is not None:
            f.write(label.numpy().astype(np.int64).tobytes())
            f.write(weight.numpy().astype(np.float32).tobytes())
            f.write(bias.numpy().astype(np.float32).tobytes())
    subprocess.run([exe, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True, timeout=300)
    out = np.fromfile(tmp_path / 'out.bin', dtype=np.float32)
    sizes = [8, 2 * B * K, B * O, B * O * V, K * O + K if K else 0]
    parts, at = [], 0
    for n in sizes:
        parts.append(torch.from_numpy(out[at:at + n].copy()))
        at += n
    assert at == out.size
    return dict(terms=parts[0], probs=parts[1].view(2, B, K) if K else None, g_cp=parts[2].view(B, O),
                g_post=parts[3].view(B, O, V), g_weight=parts[4][:K * O].view(K, O) if K else None,
                g_bias=parts[4][K * O:] if K else None)


@pytest.mark.parametrize('B,O,V,K,grid', [(21, 10, 8, 10, 2), (9, 33, 6, 3, 1), (40, 64, 4, 16, 3), (5, 32, 40, 10, 4)])
@pytest.mark.parametrize('prior_type,posterior_type', [('l2', 'entropy'), ('entropy', 'kl'), ('kl', 'l2')])
@pytest.mark.parametrize('with_label', [True, False])
def test_emulated_loss_head_kernels_match_the_oracle(emu_binary, tmp_path, B, O, V, K, grid, prior_type,
                                                     posterior_type, with_label):
    g = torch.Generator().manual_seed(B * O + V)
    cp = torch.rand(B, O, generator=g)
    post = torch.rand(B, O, V, generator=g) / O
    if prior_type != 'l2':
        cp[1, 2] = 0.0                           # log_safe's floor branch
    label = torch.randint(0, K, (B,), generator=g) if with_label else None
    weight, bias = torch.randn(K, O, generator=g) * 0.3, torch.randn(K, generator=g) * 0.3
    ws = (2.0, 0.35, 0.7, 0.2)
    consts = (float(O) / K, float(O) / K, float(B) / K)
    g_total = 1.7
    got = run_emulated(emu_binary, tmp_path, cp, post, label, weight, bias, (True, prior_type, posterior_type, ws, consts),
                       grid, g_total)
    ref = mb.loss_head_forward_backward(cp.double(), post.double(), label, weight.double(), bias.double(), K,
                                        prior_type, posterior_type, ws)
    n_terms = 6 if with_label else 4
    for i in range(n_terms):
        assert abs(float(got['terms'][i]) - float(ref['terms'][i])) <= 2e-5 * abs(float(ref['terms'][i])) + 1e-6, i
    assert rel_err(got['terms'][6], ref['total']) < 2e-5
    assert rel_err(got['g_cp'], g_total * ref['g_caps_presence']) < 1e-4
    assert rel_err(got['g_post'], g_total * ref['g_posterior']) < 1e-4
    if with_label:
        assert rel_err(got['probs'][0], ref['prior_cls_prob']) < 1e-5
        assert rel_err(got['probs'][1], ref['posterior_cls_prob']) < 1e-5
        assert rel_err(got['g_weight'], g_total * ref['g_weight']) < 1e-4
        assert rel_err(got['g_bias'], g_total * ref['g_bias']) < 1e-4


def test_emulated_loss_head_classifier_only(emu_binary, tmp_path):
    """sparsity = 0 (both prior weights 0, stacked_capsule_auto_encoder.py:243): only the classifier terms."""
    g = torch.Generator().manual_seed(2)
    B, O, V, K = 11, 12, 8, 10
    cp, post = torch.rand(B, O, generator=g), torch.rand(B, O, V, generator=g) / O
    label = torch.randint(0, K, (B,), generator=g)
    weight, bias = torch.randn(K, O, generator=g) * 0.3, torch.randn(K, generator=g) * 0.3
    ws = (0.0, 0.0, 0.7, 0.2)
    got = run_emulated(emu_binary, tmp_path, cp, post, label, weight, bias,
                       (False, 'l2', 'entropy', ws, (1.2, 1.2, 1.1)), 2, 1.0)
    ref = mb.loss_head_forward_backward(cp.double(), post.double(), label, weight.double(), bias.double(), K, 'l2',
                                        'entropy', ws, sparsity=False)
    assert float(got['terms'][:4].abs().max()) == 0.0
    assert rel_err(got['terms'][4:6], ref['terms'][4:6]) < 1e-5 and rel_err(got['terms'][6], ref['total']) < 1e-5
    assert rel_err(got['g_weight'], ref['g_weight']) < 1e-4 and rel_err(got['g_bias'], ref['g_bias']) < 1e-4
