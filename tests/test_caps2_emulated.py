"""Hot path 2, fast path (csrc/caps_ll2.cu: pair-parallel kernels staged by TMA bulk copies on mbarriers, persistent
double-buffered backward) executed on the CPU and compared with the fp64 oracle (oracle/capsule_likelihood.py, pinned
against the reference's golden vectors in tests/test_oracle_golden.py).

tests/emu/simt.h runs the CUDA threads, tests/emu/ptx_emu.h stands in for the inline PTX of csrc/ptx_sm100.cuh with
DEFERRED asynchronous copies (bytes land when the mbarrier is waited on, bulk stores read their source at
wait_group.read), so a missing wait / fence-ordering bug shows up as wrong numbers here, without a GPU.  The real
csrc headers are compiled for the host; only the kernels' launch sequence is restated (tests/emu/caps2_harness.cpp).
The GPU parity tests of the same kernels are tests/test_gpu_capsule.py."""
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from emu_io import read_arrays, write_arrays
from gpu_util import capsule_oracle, make_capsule_inputs

EMU = os.path.join(ROOT, 'tests', 'emu')
CSRC = os.path.join(ROOT, 'torch_scae_b200', 'csrc')
FAST_PATH_UPSTREAM = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence', 'vote_presence',
                      'vote', 'scale', 'presence_logit_per_caps', 'presence_logit_per_vote', 'mixing_logit')
FLAG_BITS = dict(similarity=1, learn_vote_scale=2, allow_deformations=4)


@pytest.fixture(scope='module')
def emu_binary(tmp_path_factory):
    build = tmp_path_factory.mktemp('caps2_emu')
    src = open(os.path.join(CSRC, 'caps_ll2.cu')).read()
    body = src.split('namespace scae {', 1)[1].split('// ---- host side', 1)[0]
    open(build / 'caps_ll2_device.inc', 'w').write(body)
    exe = build / 'caps2_emu'
    subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-D_GNU_SOURCE', '-I', os.path.join(EMU, 'stubs'),
                    '-I', str(build), '-I', EMU, '-I', CSRC, '-I', os.path.join(ROOT, 'include'),
                    os.path.join(EMU, 'caps2_harness.cpp'), '-o', str(exe)], check=True)
    return str(exe)


CASES = [
    # B, O, V, threads fwd, threads bwd, persistent CTAs, stages, flags, presence, noise, upstream set
    (5, 4, 5, 64, 64, 2, 2, dict(similarity=False, learn_vote_scale=True, allow_deformations=True), True, True,
     FAST_PATH_UPSTREAM),
    (4, 3, 6, 32, 96, 1, 1, dict(similarity=True, learn_vote_scale=False, allow_deformations=False), False, False,
     FAST_PATH_UPSTREAM[:4]),
    (3, 10, 7, 96, 64, 3, 2, dict(similarity=False, learn_vote_scale=True, allow_deformations=True), True, False,
     FAST_PATH_UPSTREAM[:4]),
    # the MNIST shape of the train step (O = 32, V = 40: 1280 pairs) with its launch configuration
    (3, 32, 40, 448, 640, 2, 2, dict(similarity=False, learn_vote_scale=True, allow_deformations=True), True, True,
     FAST_PATH_UPSTREAM[:4]),
]


@pytest.mark.parametrize('B,O,V,tf,tb,grid,stages,flags,presence,noise,which', CASES)
def test_emulated_capsule_fast_path_matches_the_oracle(emu_binary, tmp_path, B, O, V, tf, tb, grid, stages, flags,
                                                       presence, noise, which):
    d = make_capsule_inputs(B, O, V, presence=presence, noise=noise, seed=B + O + V, dtype=torch.float32)
    d64 = {k: (v.double() if torch.is_tensor(v) else v) for k, v in d.items()}
    d64['biases'] = [b.double() for b in d['biases']]
    d64['up'] = {k: v.double() for k, v in d['up'].items()}
    ref = capsule_oracle(d64, flags, which=which)
    bits = sum(FLAG_BITS[k] for k, on in flags.items() if on)
    arrays = dict(cfg=np.array([B, O, V, bits, tf, tb, grid, stages], dtype=np.int32), all_param=d['all_param'],
                  cpr_static=d['cpr_static'], b0=d['biases'][0], b1=d['biases'][1], b2=d['biases'][2],
                  b3=d['biases'][3], noise_caps=d['noise_caps'], noise_vote=d['noise_vote'], x=d['x'],
                  presence=d['presence'], dummy_vote=d['dummy_vote'])
    for k in which:
        arrays['g_' + k] = d['up'][k]
    write_arrays(tmp_path / 'in.bin', arrays)
    subprocess.run([emu_binary, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True, timeout=900)
    got = read_arrays(tmp_path / 'out.bin', dict(caps_presence_arg=np.int32, winner_idx=np.int64,
                                                 is_from_capsule=np.int64))
    A = 8 * V + 7
    # forward: every tensor of the reference's result dict
    for k in ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote', 'caps_presence',
              'll_per_example', 'reg_per_example', 'vote_presence_binary', 'winner', 'winner_presence', 'soft_winner',
              'soft_winner_presence', 'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit'):
        assert rel_err(got[k].view(ref[k].shape), ref[k]) < 1e-5, k
    assert torch.equal(got['is_from_capsule'].view(B, V), ref['is_from_capsule'])
    # backward
    assert rel_err(got['g_all_param'].view(B, O, A), ref['g_all_param']) < 1e-4
    gs = got['g_shared'].view(O, A)
    assert rel_err(gs[:, :6 * V].reshape(ref['g_cpr_static'].shape), ref['g_cpr_static']) < 1e-4
    for name, sl in (('g_b0', slice(6 * V, 6 * V + 6)), ('g_b1', slice(6 * V + 6, 6 * V + 7)),
                     ('g_b2', slice(6 * V + 7, 7 * V + 7)), ('g_b3', slice(7 * V + 7, A))):
        r = ref[name]
        if float(r.abs().max()) == 0.0:
            assert float(gs[:, sl].abs().max()) == 0.0, name
        else:
            assert rel_err(gs[:, sl].reshape(r.shape), r) < 1e-4, name
    if presence:
        # the kernel's g_presence is the log-likelihood term's share: g_ll * logsumexp per point
        want = d64['up']['ll_per_example'].view(B, 1) * ref['log_prob_per_point'] if 'log_prob_per_point' in ref else None
        if want is not None:
            assert rel_err(got['g_presence'].view(B, V), want) < 1e-5
