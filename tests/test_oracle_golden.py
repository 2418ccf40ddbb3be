"""Pins oracle/ against golden vectors produced by the real reference (tests/golden/make_golden.py).

The oracle uses the same torch CPU ops as the reference in the same order, so values are expected to agree to
a few ulp; the thresholds below (1e-6 fwd, 1e-5 grads, max-norm relative) leave room for BLAS/threading
differences between machines.  The numpy closed form is checked at fp64-vs-fp32 precision (5e-6).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_err, sub
from oracle import capsule_likelihood as cl
from oracle import scae_model
from oracle import template_likelihood as tl

DECODER = ['alpha_c1', 'alpha_c3_nopres', 'temp_c3_bgimage', 'temp_c1']
CAPSULE = ['default', 'similarity_plain', 'wide']
SCAE = ['enc', 'soft', 'hard', 'sparse', 'color_temp']
CAPSULE_FLAGS = dict(
    default=dict(similarity=False, learn_vote_scale=True, allow_deformations=True),
    similarity_plain=dict(similarity=True, learn_vote_scale=False, allow_deformations=False),
    wide=dict(similarity=False, learn_vote_scale=True, allow_deformations=True),
)


def decoder_kwargs(g, requires_grad=False):
    params = {k: v.clone().requires_grad_(requires_grad) for k, v in sub(g, 'param.').items()}
    return params


@pytest.mark.parametrize('case', DECODER)
def test_decoder_forward_and_grads(case):
    g = load_golden('decoder_' + case)
    params = decoder_kwargs(g, True)
    leaf = {k: g[k].clone().requires_grad_(True) for k in ('templates', 'pose', 'presence', 'bg_image') if k in g}
    loc, sigma, logits = tl.decode(leaf['templates'], leaf['pose'], g['x'].shape[-2:], leaf.get('presence'),
                                   leaf.get('bg_image'), **params)
    assert rel_err(loc, g['transformed_templates']) < 1e-6
    assert rel_err(logits, g['mixing_logits']) < 1e-6
    lp = tl.mixture_log_prob(loc, sigma, logits, g['x'])
    assert rel_err(lp, g['log_prob']) < 1e-6
    (lp * g['weight']).sum().backward()
    for k, t in leaf.items():
        assert rel_err(t.grad, g['g_' + k]) < 1e-5, k
    for k, p in params.items():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert rel_err(got, g['g_param.' + k]) < 1e-5 or float(g['g_param.' + k].abs().max()) == 0.0, k
    with torch.no_grad():
        assert rel_err(tl.mixture_mode(loc, sigma, logits), g['mode']) < 1e-6
        assert rel_err(tl.mixture_mean(loc, logits), g['mean']) < 1e-6
        assert rel_err(torch.log_softmax(logits, 1), g['mixing_log_prob']) < 1e-6
        if 'mode_maximum' in g:
            assert rel_err(tl.mixture_mode(loc, sigma, logits, maximum=True), g['mode_maximum']) < 1e-6


@pytest.mark.parametrize('case', DECODER)
def test_decoder_closed_form(case):
    g = load_golden('decoder_' + case)
    params = {k: v.numpy() for k, v in sub(g, 'param.').items()}
    lp = tl.closed_form_log_prob(g['templates'].numpy(), g['pose'].numpy(), g['x'].numpy(),
                                 g['presence'].numpy() if 'presence' in g else None,
                                 g['bg_image'].numpy() if 'bg_image' in g else None, **params)
    assert rel_err(lp, g['log_prob']) < 5e-6


@pytest.mark.parametrize('case', CAPSULE)
def test_capsule_forward_and_grads(case):
    g = load_golden('capsule_' + case)
    P = sub(g, 'param.')
    pre = 'capsule_layer.'
    shared = dict(cpr_static=P[pre + 'cpr_static'], **{f'b{i}': P[f'{pre}caps_bias_list.{i}'] for i in range(4)},
                  dummy_vote=P['dummy_vote'])
    shared = {k: v.clone().requires_grad_(True) for k, v in shared.items()}
    all_param = g['all_param'].clone().requires_grad_(True)
    x = g['x'].clone().requires_grad_(True)
    presence = g['presence'].clone().requires_grad_(True) if 'presence' in g else None
    res = cl.object_decoder_post_mlp(all_param, shared['cpr_static'], [shared[f'b{i}'] for i in range(4)],
                                     shared['dummy_vote'], x, presence, g.get('noise_caps'), g.get('noise_vote'),
                                     **CAPSULE_FLAGS[case])
    out = sub(g, 'out.')
    for k, ref in out.items():
        if ref.dtype == torch.int64:
            assert torch.equal(res[k], ref), k
        else:
            assert rel_err(res[k], ref) < 2e-6, k
    loss = 1.7 * res['log_prob'] + 0.9 * res['cpr_dynamic_reg_loss']
    for k, w in sub(g, 'weight.').items():
        loss = loss + 0.3 * (res[k] * w).sum()
    loss.backward()
    assert rel_err(all_param.grad, g['g_all_param']) < 1e-5
    assert rel_err(x.grad, g['g_x']) < 1e-5
    if presence is not None:
        assert rel_err(presence.grad, g['g_presence']) < 1e-5
    names = {'cpr_static': pre + 'cpr_static', 'dummy_vote': 'dummy_vote',
             **{f'b{i}': f'{pre}caps_bias_list.{i}' for i in range(4)}}
    for k, t in shared.items():
        ref = g['g_param.' + names[k]]
        if t.grad is None:                      # e.g. the scale bias when learn_vote_scale=False
            assert float(ref.abs().max()) == 0.0, k
        else:
            assert rel_err(t.grad, ref) < 1e-5, k


@pytest.mark.parametrize('case', CAPSULE)
def test_capsule_mlps_reproduce_all_param(case):
    g = load_golden('capsule_' + case)
    sd = {'obj_decoder.' + k: v for k, v in sub(g, 'param.').items()}
    O = g['all_param'].shape[1]
    cfg = dict(ocae_decoder_capsule=dict(n_caps=O))
    got = scae_model.capsule_mlps(sd, cfg, g['obj_encoding'])
    assert rel_err(got, g['all_param']) < 1e-6


def _factory_cfg(name):
    with open(os.path.join(GOLDEN, 'factory.json')) as f:
        return json.load(f)[name]


@pytest.mark.parametrize('case', SCAE)
def test_scae_forward_loss_grads(case):
    from golden.cases import scae_case_params
    from torch_scae_b200 import factory
    g = load_golden('scae_' + case)
    cfg = factory.prepare_model_params(**scae_case_params(case))
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sub(g, 'param.').items()}
    noise = dict(part_presence=g['noise_part_presence'], caps=g['noise_caps'], vote=g['noise_vote'])
    res = scae_model.scae_forward(sd, cfg, g['image'], noise, training=True)
    loss, log = scae_model.scae_loss(res, cfg, g['image'], g['label'])
    assert rel_err(loss, g['loss']) < 1e-6
    for k, ref in sub(g, 'log.').items():
        assert rel_err(log[k], ref) < 2e-6, k
    skip = {'rec_log_prob', 'rec_mixing_logits'}
    for k, ref in sub(g, 'out.').items():
        if k in skip:
            continue
        if ref.dtype == torch.int64:
            assert torch.equal(res[k], ref), k
        else:
            assert rel_err(res[k], ref) < 2e-6, k
    assert rel_err(scae_model.accuracy(res, g['label']), g['accuracy']) == 0.0
    loss.backward()
    for k, ref in sub(g, 'g_param.').items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        if float(ref.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, k
        else:
            assert rel_err(got, ref) < 2e-5, k


def test_factory_defaults_match_reference():
    from torch_scae_b200 import factory
    with open(os.path.join(GOLDEN, 'factory.json')) as f:
        cases = json.load(f)
    for name, c in cases.items():
        got = json.loads(json.dumps(factory.prepare_model_params(**c['args']), default=list))
        assert got == c['prepared'], name
