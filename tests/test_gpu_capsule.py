"""GPU parity, hot path 2: csrc/caps_ll.cu through the C ABI vs the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): rel <= 1e-5 on log-likelihoods / losses, rel <= 1e-4 on gradients, max-norm
relative (max|a-b| / max|b|) against the fp64 oracle evaluated on the same fp32-representable inputs.
"""
import pytest
import torch

from conftest import load_golden, rel_err, sub
from gpu_util import CAPS_UP, DEV, capsule_cuda, capsule_oracle, make_capsule_inputs, strict_fp32
from test_oracle_golden import CAPSULE, CAPSULE_FLAGS

pytestmark = pytest.mark.gpu
TOL_LL, TOL_OUT, TOL_GRAD = 1e-5, 1e-5, 1e-4
DEFAULT = dict(similarity=False, learn_vote_scale=True, allow_deformations=True)
FWD_KEYS = ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote', 'caps_presence',
            'vote_presence_binary', 'winner', 'winner_presence', 'soft_winner', 'soft_winner_presence',
            'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit', 'll_per_example', 'reg_per_example')
GRAD_KEYS = ('g_all_param', 'g_cpr_static', 'g_dummy_vote', 'g_x', 'g_presence', 'g_b0', 'g_b1', 'g_b2', 'g_b3')


def _f32(d):
    """round the fp64 synthetic inputs to fp32-representable values so both sides see identical numbers"""
    def r(t):
        if isinstance(t, torch.Tensor):
            return t.float().double()
        if isinstance(t, list):
            return [r(x) for x in t]
        if isinstance(t, dict):
            return {k: r(v) for k, v in t.items()}
        return t
    return {k: r(v) for k, v in d.items()}


def _compare(got, ref, keys, tol, ctx):
    for k in keys:
        if k not in ref or ref[k] is None:
            continue
        if ref[k].dtype == torch.int64:
            assert torch.equal(got[k].cpu(), ref[k]), (ctx, k)
            continue
        e = rel_err(got[k], ref[k].reshape(got[k].shape))
        assert e < tol, (ctx, k, e)


@pytest.mark.parametrize('B,O,V', [(8, 10, 40), (8, 32, 40), (5, 32, 64), (6, 32, 24), (3, 35, 6), (2, 3, 130),
                                   (1, 1, 1)])
def test_kernel_vs_fp64_oracle(B, O, V):
    d = _f32(make_capsule_inputs(B, O, V, seed=B * 1000 + O * 10 + V))
    ref = capsule_oracle(d, DEFAULT)
    got = capsule_cuda(d, DEFAULT)
    _compare(got, ref, FWD_KEYS, TOL_OUT, (B, O, V))
    assert rel_err(got['log_prob'], ref['log_prob']) < TOL_LL
    assert torch.equal(got['is_from_capsule'].cpu(), ref['is_from_capsule'])
    _compare(got, ref, GRAD_KEYS, TOL_GRAD, (B, O, V))


@pytest.mark.parametrize('flags', [dict(similarity=True, learn_vote_scale=False, allow_deformations=False),
                                   dict(similarity=True, learn_vote_scale=True, allow_deformations=True),
                                   dict(similarity=False, learn_vote_scale=False, allow_deformations=True)])
@pytest.mark.parametrize('presence,noise', [(True, True), (False, False)])
def test_kernel_flag_and_optional_input_combinations(flags, presence, noise):
    d = _f32(make_capsule_inputs(4, 7, 9, presence=presence, noise=noise, seed=7))
    ref = capsule_oracle(d, flags)
    got = capsule_cuda(d, flags)
    _compare(got, ref, FWD_KEYS, TOL_OUT, flags)
    _compare(got, ref, GRAD_KEYS, TOL_GRAD, flags)


def test_training_subset_of_upstream_gradients():
    """default SCAE training only feeds log_prob, posterior, caps_presence and the regulariser back (vote_type='enc')"""
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    d = _f32(make_capsule_inputs(6, 32, 40, seed=11))
    ref = capsule_oracle(d, DEFAULT, which)
    got = capsule_cuda(d, DEFAULT, which)
    _compare(got, ref, GRAD_KEYS, TOL_GRAD, 'train-subset')
    assert float(got['g_dummy_vote'].abs().max()) == 0.0      # measured in the survey: no grad in the default config


@pytest.mark.parametrize('case', CAPSULE)
def test_module_vs_reference_golden(case):
    """CapsuleObjectDecoder (batched MLPs + fused kernel) with the reference's weights, inputs and noise draws."""
    from golden.cases import CAPSULE_CASES
    from torch_scae_b200.object_decoder import CapsuleLayer, CapsuleObjectDecoder
    strict_fp32()
    c = CAPSULE_CASES[case]
    g = load_golden('capsule_' + case)
    layer = CapsuleLayer(c['O'], c['F'], c['V'], c['D'], hidden_sizes=c['hidden'],
                         learn_vote_scale=c['learn_vote_scale'], allow_deformations=c['allow_deformations'],
                         noise_type=c['noise_type'], noise_scale=c['noise_scale'], similarity_transform=c['similarity'])
    dec = CapsuleObjectDecoder(layer)
    dec.load_state_dict(sub(g, 'param.'), strict=True)
    dec.to(DEV)
    enc = g['obj_encoding'].to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    presence = g['presence'].to(DEV).requires_grad_(True) if 'presence' in g else None
    noise = (g['noise_caps'].to(DEV), g['noise_vote'].to(DEV)) if 'noise_caps' in g else None
    res = dec(enc, x, presence, noise=noise)
    out = sub(g, 'out.')
    assert set(out) == set(res.keys())
    for k, ref in out.items():
        if ref.dtype == torch.int64:
            assert torch.equal(res[k].cpu(), ref), k
        else:
            assert rel_err(res[k], ref) < (TOL_LL if k in ('log_prob', 'cpr_dynamic_reg_loss') else 2e-5), k
    loss = 1.7 * res.log_prob + 0.9 * res.cpr_dynamic_reg_loss
    for k, w in sub(g, 'weight.').items():
        loss = loss + 0.3 * (res[k] * w.to(DEV)).sum()
    loss.backward()
    assert rel_err(x.grad, g['g_x']) < TOL_GRAD
    if presence is not None:
        assert rel_err(presence.grad, g['g_presence']) < TOL_GRAD
    grads = {k: v for k, v in sub(g, 'g_param.').items()}
    assert rel_err(dec.dummy_vote.grad, grads['dummy_vote']) < TOL_GRAD
    assert rel_err(layer.cpr_static.grad, grads['capsule_layer.cpr_static']) < TOL_GRAD
    for i in range(4):
        ref = grads[f'capsule_layer.caps_bias_list.{i}']
        got = layer.caps_bias_list[i].grad
        if float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) == 0.0
        else:
            assert rel_err(got, ref) < TOL_GRAD, i


def test_full_size_properties():
    """BASELINE config sizes (B=1024, O=32, V=40): size-independent properties instead of an oracle run."""
    B, O, V = 1024, 32, 40
    d = make_capsule_inputs(B, O, V, seed=3, dtype=torch.float32)
    which = ('ll_per_example', 'posterior_mixing_prob', 'caps_presence', 'reg_per_example')
    a = capsule_cuda(d, DEFAULT, which)
    b = capsule_cuda(d, DEFAULT, which)
    for k in a:                                              # deterministic: bit-identical reruns
        assert torch.equal(a[k], b[k]), k
    post = a['posterior_mixing_prob']
    assert float(post.min()) >= 0 and float(post.sum(1).max()) <= 1 + 1e-5       # dummy component takes the rest
    mlp = a['mixing_log_prob']
    assert rel_err(mlp.exp().sum(1), torch.ones(B, V)) < 1e-5
    assert torch.equal(a['caps_presence'], a['vote_presence'].max(-1)[0])
    idx = a['posterior_mixing_prob'].argmax(1)               # hard winner = argmax of the posterior over objects
    win = torch.gather(a['vote'], 1, idx.view(B, 1, V, 1).expand(B, 1, V, 6)).squeeze(1)
    assert torch.equal(win, a['winner'])
    # batch permutation equivariance: per-example results do not depend on the batch position / CTA assignment
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    dp = dict(d)
    for k in ('all_param', 'x', 'presence', 'noise_caps', 'noise_vote'):
        dp[k] = d[k][perm]
    dp['up'] = {k: v[perm] for k, v in d['up'].items()}
    p = capsule_cuda(dp, DEFAULT, which)
    for k in ('ll_per_example', 'posterior_mixing_prob', 'soft_winner', 'g_all_param'):
        assert torch.equal(p[k], a[k][perm.to(a[k].device)]), k
    # linearity of the backward pass in the upstream gradient
    d2 = dict(d)
    d2['up'] = {k: 2 * v for k, v in d['up'].items()}
    c = capsule_cuda(d2, DEFAULT, which)
    assert rel_err(c['g_all_param'], 2 * a['g_all_param']) < 1e-6
    assert rel_err(c['g_cpr_static'], 2 * a['g_cpr_static']) < 1e-5


def test_full_size_values_vs_fp32_oracle():
    """Values at the full training batch (B = 1024, O = 32, V = 40) against the op-for-op oracle in fp32 on the host,
    chunked to bound its memory (about 5 MB of intermediates per image)."""
    B, O, V = 1024, 32, 40
    which = ('ll_per_example', 'posterior_mixing_prob', 'caps_presence', 'reg_per_example')
    d = make_capsule_inputs(B, O, V, seed=13, dtype=torch.float32)
    got = capsule_cuda(d, DEFAULT, which, part_grads=False)
    keys = ('ll_per_example', 'posterior_mixing_prob', 'caps_presence', 'vote', 'soft_winner', 'mixing_log_prob',
            'g_all_param')
    ref = {k: [] for k in keys}
    shared = {k: 0.0 for k in ('g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3')}
    for b0 in range(0, B, 128):
        part = dict(d)
        for k in ('all_param', 'x', 'presence', 'noise_caps', 'noise_vote'):
            part[k] = d[k][b0:b0 + 128]
        part['up'] = {k: v[b0:b0 + 128] for k, v in d['up'].items()}
        r = capsule_oracle(part, DEFAULT, which, dtype=torch.float32)
        for k in keys:
            ref[k].append(r[k])
        for k in shared:
            shared[k] = shared[k] + r[k]
    for k in keys:
        tol = TOL_GRAD if k.startswith('g_') else TOL_OUT
        assert rel_err(got[k], torch.cat(ref[k]).reshape(got[k].shape)) < tol, k
    for k, v in shared.items():
        assert rel_err(got[k], v.reshape(got[k].shape)) < TOL_GRAD, k
    from torch_scae_b200 import ops
    assert ops.caps_fast_path_count() > 0          # ... and it was the persistent kernels that produced them


# ---- fast path (csrc/caps_ll2.cu): pair-parallel kernels, taken when x / presence are data (SCAE training) -------------

FAST_UP = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence', 'vote_presence', 'scale',
           'vote', 'presence_logit_per_caps', 'presence_logit_per_vote', 'mixing_logit')
FAST_GRADS = ('g_all_param', 'g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3')


@pytest.mark.parametrize('B,O,V', [(8, 32, 40),      # MNIST config, 16-byte aligned image blocks
                                   (9, 10, 40),      # O*A*4 = 8 mod 16: blocks alternate between two alignments
                                   (7, 35, 6),       # O*A odd: all four alignments, several objects per warp
                                   (3, 5, 7),        # odd pair count: vote block leaves through the copy loop
                                   (330, 10, 40),    # more images than SMs: persistent backward, double-buffered prefetch
                                   (4, 32, 64),      # likelihood-stress config: single-stage backward (shared memory)
                                   (2, 3, 130), (1, 1, 1)])
def test_fast_path_vs_fp64_oracle(B, O, V):
    d = _f32(make_capsule_inputs(B, O, V, seed=B * 1000 + O * 10 + V))
    ref = capsule_oracle(d, DEFAULT, FAST_UP)
    got = capsule_cuda(d, DEFAULT, FAST_UP, part_grads=False)
    _compare(got, ref, FWD_KEYS, TOL_OUT, (B, O, V))
    _compare(got, ref, FAST_GRADS, TOL_GRAD, (B, O, V))
    assert float(got['g_dummy_vote'].abs().max()) == 0.0


@pytest.mark.parametrize('flags', [dict(similarity=True, learn_vote_scale=False, allow_deformations=False),
                                   dict(similarity=True, learn_vote_scale=True, allow_deformations=True),
                                   dict(similarity=False, learn_vote_scale=False, allow_deformations=True)])
@pytest.mark.parametrize('presence,noise', [(True, True), (False, False)])
def test_fast_path_flag_and_optional_input_combinations(flags, presence, noise):
    d = _f32(make_capsule_inputs(4, 7, 9, presence=presence, noise=noise, seed=7))
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    ref = capsule_oracle(d, flags, which)
    got = capsule_cuda(d, flags, which, part_grads=False)
    _compare(got, ref, FWD_KEYS, TOL_OUT, flags)
    _compare(got, ref, FAST_GRADS, TOL_GRAD, flags)


def test_fast_path_many_images_single_stage():
    """stress shape with several images per persistent CTA: the single-stage refill path of the backward kernel"""
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    d = _f32(make_capsule_inputs(300, 32, 64, seed=5))
    ref = capsule_oracle(d, DEFAULT, which)
    got = capsule_cuda(d, DEFAULT, which, part_grads=False)
    _compare(got, ref, ('ll_per_example', 'posterior_mixing_prob', 'caps_presence', 'vote'), TOL_OUT, 'stress')
    _compare(got, ref, FAST_GRADS, TOL_GRAD, 'stress')


def test_fast_path_fused_relu_mask():
    """SCAE_CAPS_RELU_GRAD: g_all_param is masked by (all_param > 0) (the MLP's final ReLU, nn_ext.py:19-31) while the
    gradients of cpr_static / biases still see the unmasked pre-activation gradient."""
    from torch_scae_b200 import _lib
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    d = _f32(make_capsule_inputs(6, 10, 40, seed=21))
    ref = capsule_oracle(d, DEFAULT, which)
    got = capsule_cuda(d, DEFAULT, which, part_grads=False, extra_bits=_lib.CAPS_RELU_GRAD)
    mask = (d['all_param'] > 0).double()
    assert rel_err(got['g_all_param'], ref['g_all_param'] * mask) < TOL_GRAD
    _compare(got, ref, ('g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3'), TOL_GRAD, 'relu')


def test_fast_and_general_paths_agree(monkeypatch):
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    d = make_capsule_inputs(200, 32, 40, seed=9, dtype=torch.float32)
    fast = capsule_cuda(d, DEFAULT, which, part_grads=False)
    again = capsule_cuda(d, DEFAULT, which, part_grads=False)
    for k in fast:                                           # deterministic: bit-identical reruns
        assert torch.equal(fast[k], again[k]), k
    monkeypatch.setenv('SCAE_CAPS_IMPL', 'v1')
    general = capsule_cuda(d, DEFAULT, which, part_grads=False)
    for k in FWD_KEYS:
        assert rel_err(fast[k], general[k]) < 2e-5, k
    for k in ('winner_idx', 'is_from_capsule') if 'winner_idx' in fast else ('is_from_capsule',):
        assert torch.equal(fast[k], general[k]), k
    for k in FAST_GRADS:
        assert rel_err(fast[k], general[k]) < 2e-4, k


def test_parameter_head_writes_kernel_layout():
    """PerCapsuleMLP.forward_contiguous (strided-output batched GEMM + in-place ReLU) equals the plain bmm chain, values
    and gradients, and yields the contiguous (B, O, A) block the fused kernel stages with bulk copies."""
    from torch_scae_b200.object_decoder import PerCapsuleMLP
    strict_fp32()
    torch.manual_seed(0)
    mlp = PerCapsuleMLP(10, [33, 128, 327], bias=False).to(DEV)
    x = torch.randn(64, 10, 33, device=DEV, requires_grad=True)
    w = torch.randn(64, 10, 327, device=DEV)
    ref = mlp(x)
    g_ref = torch.autograd.grad((ref * w).sum(), [x, mlp.w0, mlp.w1])
    out = mlp.forward_contiguous(x)
    assert out.is_contiguous() and out.shape == (64, 10, 327)
    assert rel_err(out, ref) < 1e-6
    g = torch.autograd.grad((out * w).sum(), [x, mlp.w0, mlp.w1])
    for a_, b_ in zip(g, g_ref):
        assert rel_err(a_, b_) < 1e-5
    # gradient already masked by the consumer (what the fused kernel does with SCAE_CAPS_RELU_GRAD)
    out2 = mlp.forward_contiguous(x, grad_is_masked=True)
    g2 = torch.autograd.grad(out2, [x, mlp.w0, mlp.w1], grad_outputs=w * (out2 > 0))
    for a_, b_ in zip(g2, g_ref):
        assert rel_err(a_, b_) < 1e-5
