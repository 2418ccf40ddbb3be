"""The hand-derived backward formulas (oracle/manual_backward.py -- the blueprint of the CUDA backward kernels) must
agree with autograd through the op-for-op oracle, in fp64, to ~1e-10."""
import pytest
import torch

from conftest import load_golden, rel_err, sub
from oracle import capsule_likelihood as cl
from oracle import manual_backward as mb
from oracle import template_likelihood as tl
from test_oracle_golden import CAPSULE, CAPSULE_FLAGS, DECODER

F64 = torch.float64
UP_KEYS = dict(g_posterior_mixing_prob='posterior_mixing_prob', g_caps_presence='caps_presence',
               g_vote_presence='vote_presence', g_soft_winner='soft_winner',
               g_soft_winner_presence='soft_winner_presence', g_winner='winner', g_winner_presence='winner_presence',
               g_vote='vote', g_scale='scale', g_presence_logit_per_caps='presence_logit_per_caps',
               g_presence_logit_per_vote='presence_logit_per_vote', g_mixing_logit='mixing_logit',
               g_mixing_log_prob='mixing_log_prob')


@pytest.mark.parametrize('case', CAPSULE)
def test_capsule_manual_backward(case):
    g = load_golden('capsule_' + case, F64)
    P = sub(g, 'param.')
    pre = 'capsule_layer.'
    leaves = dict(all_param=g['all_param'], cpr_static=P[pre + 'cpr_static'], dummy_vote=P['dummy_vote'], x=g['x'],
                  **{f'b{i}': P[f'{pre}caps_bias_list.{i}'] for i in range(4)})
    if 'presence' in g:
        leaves['presence'] = g['presence']
    leaves = {k: v.clone().requires_grad_(True) for k, v in leaves.items()}
    biases = [leaves[f'b{i}'] for i in range(4)]
    flags = CAPSULE_FLAGS[case]
    res = cl.object_decoder_post_mlp(leaves['all_param'], leaves['cpr_static'], biases, leaves['dummy_vote'],
                                     leaves['x'], leaves.get('presence'), g.get('noise_caps'), g.get('noise_vote'),
                                     **flags)
    B = g['x'].shape[0]
    torch.manual_seed(0)
    up = {k: torch.randn_like(res[o]) for k, o in UP_KEYS.items()}
    up['g_ll_per_example'] = torch.randn(B, dtype=F64)
    up['g_reg_per_example'] = torch.randn(B, dtype=F64)
    # per-example partial sums as the kernel exposes them
    lse = (res['mixing_logit'] + torch.cat([torch.zeros_like(res['posterior_mixing_prob']),
                                            torch.zeros_like(res['mixing_logit'][:, :1])], 1)).detach()  # placeholder
    V = g['x'].shape[1]
    dyn = cl.split_all_param(leaves['all_param'], V)[0]
    reg_pe = (dyn ** 2).sum((1, 2, 3)) / 2 if flags['allow_deformations'] else torch.zeros(B, dtype=F64)
    # ll_per_example: recompute through the oracle pieces
    r2 = cl.capsule_likelihood(res['vote'], res['scale'], res['vote_presence'], leaves['dummy_vote'], leaves['x'],
                               None)
    per_point = (r2['mixing_logit'] + torch.cat([
        (-((leaves['x'].unsqueeze(1) - res['vote']) ** 2) / (2 * res['scale'].unsqueeze(-1) ** 2)
         - torch.log(res['scale'].unsqueeze(-1)) - mb.HALF_LOG_2PI).sum(-1),
        torch.full((B, 1, V), cl.DUMMY_LOG, dtype=F64)], 1)).logsumexp(1)
    if 'presence' in leaves:
        per_point = per_point * leaves['presence']
    ll_pe = per_point.sum(1)
    assert rel_err(ll_pe.mean(), res['log_prob']) < 1e-12
    loss = (up['g_ll_per_example'] * ll_pe).sum() + (up['g_reg_per_example'] * reg_pe).sum()
    for k, o in UP_KEYS.items():
        loss = loss + (up[k].reshape(res[o].shape) * res[o]).sum()
    loss.backward()

    flat = dict(up)
    flat['g_presence_logit_per_caps'] = up['g_presence_logit_per_caps'].reshape(B, -1)
    got = mb.capsule_forward_backward(g['all_param'], P[pre + 'cpr_static'],
                                      [P[f'{pre}caps_bias_list.{i}'] for i in range(4)], P['dummy_vote'], g['x'],
                                      g.get('presence'), g.get('noise_caps'), g.get('noise_vote'), flat, **flags)
    assert rel_err(got['g_all_param'], leaves['all_param'].grad) < 1e-9
    assert rel_err(got['g_x'], leaves['x'].grad) < 1e-9
    assert rel_err(got['g_dummy_vote'], leaves['dummy_vote'].grad.reshape(V, 6)) < 1e-9
    if 'presence' in leaves:
        assert rel_err(got['g_presence'], leaves['presence'].grad) < 1e-9
    O = g['all_param'].shape[1]
    gs = got['g_shared']
    assert rel_err(gs[:, :6 * V].reshape(1, O, V, 6), leaves['cpr_static'].grad) < 1e-9
    assert rel_err(gs[:, 6 * V:6 * V + 6].reshape(1, O, 1, 6), leaves['b0'].grad) < 1e-9
    assert rel_err(gs[:, 6 * V + 6:6 * V + 7].reshape(1, O, 1), leaves['b1'].grad) < 1e-9
    assert rel_err(gs[:, 6 * V + 7:7 * V + 7].reshape(1, O, V), leaves['b2'].grad) < 1e-9
    if flags['learn_vote_scale']:
        assert rel_err(gs[:, 7 * V + 7:].reshape(1, O, V), leaves['b3'].grad) < 1e-9
    else:
        assert float(gs[:, 7 * V + 7:].abs().max()) == 0.0


@pytest.mark.parametrize('case', DECODER)
def test_template_manual_backward(case):
    g = load_golden('decoder_' + case, F64)
    params = {k: v.clone().requires_grad_(True) for k, v in sub(g, 'param.').items()}
    leaf = {k: g[k].clone().requires_grad_(True) for k in ('templates', 'pose', 'presence', 'bg_image') if k in g}
    lp = tl.image_log_likelihood(leaf['templates'], leaf['pose'], g['x'], leaf.get('presence'), leaf.get('bg_image'),
                                 **params)
    (lp * g['weight']).sum().backward()
    got = mb.template_forward_backward(g['templates'], g['pose'], g['x'], g['weight'], g.get('presence'),
                                       g.get('bg_image'), **{k: v.detach() for k, v in params.items()})
    assert rel_err(got['log_prob'], lp) < 1e-12
    assert rel_err(got['g_templates'], leaf['templates'].grad) < 1e-9
    assert rel_err(got['g_pose'], leaf['pose'].grad) < 1e-9
    if 'presence' in leaf:
        assert rel_err(got['g_presence'], leaf['presence'].grad) < 1e-9
    if 'bg_image' in leaf:
        assert rel_err(got['g_bg_image'], leaf['bg_image'].grad) < 1e-9
    if 'templates_alpha' in params:
        assert rel_err(got['g_alpha'], params['templates_alpha'].grad.reshape(got['g_alpha'].shape)) < 1e-9
    for i, name in enumerate(('bg_value', 'bg_mixing_logit', 'temperature_logit', 'scale')):
        if name in params and params[name].grad is not None:
            assert rel_err(got['g_scalars'][i], params[name].grad.reshape(())) < 1e-9, name


@pytest.mark.parametrize('prior_type,posterior_type', [('l2', 'entropy'), ('entropy', 'kl'), ('kl', 'l2')])
@pytest.mark.parametrize('with_label', [True, False])
def test_loss_head_manual_backward(prior_type, posterior_type, with_label):
    """The loss-head formulas (sparsity losses + classifier cross-entropies on softmax outputs) vs autograd through the
    product's PyTorch restatement of object_decoder.py:431-493 / stacked_capsule_auto_encoder.py:243-285."""
    import torch.nn.functional as F
    from torch_scae_b200.object_decoder import sparsity_loss
    torch.manual_seed(3)
    B, O, V, K = 7, 5, 4, 3
    cp = torch.rand(B, O, dtype=F64)
    cp[2, 1] = 0.0                                      # log_safe's floor branch (k p < 1e-16)
    post = torch.rand(B, O, V, dtype=F64)
    post[4, 3] = 0.0
    label = torch.randint(0, K, (B,)) if with_label else None
    weight, bias = torch.randn(K, O, dtype=F64), torch.randn(K, dtype=F64)
    ws = (2.0, 0.35, 0.7, 0.2)
    leaves = [t.clone().requires_grad_(True) for t in (cp, post, weight, bias)]
    a, b, w_, b_ = leaves
    pw, pb = sparsity_loss(prior_type, a, n_classes=K, within_example_constant=None)
    qw, qb = sparsity_loss(posterior_type, b.sum(-1) / V, n_classes=K)
    total = ws[0] * pw + ws[1] * pb + ws[2] * qw + ws[3] * qb
    terms = [pw, pb, qw, qb]
    if with_label:
        p1 = torch.softmax(F.linear(a.detach(), w_, b_), -1)
        p2 = torch.softmax(F.linear(b.sum(-1).detach(), w_, b_), -1)
        x1, x2 = F.cross_entropy(p1, label), F.cross_entropy(p2, label)
        total = total + x1 + x2
        terms += [x1, x2]
    grads = torch.autograd.grad(total, leaves, allow_unused=True)
    got = mb.loss_head_forward_backward(cp, post, label, weight, bias, K, prior_type, posterior_type, ws)
    assert rel_err(got['total'], total) < 1e-12
    for i, t in enumerate(terms):
        assert rel_err(got['terms'][i], t) < 1e-12, i
    assert rel_err(got['g_caps_presence'], grads[0]) < 1e-10
    assert rel_err(got['g_posterior'], grads[1]) < 1e-10
    if with_label:
        assert rel_err(got['prior_cls_prob'], p1) < 1e-12 and rel_err(got['posterior_cls_prob'], p2) < 1e-12
        assert rel_err(got['g_weight'], grads[2]) < 1e-10 and rel_err(got['g_bias'], grads[3]) < 1e-10


def test_attention_head_as_gemm_is_the_same_function():
    """The restructured capsule head of the part encoder (1x1 convolution as a GEMM over positions, pooling on the
    channels-last result, bias after the pooling: tests/formulations.py) equals the reference formulation
    conv2d -> multiple_attention_pooling_2d (part_encoder.py:95-101, nn_ext.py:76-101), values and gradients, in fp64.
    The bias gradient of the groups' logit channels is analytically zero (softmax is shift invariant); autograd through
    the reference formulation leaves rounding noise there."""
    import torch.nn.functional as F
    import formulations
    from torch_scae_b200 import nn_ext
    torch.manual_seed(1)
    B, Cin, H, W, n, G = 3, 6, 4, 5, 7, 5
    x = torch.randn(B, Cin, H, W, dtype=F64, requires_grad=True)
    w = torch.randn(n * G, Cin, 1, 1, dtype=F64, requires_grad=True)
    b = torch.randn(n * G, dtype=F64, requires_grad=True)
    up = torch.randn(B, n * (G - 1), 1, 1, dtype=F64)
    ref = nn_ext.multiple_attention_pooling_2d(F.conv2d(x, w, b), n)
    got = formulations.attention_conv_pool_reference(x, w, b, n)
    assert rel_err(got, ref) < 1e-12
    g_ref = torch.autograd.grad((ref * up).sum(), [x, w, b])
    g_got = torch.autograd.grad((got * up).sum(), [x, w, b])
    for a, r in zip(g_got, g_ref):
        assert rel_err(a, r) < 1e-12
    assert float(g_got[2].view(n, G)[:, -1].abs().max()) == 0.0
