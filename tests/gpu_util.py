"""Helpers shared by the GPU parity tests: seeded synthetic inputs and kernel-level runners."""
import math

import torch

from oracle import capsule_likelihood as cl
from oracle import template_likelihood as tl
from oracle.pose import pose_to_affine

DEV = 'cuda'


def strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def make_template_inputs(B, M, C, h, w, H, W, *, alpha, presence=True, bg_image=False, learn_scale=False, seed=0,
                         dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=dtype)
    n = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    d = dict(templates=r(B, M, C, h, w), pose=pose_to_affine(0.5 * n(B, M, 6)), x=r(B, C, H, W))
    d['presence'] = r(B, M) if presence else None
    if presence and B * M > 3:
        d['presence'].view(-1)[3] = 0.0
    d['bg_image'] = r(B, C, H, W) if bg_image else None
    params = dict(bg_mixing_logit=0.3 * n(1))
    if not bg_image:
        params['bg_value'] = 0.3 * n(1)
    if alpha:
        params['templates_alpha'] = n(1, M, 1, h, w)
    else:
        params['temperature_logit'] = r(1)
        params.pop('bg_mixing_logit')
    if learn_scale:
        params['scale'] = r(1)
    d['params'] = params
    d['weight'] = n(B, C, H, W)
    return d


def template_oracle(d, dtype=torch.float64):
    """fp64 (or fp32) oracle forward + autograd backward on CPU.  Returns dict of outputs and gradients."""
    cast = lambda t: None if t is None else t.to(dtype).clone().requires_grad_(True)
    leaf = {k: cast(d[k]) for k in ('templates', 'pose', 'presence', 'bg_image')}
    params = {k: cast(v) for k, v in d['params'].items()}
    lp = tl.image_log_likelihood(leaf['templates'], leaf['pose'], d['x'].to(dtype), leaf['presence'],
                                 leaf['bg_image'], **params)
    (lp * d['weight'].to(dtype)).sum().backward()
    out = dict(log_prob=lp.detach())
    for k, t in list(leaf.items()) + list(params.items()):
        if t is not None:
            out['g_' + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    return out


def template_cuda(d):
    """The same through TemplateMixtureLogProb on the GPU in fp32."""
    from torch_scae_b200 import ops
    cast = lambda t: None if t is None else t.to(DEV, torch.float32).clone().requires_grad_(True)
    leaf = {k: cast(d[k]) for k in ('templates', 'pose', 'presence', 'bg_image')}
    params = {k: cast(v) for k, v in d['params'].items()}
    H, W = d['x'].shape[-2:]
    lp, ll = ops.TemplateMixtureLogProb.apply(
        leaf['templates'], leaf['pose'], leaf['presence'], leaf['bg_image'], d['x'].to(DEV, torch.float32),
        params.get('templates_alpha'), params.get('bg_value'), params.get('bg_mixing_logit'),
        params.get('temperature_logit'), params.get('scale'), (H, W))
    (lp * d['weight'].to(DEV, torch.float32)).sum().backward()
    out = dict(log_prob=lp.detach(), ll=ll.detach())
    for k, t in list(leaf.items()) + list(params.items()):
        if t is not None:
            out['g_' + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    return out


CAPS_UP = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence', 'vote_presence',
           'soft_winner', 'soft_winner_presence', 'winner', 'winner_presence', 'vote', 'scale',
           'presence_logit_per_caps', 'presence_logit_per_vote', 'mixing_logit', 'mixing_log_prob')


def make_capsule_inputs(B, O, V, *, presence=True, noise=True, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=dtype)
    n = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    A = 8 * V + 7
    d = dict(all_param=torch.relu(n(B, O, A)), cpr_static=0.1 * n(1, O, V, 6),
             biases=[0.1 * n(1, O, 1, 6), 0.1 * n(1, O, 1), 0.1 * n(1, O, V), 0.1 * n(1, O, V)],
             dummy_vote=0.1 * n(1, 1, V, 6), x=pose_to_affine(0.5 * n(B, V, 6)))
    d['presence'] = r(B, V) if presence else None
    if presence and B * V > 1:
        d['presence'].view(-1)[1] = 0.0
    d['noise_caps'] = (r(B, O, 1) - .5) * 4 if noise else None
    d['noise_vote'] = (r(B, O, V) - .5) * 4 if noise else None
    shapes = dict(ll_per_example=(B,), reg_per_example=(B,), posterior_mixing_prob=(B, O, V), caps_presence=(B, O),
                  vote_presence=(B, O, V), soft_winner=(B, V, 6), soft_winner_presence=(B, V), winner=(B, V, 6),
                  winner_presence=(B, V), vote=(B, O, V, 6), scale=(B, O, V), presence_logit_per_caps=(B, O, 1),
                  presence_logit_per_vote=(B, O, V), mixing_logit=(B, O + 1, V), mixing_log_prob=(B, O + 1, V))
    d['up'] = {k: n(*s) for k, s in shapes.items()}
    return d


def _caps_loss(res, up, which, dtype, dev):
    loss = 0
    for k in which:
        loss = loss + (res[k] * up[k].to(dev, dtype)).sum()
    return loss


def capsule_oracle(d, flags, which=CAPS_UP, dtype=torch.float64):
    cast = lambda t: None if t is None else t.to(dtype).clone().requires_grad_(True)
    leaf = dict(all_param=cast(d['all_param']), cpr_static=cast(d['cpr_static']), dummy_vote=cast(d['dummy_vote']),
                x=cast(d['x']), presence=cast(d['presence']))
    biases = [cast(b) for b in d['biases']]
    nz = lambda t: None if t is None else t.to(dtype)
    res = cl.object_decoder_post_mlp(leaf['all_param'], leaf['cpr_static'], biases, leaf['dummy_vote'], leaf['x'],
                                     leaf['presence'], nz(d['noise_caps']), nz(d['noise_vote']), **flags)
    B, O, A = d['all_param'].shape
    V = d['x'].shape[1]
    # per-example partial sums as the kernel exposes them
    dyn = cl.split_all_param(leaf['all_param'], V)[0]
    res['reg_per_example'] = (dyn ** 2).sum((1, 2, 3)) / 2 if flags['allow_deformations'] else \
        torch.zeros(B, dtype=dtype) + 0 * leaf['all_param'].sum()
    s = res['scale'].unsqueeze(-1)
    lp = (-((leaf['x'].unsqueeze(1) - res['vote']) ** 2) / (2 * s ** 2) - torch.log(s) - cl.HALF_LOG_2PI).sum(-1)
    lp = torch.cat([lp, torch.full((B, 1, V), cl.DUMMY_LOG, dtype=dtype)], 1)
    per_point = (res['mixing_logit'] + lp).logsumexp(1)
    if leaf['presence'] is not None:
        per_point = per_point * leaf['presence']
    res['ll_per_example'] = per_point.sum(1)
    _caps_loss(res, d['up'], which, dtype, 'cpu').backward()
    out = {k: v.detach() for k, v in res.items()}
    for k, t in leaf.items():
        if t is not None:
            out['g_' + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    for i, b in enumerate(biases):
        out[f'g_b{i}'] = b.grad if b.grad is not None else torch.zeros_like(b)
    return out


def capsule_cuda(d, flags, which=CAPS_UP, part_grads=True, extra_bits=0):
    """part_grads=False: x / presence are data (what SCAE training does: stop_grad_caps_target), which is what lets the
    backward take the fast path (csrc/caps_ll2.cu); extra_bits: additional SCAE_CAPS_* flag bits (e.g. RELU_GRAD)."""
    from torch_scae_b200 import _lib, ops
    cast = lambda t: None if t is None else t.to(DEV, torch.float32).clone().requires_grad_(True)
    data = lambda t: None if t is None else t.to(DEV, torch.float32).clone()
    leaf = dict(all_param=cast(d['all_param']), cpr_static=cast(d['cpr_static']), dummy_vote=cast(d['dummy_vote']),
                x=(cast if part_grads else data)(d['x']), presence=(cast if part_grads else data)(d['presence']))
    biases = [cast(b) for b in d['biases']]
    nz = lambda t: None if t is None else t.to(DEV, torch.float32)
    bits = (_lib.CAPS_SIMILARITY if flags['similarity'] else 0) \
        | (_lib.CAPS_LEARN_VOTE_SCALE if flags['learn_vote_scale'] else 0) \
        | (_lib.CAPS_ALLOW_DEFORM if flags['allow_deformations'] else 0) | extra_bits
    res = dict(zip(ops.CAPS_RETURNS, ops.CapsuleVoteLikelihood.apply(
        leaf['all_param'], leaf['cpr_static'], *biases, leaf['dummy_vote'], leaf['x'], leaf['presence'],
        nz(d['noise_caps']), nz(d['noise_vote']), bits)))
    _caps_loss(res, d['up'], which, torch.float32, DEV).backward()
    out = {k: v.detach() for k, v in res.items()}
    out['log_prob'] = out['ll_per_example'].mean()
    for k, t in leaf.items():
        if t is not None:
            out['g_' + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    for i, b in enumerate(biases):
        out[f'g_b{i}'] = b.grad if b.grad is not None else torch.zeros_like(b)
    return out
