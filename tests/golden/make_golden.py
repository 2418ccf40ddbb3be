"""Generate golden input/output vectors by running the UNMODIFIED reference from /root/reference.

Run in the build container only:  ``python tests/golden/make_golden.py``  (writes tests/golden/*.npz, *.json).
The fixtures are committed; tests never import the reference.  See _reference_loader.py for the two
accommodations (monty shim, out-of-place theta*2pi) needed to run the reference on torch 2.11.

Fixtures
  decoder_<case>.npz   reference TemplateBasedImageDecoder + GaussianMixture (hot path 1)
  capsule_<case>.npz   reference CapsuleObjectDecoder incl. per-capsule MLPs (hot path 2), all_param captured by hooks
  scae_<case>.npz      reference SCAE forward + loss + backward on a tiny model (state dict, noise, outputs, grads)
  factory.json         reference factory.prepare_model_params outputs
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

load_reference()
from cases import CAPSULE_CASES, DECODER_CASES, SCAE_CASES, scae_case_params, tiny_model_params  # noqa: E402
from torch_scae import cv_ops, factory  # noqa: E402
from torch_scae.object_decoder import CapsuleLayer, CapsuleObjectDecoder  # noqa: E402
from torch_scae.part_decoder import TemplateBasedImageDecoder  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrays)} arrays')


class _RecordRand:
    """Records every torch.rand_like result while active (the reference draws its noises with it)."""

    def __enter__(self):
        self.drawn = []
        self._orig = torch.rand_like

        def rec(*a, **k):
            out = self._orig(*a, **k)
            self.drawn.append(out.clone())
            return out
        torch.rand_like = rec
        return self

    def __exit__(self, *exc):
        torch.rand_like = self._orig


# ---------------------------------------------------------------------------------------------------------------
def make_decoder(name, c, seed):
    torch.manual_seed(seed)
    B, M, C = c['B'], c['M'], c['C']
    dec = TemplateBasedImageDecoder(M, c['tsize'], c['osize'], learn_output_scale=c['learn_scale'],
                                    use_alpha_channel=c['alpha'], background_value=c['bg_value'])
    with torch.no_grad():
        for p in dec.parameters():
            p.copy_(torch.randn_like(p) * 0.7)
    templates = torch.rand(B, M, C, *c['tsize'], requires_grad=True)
    raw = 0.5 * torch.randn(B, M, 6)
    raw[0, 0] = torch.tensor([3., 3., 0.1, 0., 0.15, -0.1])      # scale ~1: template partly outside the image
    raw[-1, -1] = torch.tensor([-3., -2., 0.4, 0.3, 0.02, 0.])    # tiny scale: strong magnification
    pose = cv_ops.geometric_transform(raw).detach().requires_grad_(True)
    presence = None
    if c['presence']:
        presence = torch.rand(B, M)
        presence[0, 1] = 0.0                                       # exercises log_safe's -1e8 branch
        presence.requires_grad_(True)
    bg_image = torch.rand(B, C, *c['osize'], requires_grad=True) if c['bg_image'] else None
    x = torch.rand(B, C, *c['osize'])
    res = dec(templates, pose, presence, bg_image)
    lp = res.pdf.log_prob(x)
    wgt = torch.randn_like(lp)
    (lp * wgt).sum().backward()
    out = dict(templates=_np(templates), pose=_np(pose), x=_np(x), weight=_np(wgt), log_prob=_np(lp),
               transformed_templates=_np(res.transformed_templates), mixing_logits=_np(res.mixing_logits),
               g_templates=_np(templates.grad), g_pose=_np(pose.grad))
    with torch.no_grad():
        res2 = dec(templates, pose, presence, bg_image)
        out.update(mode=_np(res2.pdf.mode()), mean=_np(res2.pdf.mean()),
                   mixing_log_prob=_np(res2.pdf.mixing_log_prob()))
        res3 = dec(templates, pose, presence, bg_image)
        try:
            out.update(mode_maximum=_np(res3.pdf.mode(maximum=True)))
        except RuntimeError:
            # reference quirk: the in-place `+=` at distributions.py:65 cannot broadcast [B,K,1,H,W] logits
            # against C>1 component densities, so alpha mode with C>1 raises; no golden for that combination.
            pass
    if presence is not None:
        out.update(presence=_np(presence), g_presence=_np(presence.grad))
    if bg_image is not None:
        out.update(bg_image=_np(bg_image), g_bg_image=_np(bg_image.grad))
    for k, p in dec.named_parameters():
        out['param.' + k] = _np(p)
        out['g_param.' + k] = _np(p.grad) if p.grad is not None else np.zeros_like(_np(p))
    _save('decoder_' + name, **out)


# ---------------------------------------------------------------------------------------------------------------
def make_capsule(name, c, seed):
    torch.manual_seed(seed)
    B, O, V = c['B'], c['O'], c['V']
    layer = CapsuleLayer(O, c['F'], V, c['D'], hidden_sizes=c['hidden'], learn_vote_scale=c['learn_vote_scale'],
                         allow_deformations=c['allow_deformations'], noise_type=c['noise_type'],
                         noise_scale=c['noise_scale'], similarity_transform=c['similarity'])
    dec = CapsuleObjectDecoder(layer)
    with torch.no_grad():
        for k, p in dec.named_parameters():
            if 'mlps' not in k:
                p.copy_(torch.randn_like(p) * 0.3)
    captured = []

    def grab(mod, inp, out):
        out.retain_grad()
        captured.append(out)

    for m in layer.caps_mlps:
        m.register_forward_hook(grab)
    enc = torch.randn(B, O, c['F'])
    x = cv_ops.geometric_transform(0.5 * torch.randn(B, V, 6)).detach().requires_grad_(True)
    presence = None
    if c['presence']:
        presence = torch.rand(B, V)
        presence[0, 0] = 0.0
        presence.requires_grad_(True)
    with _RecordRand() as rr:
        res = dec(enc, x, presence)
    ns = c['noise_scale']
    out = dict(obj_encoding=_np(enc), x=_np(x))
    if c['noise_type'] == 'uniform':
        out.update(noise_caps=_np((rr.drawn[0] - 0.5) * ns), noise_vote=_np((rr.drawn[1] - 0.5) * ns))
    w = {k: torch.randn_like(res[k]) for k in ('posterior_mixing_prob', 'caps_presence', 'soft_winner',
                                                'soft_winner_presence', 'winner', 'winner_presence', 'vote',
                                                'scale', 'vote_presence', 'mixing_logit', 'mixing_log_prob',
                                                'presence_logit_per_caps', 'presence_logit_per_vote')}
    # vote_presence can be exactly 0 only through underflow; weights on mixing_* kept small to stay well-scaled
    loss = 1.7 * res.log_prob + 0.9 * res.cpr_dynamic_reg_loss
    for k, t in w.items():
        loss = loss + 0.3 * (res[k] * t).sum()
        out['weight.' + k] = _np(t)
    loss.backward()
    for k, v in res.items():
        out['out.' + k] = _np(v)
    out['all_param'] = np.stack([_np(t) for t in captured], 1)
    out['g_all_param'] = np.stack([_np(t.grad) for t in captured], 1)
    out['g_x'] = _np(x.grad)
    if presence is not None:
        out.update(presence=_np(presence), g_presence=_np(presence.grad))
    for k, p in dec.state_dict().items():
        out['param.' + k] = _np(p)
    for k, p in dec.named_parameters():
        if 'mlps' not in k:
            out['g_param.' + k] = _np(p.grad) if p.grad is not None else np.zeros_like(_np(p))
    _save('capsule_' + name, **out)


# ---------------------------------------------------------------------------------------------------------------
def make_scae(name, scae_kwargs, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    params = scae_case_params(name)
    model = factory.make_scae(params)
    with torch.no_grad():                       # zero-initialised tensors would hide wiring mistakes
        for k, p in model.named_parameters():
            if k.endswith(('templates_alpha', 'cpr_static', 'dummy_vote', 'img_embedding_bias')) or \
                    'caps_bias_list' in k or 'bg_' in k:
                p.copy_(torch.randn_like(p) * 0.3)
    model.train()
    image = torch.rand(3, *params['image_shape'])
    label = torch.randint(0, 10, (3,))
    with _RecordRand() as rr:
        res = model(image)
    loss, log = model.loss(res, image, label)
    loss.backward()
    acc = model.calculate_accuracy(res, label)
    out = dict(image=_np(image), label=_np(label), loss=_np(loss), accuracy=_np(acc),
               noise_part_presence=_np((rr.drawn[0] - .5) * 4.), noise_caps=_np((rr.drawn[1] - .5) * 4.),
               noise_vote=_np((rr.drawn[2] - .5) * 4.))
    assert len(rr.drawn) == 3
    for k, v in log.items():
        out['log.' + k] = _np(v)
    for k in ('log_prob', 'caps_presence', 'posterior_mixing_prob', 'vote', 'scale', 'vote_presence', 'winner',
              'soft_winner', 'soft_winner_presence', 'winner_presence', 'prior_cls_prob', 'posterior_cls_prob',
              'templates', 'transformed_templates', 'is_from_capsule', 'mixing_logit', 'mixing_log_prob',
              'vote_presence_binary', 'part_presence', 'cpr_dynamic_reg_loss'):
        out['out.' + k] = _np(res[k])
    out['out.rec_log_prob'] = _np(res.rec.pdf.log_prob(image))
    out['out.rec_mixing_logits'] = _np(res.rec.mixing_logits)
    for k, p in model.state_dict().items():
        out['param.' + k] = _np(p)
    for k, p in model.named_parameters():
        out['g_param.' + k] = _np(p.grad) if p.grad is not None else np.zeros_like(_np(p))
    _save('scae_' + name, **out)
    return params


def make_explicit_likelihood(seed=400):
    """The reference's standalone CapsuleLikelihood on explicit vote tensors (object_decoder.py:243-372), as its own test
    builds it (tests/test_object_decoder.py:62-112), outputs and gradients w.r.t. every input."""
    from torch_scae.object_decoder import CapsuleLikelihood
    torch.manual_seed(seed)
    B, O, V, P = 3, 5, 7, 6
    leaf = lambda *s: torch.rand(*s).requires_grad_(True)
    vote, scale, vote_presence, dummy_vote = leaf(B, O, V, P), leaf(B, O, V), leaf(B, O, V), leaf(1, 1, V, P)
    with torch.no_grad():
        scale.add_(0.2)
        vote_presence[0, 1, 2] = 0.0               # log_safe floor
    x, presence = leaf(B, V, P), leaf(B, V)
    res = CapsuleLikelihood(vote=vote, scale=scale, vote_presence=vote_presence, dummy_vote=dummy_vote)(x, presence)
    out = dict(vote=_np(vote), scale=_np(scale), vote_presence=_np(vote_presence), dummy_vote=_np(dummy_vote), x=_np(x),
               presence=_np(presence))
    loss = 1.3 * res.log_prob
    for k in ('winner', 'winner_presence', 'soft_winner', 'soft_winner_presence', 'posterior_mixing_prob',
              'mixing_log_prob', 'mixing_logit'):
        w = torch.randn_like(res[k])
        out['weight.' + k] = _np(w)
        loss = loss + 0.4 * (res[k] * w).sum()
    for k, v in res.items():
        out['out.' + k] = _np(v)
    loss.backward()
    for k, t in dict(vote=vote, scale=scale, vote_presence=vote_presence, dummy_vote=dummy_vote, x=x,
                     presence=presence).items():
        out['g_' + k] = _np(t.grad)
    _save('capsule_likelihood_explicit', **out)


def make_hierarchical(seed=500):
    """The reference's CapsuleLayer with the hierarchical inputs parent_transform (B,O,1,3,3) and parent_presence (B,O,1)
    (object_decoder.py:183-188, :214-217): outputs and the gradients w.r.t. the feature, both parents and the
    non-MLP parameters."""
    torch.manual_seed(seed)
    c = CAPSULE_CASES['default']
    B, O, V = c['B'], c['O'], c['V']
    layer = CapsuleLayer(O, c['F'], V, c['D'], hidden_sizes=c['hidden'], learn_vote_scale=True, allow_deformations=True,
                         noise_type='uniform', noise_scale=4., similarity_transform=False)
    with torch.no_grad():
        for k, p in layer.named_parameters():
            if 'mlps' not in k:
                p.copy_(torch.randn_like(p) * 0.3)
    feature = torch.randn(B, O, c['F']).requires_grad_(True)
    parent_transform = cv_ops.geometric_transform(0.5 * torch.randn(B, O, 1, 6), as_matrix=True).detach().requires_grad_(True)
    parent_presence = torch.rand(B, O, 1).requires_grad_(True)
    with _RecordRand() as rr:
        res = layer(feature, parent_transform, parent_presence)
    out = dict(feature=_np(feature), parent_transform=_np(parent_transform), parent_presence=_np(parent_presence),
               noise_caps=_np((rr.drawn[0] - 0.5) * 4.), noise_vote=_np((rr.drawn[1] - 0.5) * 4.))
    loss = 0.9 * res.cpr_dynamic_reg_loss
    for k in ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote'):
        w = torch.randn_like(res[k])
        out['weight.' + k] = _np(w)
        loss = loss + 0.3 * (res[k] * w).sum()
    loss.backward()
    for k, v in res.items():
        out['out.' + k] = _np(v)
    out.update(g_feature=_np(feature.grad), g_parent_transform=_np(parent_transform.grad),
               g_parent_presence=_np(parent_presence.grad))
    for k, p in layer.state_dict().items():
        out['param.' + k] = _np(p)
    for k, p in layer.named_parameters():
        if 'mlps' not in k:
            out['g_param.' + k] = _np(p.grad) if p.grad is not None else np.zeros_like(_np(p))
    _save('capsule_hierarchical', **out)


def make_factory():
    cases = dict(
        mnist=dict(image_shape=(1, 40, 40), n_classes=10, n_part_caps=40, n_obj_caps=32),
        tiny=tiny_model_params(),
        color=dict(image_shape=(3, 32, 32), n_classes=10, n_part_caps=24, n_obj_caps=32,
                   pcae_decoder_params=dict(use_alpha_channel=False, learn_output_scale=True),
                   scae_params=dict(vote_type='soft')),
    )
    out = {k: dict(args=v, prepared=factory.prepare_model_params(**v)) for k, v in cases.items()}
    with open(os.path.join(HERE, 'factory.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True, default=list)
    print('factory.json written')


if __name__ == '__main__':
    for i, (n, c) in enumerate(DECODER_CASES.items()):
        make_decoder(n, c, 100 + i)
    for i, (n, c) in enumerate(CAPSULE_CASES.items()):
        make_capsule(n, c, 200 + i)
    for i, (n, c) in enumerate(SCAE_CASES.items()):
        make_scae(n, c, 300 + i)
    make_explicit_likelihood()
    make_hierarchical()
    make_factory()
