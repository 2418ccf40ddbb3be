"""Oracle for hot path 2: object->part vote composition + part-pose mixture likelihood.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the post-MLP half of ``CapsuleLayer.forward``
(object_decoder.py:160-236), ``CapsuleObjectDecoder.forward`` (:407-428) and ``CapsuleLikelihood.__call__``
(:257-372) as one function of plain tensors, differentiable through autograd, any dtype.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .pose import pose_to_affine
from .template_likelihood import log_safe

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
# The dummy component's logit and log-density are built as fp32 tensors whatever the input dtype
# (object_decoder.py:273-274, :281-282), so the constant is fp32(log 0.01).
DUMMY_LOG = float(np.float32(np.log(0.01)))


def split_all_param(all_param, n_votes):
    """object_decoder.py:91-99,:160-162: [B,O,8V+7] -> cpr_dynamic [B,O,V,6], cvr [B,O,1,6], caps logit [B,O,1],
    vote logit [B,O,V], raw scale [B,O,V]."""
    B, O, A = all_param.shape
    V = n_votes
    assert A == 8 * V + 7
    a, b, c, d, e = torch.split(all_param, [6 * V, 6, 1, V, V], -1)
    return a.reshape(B, O, V, 6), b.reshape(B, O, 1, 6), c, d, e


def capsule_votes(all_param, cpr_static, biases, noise_caps=None, noise_vote=None, *, similarity=False,
                  learn_vote_scale=True, allow_deformations=True):
    """Post-MLP half of CapsuleLayer.forward.  ``biases`` = caps_bias_list (4 tensors); ``noise_*`` are the
    *already scaled* additive noises ((rand-0.5)*noise_scale, object_decoder.py:201) or None.

    Returns dict(vote [B,O,V,6], scale, vote_presence, presence_logit_per_caps [B,O,1],
    presence_logit_per_vote [B,O,V], cpr_dynamic_reg_loss []).
    """
    B, O, _ = all_param.shape
    V = cpr_static.shape[2]
    cpr_dyn, cvr, caps_logit, vote_logit, raw_scale = split_all_param(all_param, V)
    if not allow_deformations:                                                     # :168-169
        cpr_dyn = torch.zeros_like(cpr_dyn)
    reg = torch.sum(cpr_dyn ** 2) / 2 / B                                          # :170, math_ops.py:33-34
    cpr = pose_to_affine(cpr_dyn + cpr_static, similarity, True, True)            # :171
    cvr = cvr + biases[0]                                                          # :176-179
    caps_logit = caps_logit + biases[1]
    vote_logit = vote_logit + biases[2]
    raw_scale = raw_scale + biases[3]
    cvr = pose_to_affine(cvr, similarity, True, True)                              # :185
    vote = torch.matmul(cvr.expand(B, O, V, 3, 3), cpr)                            # :189-191
    if noise_caps is not None:                                                     # :198-212
        caps_logit = caps_logit + noise_caps
    if noise_vote is not None:
        vote_logit = vote_logit + noise_vote
    vote_presence = torch.sigmoid(caps_logit) * torch.sigmoid(vote_logit)          # :217-219
    if learn_vote_scale:                                                           # :223-227
        scale = F.softplus(raw_scale + .5) + 1e-2
    else:
        scale = torch.ones_like(raw_scale)
    return dict(vote=vote[..., :-1, :].reshape(B, O, V, 6),                         # :413
                scale=scale, vote_presence=vote_presence, presence_logit_per_caps=caps_logit,
                presence_logit_per_vote=vote_logit, cpr_dynamic_reg_loss=reg)


def capsule_likelihood(vote, scale, vote_presence, dummy_vote, x, presence=None):
    """CapsuleLikelihood.__call__ (object_decoder.py:257-372).  x [B,V,6], presence [B,V] or None."""
    B, O, V, P = vote.shape
    s = scale.unsqueeze(-1)
    lp = (-((x.unsqueeze(1) - vote) ** 2) / (2 * s ** 2) - torch.log(s) - HALF_LOG_2PI).sum(-1)   # :263-269
    f32 = dict(dtype=torch.float32, device=x.device)
    dummy_lp = torch.zeros(B, 1, V, **f32) + np.log(0.01)                            # :273-274
    lp = torch.cat([lp, dummy_lp], 1)
    dummy_logit = torch.full((B, 1, V), fill_value=np.log(0.01), **f32)             # :281-282
    mixing_logit = torch.cat([log_safe(vote_presence), dummy_logit], 1)             # :284-285
    mixing_log_prob = mixing_logit - mixing_logit.logsumexp(1, keepdim=True)       # :286
    binary = (mixing_logit[:, :-1] > mixing_logit[:, -1:]).float()                  # :289
    post_logit = mixing_logit + lp                                                  # :292
    per_point = post_logit.logsumexp(1)                                             # :296
    if presence is not None:
        per_point = per_point * presence.float()                                    # :299
    log_prob = per_point.sum(1).mean()                                              # :302-306
    win = torch.argmax(post_logit[:, :-1], 1)                                       # :310-311
    gather_idx = win.unsqueeze(1)
    winner = torch.gather(vote, 1, gather_idx.unsqueeze(-1).expand(B, 1, V, P)).squeeze(1)     # :324
    winner_presence = torch.gather(vote_presence, 1, gather_idx).squeeze(1)         # :328-329
    is_from_capsule = win // V                                                      # :334 (sic)
    post = F.softmax(post_logit, 1)                                                 # :338
    votes_ext = torch.cat([vote, dummy_vote.expand(B, 1, V, P)], 1)                 # :341-344
    pres_ext = torch.cat([vote_presence, torch.zeros(B, 1, V, dtype=vote_presence.dtype, device=vote_presence.device)], 1)
    soft_winner = torch.sum(post.unsqueeze(-1) * votes_ext, 1)                      # :350
    soft_winner_presence = torch.sum(post * pres_ext, 1)                            # :354
    return dict(log_prob=log_prob, vote_presence_binary=binary, winner=winner, winner_presence=winner_presence,
                soft_winner=soft_winner, soft_winner_presence=soft_winner_presence,
                posterior_mixing_prob=post[:, :-1], mixing_log_prob=mixing_log_prob, mixing_logit=mixing_logit,
                is_from_capsule=is_from_capsule)


def object_decoder_post_mlp(all_param, cpr_static, biases, dummy_vote, x, presence=None, noise_caps=None,
                            noise_vote=None, **flags):
    """Everything CapsuleObjectDecoder.forward does after the per-capsule MLPs (object_decoder.py:410-428)."""
    res = capsule_votes(all_param, cpr_static, biases, noise_caps, noise_vote, **flags)
    res['caps_presence'] = res['vote_presence'].max(-1)[0]                          # :415
    res.update(capsule_likelihood(res['vote'], res['scale'], res['vote_presence'], dummy_vote, x, presence))
    return res
