"""Oracle: hand-derived backward passes of both hot paths, written out the way the CUDA kernels compute them.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference gets its gradients from autograd; the kernels cannot,
so the analytic formulas are stated here once in vectorised torch, checked against autograd of the op-for-op
restatements (tests/test_manual_backward.py) and then transcribed to CUDA (csrc/caps_ll.cu, csrc/tmpl_ll.cu).
Symbols follow DESIGN.md section 4.
"""
import math

import torch
import torch.nn.functional as F

from .capsule_likelihood import DUMMY_LOG, split_all_param

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
TWO_PI = 2.0 * math.pi


# =============================================================================================================
# hot path 2
# =============================================================================================================

def _transform_fwd(t, similarity):
    """t [...,6] -> (a [...,6], saved) with a = [a00,a01,a02,a10,a11,a12] (cv_ops.py:36-63)."""
    sx = torch.sigmoid(t[..., 0]) + 1e-2
    sy = torch.sigmoid(t[..., 1]) + 1e-2
    th = t[..., 2] * TWO_PI
    sh = torch.tanh(t[..., 3] * 5.)
    tx = torch.tanh(t[..., 4] * 5.)
    ty = torch.tanh(t[..., 5] * 5.)
    c, s = torch.cos(th), torch.sin(th)
    if similarity:
        a = torch.stack([sx * c, -sx * s, tx, sx * s, sx * c, ty], -1)
    else:
        a = torch.stack([sx * c + sh * sy * s, -sx * s + sh * sy * c, tx, sy * s, sy * c, ty], -1)
    return a, (sx, sy, sh, tx, ty, c, s)


def _transform_bwd(ga, saved, similarity):
    """ga [...,6] -> g_t [...,6]."""
    sx, sy, sh, tx, ty, c, s = saved
    g00, g01, g02, g10, g11, g12 = ga.unbind(-1)
    if similarity:
        g_sx = g00 * c - g01 * s + g10 * s + g11 * c
        g_sy = torch.zeros_like(g_sx)
        g_sh = torch.zeros_like(g_sx)
        g_c = (g00 + g11) * sx
        g_s = (g10 - g01) * sx
    else:
        g_sx = g00 * c - g01 * s
        g_sy = (g00 * s + g01 * c) * sh + g10 * s + g11 * c
        g_sh = (g00 * s + g01 * c) * sy
        g_c = g00 * sx + g01 * sh * sy + g11 * sy
        g_s = g00 * sh * sy - g01 * sx + g10 * sy
    g_th = g_s * c - g_c * s
    e0, e1 = sx - 1e-2, sy - 1e-2
    return torch.stack([g_sx * e0 * (1 - e0), g_sy * e1 * (1 - e1), g_th * TWO_PI, g_sh * 5. * (1 - sh * sh),
                        g02 * 5. * (1 - tx * tx), g12 * 5. * (1 - ty * ty)], -1)


def capsule_forward_backward(all_param, cpr_static, biases, dummy_vote, x, presence, noise_caps, noise_vote, up, *,
                             similarity=False, learn_vote_scale=True, allow_deformations=True):
    """Manual forward + backward of hot path 2.

    ``up``: dict of upstream gradients keyed like scae_caps_upstream in include/scae_b200.h (missing = zero).
    Returns dict(g_all_param [B,O,A], g_shared [O,A], g_dummy_vote [V,6], g_x [B,V,6], g_presence [B,V]).
    """
    B, O, A = all_param.shape
    V = cpr_static.shape[2]
    dt = all_param.dtype
    z = lambda *shape: torch.zeros(*shape, dtype=dt)
    G = lambda k, *shape: up[k] if up.get(k) is not None else z(*shape)

    dyn, cvr_raw, caps_raw, vote_raw, scale_raw = split_all_param(all_param, V)
    dyn_used = dyn if allow_deformations else torch.zeros_like(dyn)
    t = dyn_used + cpr_static                                         # [B,O,V,6]
    a, a_saved = _transform_fwd(t, similarity)
    r, r_saved = _transform_fwd((cvr_raw + biases[0]).squeeze(2), similarity)      # [B,O,6]
    r00, r01, r02, r10, r11, r12 = (r[..., i].unsqueeze(-1) for i in range(6))
    a00, a01, a02, a10, a11, a12 = a.unbind(-1)
    vote = torch.stack([r00 * a00 + r01 * a10, r00 * a01 + r01 * a11, r00 * a02 + r01 * a12 + r02,
                        r10 * a00 + r11 * a10, r10 * a01 + r11 * a11, r10 * a02 + r11 * a12 + r12], -1)
    lc = caps_raw + biases[1] + (noise_caps if noise_caps is not None else 0)       # [B,O,1]
    lv = vote_raw + biases[2] + (noise_vote if noise_vote is not None else 0)       # [B,O,V]
    pc, pv = torch.sigmoid(lc), torch.sigmoid(lv)
    vp = pc * pv
    u = scale_raw + biases[3]
    sc = F.softplus(u + .5) + 1e-2 if learn_vote_scale else torch.ones_like(u)
    diff = x.unsqueeze(1) - vote                                      # [B,O,V,6]
    q = (diff * diff).sum(-1)
    lp = -q / (2 * sc * sc) - 6 * torch.log(sc) - 6 * HALF_LOG_2PI
    tiny = vp < 1e-16
    ml = torch.where(tiny, torch.full_like(vp, -1e8), torch.log(torch.where(tiny, torch.ones_like(vp), vp)))
    dl = torch.full((B, 1, V), DUMMY_LOG, dtype=dt)
    ml_ext = torch.cat([ml, dl], 1)
    pl_ext = torch.cat([ml + lp, dl + dl], 1)
    post_ext = torch.softmax(pl_ext, 1)
    post = post_ext[:, :O]
    win = pl_ext[:, :O].argmax(1)                                     # [B,V]
    cp_arg = vp.argmax(-1)                                            # [B,O]
    pres = presence if presence is not None else torch.ones(B, V, dtype=dt)

    # ---- upstream -> gradient w.r.t. the posterior logits pl -------------------------------------------------
    g_sw, g_swp = G('g_soft_winner', B, V, 6), G('g_soft_winner_presence', B, V)
    h = G('g_posterior_mixing_prob', B, O, V) + (g_sw.unsqueeze(1) * vote).sum(-1) + g_swp.unsqueeze(1) * vp
    h_dummy = (g_sw * dummy_vote.reshape(1, V, 6)).sum(-1)            # [B,V]
    S = (post * h).sum(1) + post_ext[:, O] * h_dummy
    g_ll = G('g_ll_per_example', B).view(B, 1, 1)
    g_pl = post * (h - S.unsqueeze(1)) + g_ll * pres.unsqueeze(1) * post

    # ---- direct terms ------------------------------------------------------------------------------------------
    is_win = F.one_hot(win, O).permute(0, 2, 1).to(dt)                # [B,O,V]
    g_vote = G('g_vote', B, O, V, 6) + g_sw.unsqueeze(1) * post.unsqueeze(-1) \
        + is_win.unsqueeze(-1) * G('g_winner', B, V, 6).unsqueeze(1)
    is_cp = F.one_hot(cp_arg, V).to(dt)                               # [B,O,V]
    g_vp = G('g_vote_presence', B, O, V) + g_swp.unsqueeze(1) * post + is_win * G('g_winner_presence', B, V).unsqueeze(1) \
        + is_cp * G('g_caps_presence', B, O).unsqueeze(-1)
    g_mlp = G('g_mixing_log_prob', B, O + 1, V)
    g_ml = G('g_mixing_logit', B, O + 1, V)[:, :O] + g_mlp[:, :O] \
        - torch.softmax(ml_ext, 1)[:, :O] * g_mlp.sum(1, keepdim=True) + g_pl
    g_vp = g_vp + torch.where(tiny, torch.zeros_like(vp), g_ml / torch.where(tiny, torch.ones_like(vp), vp))
    # ---- Gaussian ---------------------------------------------------------------------------------------------
    inv2 = 1 / (sc * sc)
    g_vote = g_vote + (g_pl * inv2).unsqueeze(-1) * diff
    g_x = -((g_pl * inv2).unsqueeze(-1) * diff).sum(1)
    g_sc = g_pl * (q * inv2 / sc - 6 / sc) + G('g_scale', B, O, V)
    g_u = g_sc * torch.sigmoid(u + .5) if learn_vote_scale else torch.zeros_like(u)
    g_presence = G('g_ll_per_example', B).view(B, 1) * torch.logsumexp(pl_ext, 1)
    # ---- presence logits -----------------------------------------------------------------------------------------
    g_lv = g_vp * pc * pv * (1 - pv) + G('g_presence_logit_per_vote', B, O, V)
    g_lc = (g_vp * pv).sum(-1, keepdim=True) * pc * (1 - pc) + G('g_presence_logit_per_caps', B, O).unsqueeze(-1)
    # ---- vote = R . A --------------------------------------------------------------------------------------------
    gv = g_vote.unbind(-1)
    g_a = torch.stack([r00 * gv[0] + r10 * gv[3], r00 * gv[1] + r10 * gv[4], r00 * gv[2] + r10 * gv[5],
                       r01 * gv[0] + r11 * gv[3], r01 * gv[1] + r11 * gv[4], r01 * gv[2] + r11 * gv[5]], -1)
    g_r = torch.stack([(gv[0] * a00 + gv[1] * a01 + gv[2] * a02).sum(-1),
                       (gv[0] * a10 + gv[1] * a11 + gv[2] * a12).sum(-1), gv[2].sum(-1),
                       (gv[3] * a00 + gv[4] * a01 + gv[5] * a02).sum(-1),
                       (gv[3] * a10 + gv[4] * a11 + gv[5] * a12).sum(-1), gv[5].sum(-1)], -1)    # [B,O,6]
    g_t = _transform_bwd(g_a, a_saved, similarity)                    # [B,O,V,6]
    g_cvr = _transform_bwd(g_r, r_saved, similarity)                  # [B,O,6]
    # ---- assemble ------------------------------------------------------------------------------------------------
    g_pre = torch.cat([g_t.reshape(B, O, 6 * V), g_cvr, g_lc, g_lv, g_u], -1)       # w.r.t. pre-activation sums
    g_reg = G('g_reg_per_example', B).view(B, 1, 1, 1)
    g_dyn = (g_t + g_reg * dyn) if allow_deformations else torch.zeros_like(dyn)
    g_all = torch.cat([g_dyn.reshape(B, O, 6 * V), g_cvr, g_lc, g_lv, g_u], -1)
    g_dummy = (g_sw * post_ext[:, O].unsqueeze(-1)).sum(0)            # [V,6]
    return dict(g_all_param=g_all, g_shared=g_pre.sum(0), g_dummy_vote=g_dummy, g_x=g_x, g_presence=g_presence)


# =============================================================================================================
# hot path 1
# =============================================================================================================

def _base_coords(n, dtype):
    return (2.0 * torch.arange(n, dtype=dtype) + 1.0) / n - 1.0


def _taps(tex, ix, iy):
    """tex [..., h, w] broadcast against ix, iy [..., H, W]: returns the 4 zero-padded taps (nw, ne, sw, se), the
    fractional offsets and (x0, y0)."""
    h, w = tex.shape[-2:]
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx, fy = ix - x0, iy - y0
    x0, y0 = x0.long(), y0.long()

    lead = torch.broadcast_shapes(tex.shape[:-2], ix.shape[:-2])
    HW = ix.shape[-2:]

    def tap(dx, dy):
        xx, yy = x0 + dx, y0 + dy
        ok = ((xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)).expand(*lead, *HW)
        flat = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).expand(*lead, *HW)
        vals = torch.gather(tex.reshape(*tex.shape[:-2], h * w).expand(*lead, h * w), -1,
                            flat.reshape(*lead, -1)).reshape(*lead, *HW)
        return torch.where(ok, vals, torch.zeros_like(vals)), ok, flat
    return [tap(0, 0), tap(1, 0), tap(0, 1), tap(1, 1)], fx, fy


def template_forward_backward(templates, pose, x, grad_out, presence=None, bg_image=None, *, templates_alpha=None,
                              temperature_logit=None, scale=None, bg_mixing_logit=None, bg_value=None):
    """Manual forward + backward of hot path 1 (per-pixel formulas of DESIGN.md section 4).

    Returns dict(log_prob, g_templates, g_pose, g_presence, g_bg_image, g_alpha, g_scalars[4]).
    """
    B, M, C, h, w = templates.shape
    H, W = x.shape[-2:]
    dt = templates.dtype
    alpha_mode = templates_alpha is not None
    Xs, Ys = _base_coords(W, dt).view(1, 1, 1, W), _base_coords(H, dt).view(1, 1, H, 1)
    p = pose.view(B, M, 6, 1, 1)
    gx = p[:, :, 0] * Xs + p[:, :, 1] * Ys + p[:, :, 2]               # [B,M,H,W]
    gy = p[:, :, 3] * Xs + p[:, :, 4] * Ys + p[:, :, 5]
    ix = ((gx + 1) * w - 1) / 2
    iy = ((gy + 1) * h - 1) / 2
    sigma = (F.softplus(scale) + 1e-4).reshape(()) if scale is not None else torch.ones((), dtype=dt)
    i2s = 1 / (2 * sigma * sigma)

    taps_t, fx, fy = _taps(templates, ix.unsqueeze(2), iy.unsqueeze(2))          # values [B,M,C,H,W]
    fx, fy = fx.squeeze(2), fy.squeeze(2)
    wts = [(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy]           # [B,M,H,W]
    loc = sum(tv * wt.unsqueeze(2) for (tv, _, _), wt in zip(taps_t, wts))       # [B,M,C,H,W]
    if presence is not None:
        tiny = presence < 1e-16
        lpres = torch.where(tiny, torch.full_like(presence, -1e8),
                            torch.log(torch.where(tiny, torch.ones_like(presence), presence)))
    else:
        lpres = torch.zeros(B, M, dtype=dt)
    if alpha_mode:
        taps_a, _, _ = _taps(templates_alpha.reshape(1, M, h, w), ix, iy)
        a = sum(tv * wt for (tv, _, _), wt in zip(taps_a, wts))                   # [B,M,H,W]
        logit = (a + lpres.view(B, M, 1, 1)).unsqueeze(2).expand(B, M, C, H, W)
        bg_logit = F.softplus(bg_mixing_logit).reshape(()).expand(B, 1, C, H, W)
    else:
        tau = (F.softplus(temperature_logit + .5) + 1e-4).reshape(())
        logit = loc / tau + lpres.view(B, M, 1, 1, 1)
    bg_loc = bg_image.unsqueeze(1) if bg_image is not None else torch.sigmoid(bg_value).reshape(()).expand(B, 1, C, H, W)
    if not alpha_mode:
        bg_logit = bg_loc / tau
    loc_e = torch.cat([loc, bg_loc], 1)
    logit_e = torch.cat([logit, bg_logit], 1)
    d = x.unsqueeze(1) - loc_e
    n = -d * d * i2s + logit_e
    N = torch.logsumexp(n, 1)
    D = torch.logsumexp(logit_e, 1)
    log_prob = N - D - torch.log(sigma) - HALF_LOG_2PI

    # ---- backward ------------------------------------------------------------------------------------------------
    Gc = grad_out.unsqueeze(1)                                          # [B,1,C,H,W]
    pN = torch.exp(n - N.unsqueeze(1))
    pD = torch.exp(logit_e - D.unsqueeze(1))
    g_loc = Gc * pN * d * (2 * i2s)
    g_logit = Gc * (pN - pD)
    g_sigma = (Gc * (pN * d * d / sigma ** 3)).sum() - grad_out.sum() / sigma
    g_scalars = torch.zeros(4, dtype=dt)
    if not alpha_mode:
        g_loc = g_loc + g_logit / tau
        g_tau = -(g_logit * loc_e).sum() / (tau * tau)
        g_scalars[2] = g_tau * torch.sigmoid(temperature_logit.reshape(()) + .5)
    else:
        # alpha-mode logits are shared by the C channels; the bg logit enters each channel's mixture, so its
        # gradient is the plain sum over channels and pixels
        g_scalars[1] = g_logit[:, M].sum() * torch.sigmoid(bg_mixing_logit.reshape(()))
    if scale is not None:
        g_scalars[3] = g_sigma * torch.sigmoid(scale.reshape(()))
    g_bg_image = None
    if bg_image is not None:
        g_bg_image = g_loc[:, M]
    else:
        sb = torch.sigmoid(bg_value.reshape(()))
        g_scalars[0] = g_loc[:, M].sum() * sb * (1 - sb)
    g_loc_m = g_loc[:, :M]                                              # [B,M,C,H,W]
    g_logit_m = g_logit[:, :M]
    g_lpres = g_logit_m.sum((2, 3, 4))
    g_presence = None
    if presence is not None:
        g_presence = torch.where(tiny, torch.zeros_like(presence), g_lpres / torch.where(tiny, torch.ones_like(presence), presence))
    g_a = g_logit_m.sum(2) if alpha_mode else None                       # [B,M,H,W]

    # template / alpha gradients: transposed bilinear interpolation (scatter-add of g * weight into the 4 taps)
    g_templates = torch.zeros(B, M, C, h * w, dtype=dt)
    g_alpha = torch.zeros(M, h * w, dtype=dt) if alpha_mode else None
    for (tv, ok, flat), wt in zip(taps_t, wts):
        contrib = torch.where(ok, g_loc_m * wt.unsqueeze(2), torch.zeros_like(g_loc_m))
        g_templates.scatter_add_(-1, flat.reshape(B, M, C, -1), contrib.reshape(B, M, C, -1))
    if alpha_mode:
        for (tv, ok, flat), wt in zip(taps_a, wts):
            contrib = torch.where(ok, g_a * wt, torch.zeros_like(g_a))
            idx = flat.reshape(B, M, -1)
            g_alpha.scatter_add_(-1, idx.permute(1, 0, 2).reshape(M, -1), contrib.reshape(B, M, -1).permute(1, 0, 2).reshape(M, -1))
    # pose gradients: d loc / d ix = (ne - nw)(1 - fy) + (se - sw) fy ; d loc / d iy = (sw - nw)(1 - fx) + (se - ne) fx
    (nw, _, _), (ne, _, _), (sw, _, _), (se, _, _) = taps_t
    dldx = (ne - nw) * (1 - fy).unsqueeze(2) + (se - sw) * fy.unsqueeze(2)
    dldy = (sw - nw) * (1 - fx).unsqueeze(2) + (se - ne) * fx.unsqueeze(2)
    g_ix = (g_loc_m * dldx).sum(2)
    g_iy = (g_loc_m * dldy).sum(2)
    if alpha_mode:
        (nw, _, _), (ne, _, _), (sw, _, _), (se, _, _) = taps_a
        g_ix = g_ix + g_a * ((ne - nw) * (1 - fy) + (se - sw) * fy)
        g_iy = g_iy + g_a * ((sw - nw) * (1 - fx) + (se - ne) * fx)
    g_gx, g_gy = g_ix * (w / 2), g_iy * (h / 2)
    g_pose = torch.stack([(g_gx * Xs).sum((2, 3)), (g_gx * Ys).sum((2, 3)), g_gx.sum((2, 3)),
                          (g_gy * Xs).sum((2, 3)), (g_gy * Ys).sum((2, 3)), g_gy.sum((2, 3))], -1)
    return dict(log_prob=log_prob, g_templates=g_templates.view(B, M, C, h, w), g_pose=g_pose, g_presence=g_presence,
                g_bg_image=g_bg_image, g_alpha=g_alpha.view(M, h, w) if alpha_mode else None, g_scalars=g_scalars)


# =============================================================================================================
# loss head: capsule sparsity losses + classifier cross-entropies
# =============================================================================================================

LOG_SAFE_EPS, LOG_SAFE_FLOOR = 1e-16, -1e8


def _head_within(kind, x, within_constant):
    """Per-row within-example term t[B] and d t / d x [B,O] (object_decoder.py:431-470).  ``kind``: 'l2' |
    'entropy' | 'kl'."""
    rs = x.sum(1, keepdim=True)
    if kind == 'l2':
        d = rs - within_constant
        return (d * d).squeeze(1), (2.0 * d).expand_as(x)
    k = 1.0 if kind == 'entropy' else float(x.shape[1])
    den = rs + 1e-8
    p = x / den
    tiny = p * k < LOG_SAFE_EPS
    L = torch.where(tiny, torch.full_like(p, LOG_SAFE_FLOOR), torch.log(torch.where(tiny, torch.ones_like(p), p * k)))
    t = -(p * L).sum(1)
    gp = -(L + (~tiny).to(x.dtype))            # d/dp of -sum p log_safe(k p): log_safe passes no gradient when tiny
    dot = (gp * p).sum(1, keepdim=True)
    return t, (gp - dot) / den


def _head_between(kind, col, batch_size, n_classes):
    """Between-example term (scalar) and its gradient w.r.t. the column sums col[O]."""
    O = col.shape[0]
    if kind == 'l2':
        d = col - float(batch_size) / n_classes
        return (d * d).mean(), 2.0 * d / O
    t, g = _head_within(kind, col.unsqueeze(0), None)
    return -t[0], -g[0]                        # negated: this entropy is to be increased (object_decoder.py:469)


def _head_classifier(x, weight, bias, label):
    """mean cross_entropy(softmax(linear(x)), label) -- on softmax OUTPUTS, sic (stacked_capsule_auto_encoder.py:
    67-74, :281-282) -- with the gradients of the head's weight / bias; x is detached in the reference (:209-211)."""
    B = x.shape[0]
    p = torch.softmax(x @ weight.t() + bias, -1)
    u = torch.log_softmax(p, -1)
    onehot = F.one_hot(label, p.shape[1]).to(x.dtype)
    xe = -(u * onehot).sum(1).mean()
    r = torch.softmax(p, -1) - onehot
    dz = p * (r - (p * r).sum(1, keepdim=True)) / B
    return xe, p, dz.t() @ x, dz.sum(0)


def loss_head_forward_backward(caps_presence, posterior, label, weight, bias, n_classes, prior_type, posterior_type,
                               weights, within_constant=None, sparsity=True):
    """What csrc/loss_head.cu computes.  caps_presence [B,O], posterior [B,O,V] (posterior_mixing_prob), label [B]
    int64 | None, weight [K,O] / bias [K] of prior_classifier[0] (used for BOTH heads, sic :211), ``weights`` =
    (prior within, prior between, posterior within, posterior between).  Returns terms (6 scalars, 0 where off), the
    weighted total, the two class-probability matrices and the gradients of the total."""
    B, O, V = posterior.shape
    dt = posterior.dtype
    zero = torch.zeros((), dtype=dt)
    terms = [zero] * 6
    g_cp = torch.zeros_like(caps_presence)
    g_post = torch.zeros_like(posterior)
    g_w = torch.zeros_like(weight) if weight is not None else None
    g_b = torch.zeros_like(bias) if bias is not None else None
    mass = posterior.sum(-1)
    total = zero
    if sparsity:
        wc = float(O) / n_classes if (within_constant is None and n_classes) else within_constant
        for slot, (kind, x, c, scale) in enumerate(((prior_type, caps_presence, wc, 1.0),
                                                    (posterior_type, mass / V,
                                                     float(O) / n_classes if n_classes else None, 1.0 / V))):
            ww, wb = weights[2 * slot], weights[2 * slot + 1]
            t, gx = _head_within(kind, x, c)
            tb, gcol = _head_between(kind, x.sum(0), B, n_classes)
            terms[2 * slot], terms[2 * slot + 1] = t.mean(), tb
            total = total + ww * t.mean() + wb * tb
            g = (ww * gx / B + wb * gcol.unsqueeze(0)) * scale
            if slot == 0:
                g_cp = g
            else:
                g_post = g.unsqueeze(-1).expand(B, O, V).clone()
    probs = [None, None]
    if label is not None:
        for slot, x in enumerate((caps_presence, mass)):
            xe, p, gw, gb = _head_classifier(x, weight, bias, label)
            terms[4 + slot], probs[slot] = xe, p
            total = total + xe
            g_w, g_b = g_w + gw, g_b + gb
    return dict(terms=torch.stack(terms), total=total, prior_cls_prob=probs[0], posterior_cls_prob=probs[1],
                g_caps_presence=g_cp, g_posterior=g_post, g_weight=g_w, g_bias=g_b)
