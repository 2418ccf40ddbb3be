"""Oracle for hot path 1: template warp + per-pixel template-mixture Gaussian log-likelihood.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two independent restatements:

* ``decode`` / ``mixture_log_prob`` / ``mixture_mode`` / ``mixture_mean``: the reference's own op sequence
  (``F.affine_grid`` + ``F.grid_sample`` + Normal log-density + log-softmax + logsumexp) following
  part_decoder.py:152-243 and distributions.py:20-89, differentiable through autograd, any dtype.
* ``closed_form_log_prob``: numpy fp64, no ATen sampler -- explicit bilinear zero-padded sampling per
  SURVEY.md section 8(a).  It documents exactly what the CUDA kernel computes per (pixel, template).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def log_safe(t, eps=1e-16):
    """math_ops.py:18-22: log(t) where t >= eps, else the constant -1e8."""
    tiny = t < eps
    return torch.where(tiny, torch.full_like(t, -1e8), torch.log(torch.where(tiny, torch.ones_like(t), t)))


def decode(templates, pose, output_size, presence=None, bg_image=None, *, templates_alpha=None,
           temperature_logit=None, scale=None, bg_mixing_logit=None, bg_value=None):
    """Returns (loc [B,M+1,C,H,W], sigma [1], mixing_logits [B,M+1,1|C,H,W]).

    templates [B,M,C,h,w]; pose [B,M,6] used directly as the 2x3 theta (part_decoder.py:176);
    alpha mode iff ``templates_alpha`` [1,M,1,h,w] is given, else temperature mode (part_decoder.py:198-218);
    ``scale`` = raw learnt output scale or None for sigma=1 (part_decoder.py:220-223).
    """
    B, M, C, h, w = templates.shape
    H, W = output_size
    theta = pose.reshape(B * M, 2, 3)
    grid = F.affine_grid(theta, [B * M, C, H, W], align_corners=False)                 # part_decoder.py:181
    warped = F.grid_sample(templates.reshape(B * M, C, h, w), grid, align_corners=False)  # :182-183
    warped = warped.view(B, M, C, H, W)
    if bg_image is not None:                                                            # :189-193
        bg = bg_image.unsqueeze(1)
    else:
        bg = torch.sigmoid(bg_value).expand(B, 1, C, H, W)
    loc = torch.cat([warped, bg], 1)                                                    # :195
    if templates_alpha is not None:                                                     # :198-214
        alpha = templates_alpha.expand(B, M, 1, h, w).reshape(B * M, 1, h, w)
        logits = F.grid_sample(alpha, grid, align_corners=False).view(B, M, 1, H, W)
        bg_logit = F.softplus(bg_mixing_logit).expand(B, 1, 1, H, W)
        logits = torch.cat([logits, bg_logit], 1)
    else:                                                                               # :215-218
        temperature = F.softplus(temperature_logit + .5) + 1e-4
        logits = loc / temperature
    if scale is not None:
        sigma = F.softplus(scale) + 1e-4
    else:
        sigma = torch.ones(1, dtype=templates.dtype, device=templates.device)
    if presence is not None:                                                            # :225-231
        full = torch.cat([presence, presence.new_ones(B, 1)], 1)
        logits = logits + log_safe(full).view(B, M + 1, 1, 1, 1)
    return loc, sigma, logits


def mixture_log_prob(loc, sigma, logits, x):
    """distributions.py:41-48 with torch Normal.log_prob written out: [B,C,H,W]."""
    x = x.unsqueeze(1)
    comp = -((x - loc) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - HALF_LOG_2PI
    return torch.logsumexp(comp + F.log_softmax(logits, 1), 1)


def mixture_mean(loc, logits):
    """distributions.py:37-39."""
    return torch.sum(F.softmax(logits, 1) * loc, 1)


def mixture_mode(loc, sigma, logits, straight_through_gradient=False, maximum=False):
    """distributions.py:50-77."""
    mlp = F.log_softmax(logits, 1)
    if maximum:
        mlp = mlp + (-torch.log(sigma) - HALF_LOG_2PI)
    K = mlp.shape[1]
    mask = F.one_hot(mlp.argmax(1), K).movedim(-1, 1)
    if straight_through_gradient:
        soft = F.softmax(mlp, 1)
        mask = (mask - soft).detach() + soft
    return torch.sum(mask * loc, 1)


def image_log_likelihood(templates, pose, x, presence=None, bg_image=None, **params):
    """Per-pixel log-prob as evaluated by SCAE.loss (stacked_capsule_auto_encoder.py:220)."""
    loc, sigma, logits = decode(templates, pose, x.shape[-2:], presence, bg_image, **params)
    return mixture_log_prob(loc, sigma, logits, x)


# --------------------------------------------------------------------------------------------------------------
# closed form, numpy fp64
# --------------------------------------------------------------------------------------------------------------

def _bilinear_zero_pad(tex, ix, iy):
    """tex [h,w]; ix, iy [H,W] unnormalised source coordinates -> [H,W] (ATen grid_sampler_2d, zeros padding)."""
    h, w = tex.shape
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    fx = ix - x0
    fy = iy - y0
    out = np.zeros_like(ix)
    for dy, wy in ((0, 1.0 - fy), (1, fy)):
        for dx, wx in ((0, 1.0 - fx), (1, fx)):
            xx = x0 + dx
            yy = y0 + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            out += np.where(ok, tex[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0.0) * wx * wy
    return out


def _softplus(v):
    return np.logaddexp(0.0, v)


def closed_form_log_prob(templates, pose, x, presence=None, bg_image=None, *, templates_alpha=None,
                         temperature_logit=None, scale=None, bg_mixing_logit=None, bg_value=None):
    """numpy fp64 per-pixel log-prob [B,C,H,W]; the formula block of SURVEY.md section 8(a)."""
    T = np.asarray(templates, np.float64)
    P = np.asarray(pose, np.float64)
    X = np.asarray(x, np.float64)
    B, M, C, h, w = T.shape
    H, W = X.shape[-2:]
    xs = (2.0 * np.arange(W) + 1.0) / W - 1.0
    ys = (2.0 * np.arange(H) + 1.0) / H - 1.0
    Xg, Yg = np.meshgrid(xs, ys)
    sigma = 1.0 if scale is None else float(_softplus(np.asarray(scale, np.float64)).reshape(-1)[0]) + 1e-4
    alpha_mode = templates_alpha is not None
    if alpha_mode:
        A = np.asarray(templates_alpha, np.float64)[0, :, 0]
        bg_logit = float(_softplus(np.asarray(bg_mixing_logit, np.float64)).reshape(-1)[0])
    else:
        temperature = float(_softplus(np.asarray(temperature_logit, np.float64) + .5).reshape(-1)[0]) + 1e-4
    out = np.zeros((B, C, H, W))
    for b in range(B):
        comp = np.zeros((M + 1, C, H, W))
        logit = np.zeros((M + 1, C, H, W))
        for m in range(M):
            p = P[b, m]
            gx = p[0] * Xg + p[1] * Yg + p[2]
            gy = p[3] * Xg + p[4] * Yg + p[5]
            ix = ((gx + 1.0) * w - 1.0) / 2.0
            iy = ((gy + 1.0) * h - 1.0) / 2.0
            lp = 0.0
            if presence is not None:
                pr = float(np.asarray(presence, np.float64)[b, m])
                lp = -1e8 if pr < 1e-16 else math.log(pr)
            a = _bilinear_zero_pad(A[m], ix, iy) if alpha_mode else None
            for c in range(C):
                loc = _bilinear_zero_pad(T[b, m, c], ix, iy)
                comp[m, c] = -0.5 * ((X[b, c] - loc) / sigma) ** 2
                logit[m, c] = (a if alpha_mode else loc / temperature) + lp
        for c in range(C):
            bg = np.asarray(bg_image, np.float64)[b, c] if bg_image is not None else \
                1.0 / (1.0 + math.exp(-float(np.asarray(bg_value, np.float64).reshape(-1)[0])))
            comp[M, c] = -0.5 * ((X[b, c] - bg) / sigma) ** 2
            logit[M, c] = bg_logit if alpha_mode else bg / temperature
        comp = comp - math.log(sigma) - HALF_LOG_2PI
        num = np.logaddexp.reduce(comp + logit, axis=0)
        den = np.logaddexp.reduce(logit, axis=0)
        out[b] = num - den
    return out
