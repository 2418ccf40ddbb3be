"""Micro-benchmark of the four C-ABI entry points alone, at every BASELINE.json config shape and several batch sizes.

    python tools/kernel_bench.py [--configs mnist32,mnist10,stress,color] [--batches 1024,8192] [--iters 20]

Inputs follow SURVEY.md section 8(d) (realistic pose range, presences in (0,1), ReLU'd capsule parameters).  Each
iteration is preceded by an L2 flush (a 512 MB fill) so that inputs come from HBM even when the working set would fit
the 126 MB L2; the timed region is the CUDA-event pair `ops.KernelTimer` puts around the C-ABI call only.  Prints one
JSON line per (config, batch): ms, algorithmic GB/s and fraction of the measured HBM copy peak per entry point.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bench import algorithmic_bytes, measured_peaks  # noqa: E402
from torch_scae_b200 import _lib, ops  # noqa: E402
from torch_scae_b200.cv_ops import geometric_transform  # noqa: E402

CONFIGS = dict(
    mnist32=dict(M=40, C=1, h=11, w=11, H=40, W=40, O=32),     # BASELINE configs[0..2] with mnist.yaml's O
    mnist10=dict(M=40, C=1, h=11, w=11, H=40, W=40, O=10),     # ... with BASELINE's parenthetical O
    stress=dict(M=64, C=1, h=21, w=21, H=64, W=64, O=32),      # configs[3]
    color=dict(M=24, C=3, h=11, w=11, H=32, W=32, O=32),       # configs[4]
)


def run(name, cfg, B, iters, flush, no_flush=False, only='all'):
    dev = 'cuda'
    M, C, h, w, H, W, O = (cfg[k] for k in ('M', 'C', 'h', 'w', 'H', 'W', 'O'))
    V, A = M, 8 * M + 7
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g, device=dev)
    n = lambda *s: torch.randn(*s, generator=g, device=dev)
    # path 1
    templates = r(B, M, C, h, w).requires_grad_(True)
    pose = geometric_transform(0.5 * n(B, M, 6), similarity=False).requires_grad_(True)
    presence = r(B, M).requires_grad_(True)
    x = r(B, C, H, W)
    alpha = n(1, M, 1, h, w).requires_grad_(True)
    bg_value = torch.zeros(1, device=dev, requires_grad=True)
    bg_logit = torch.zeros(1, device=dev, requires_grad=True)
    # path 2
    all_param = torch.relu(n(B, O, A)).requires_grad_(True)
    cpr_static = (0.1 * n(1, O, V, 6)).requires_grad_(True)
    biases = [torch.zeros(s, device=dev, requires_grad=True) for s in ((1, O, 1, 6), (1, O, 1), (1, O, V), (1, O, V))]
    dummy = torch.zeros(1, 1, V, 6, device=dev)
    px = geometric_transform(0.5 * n(B, V, 6), similarity=False)
    ppres = r(B, V)
    noise_caps, noise_vote = (r(B, O, 1) - .5) * 4, (r(B, O, V) - .5) * 4
    flags = _lib.CAPS_LEARN_VOTE_SCALE | _lib.CAPS_ALLOW_DEFORM
    up_post, up_cp = n(B, O, V), n(B, O)

    def once():
        if only in ('all', 'tmpl'):
            lp, ll = ops.TemplateMixtureLogProb.apply(templates, pose, presence, None, x, alpha, bg_value, bg_logit, None,
                                                      None, (H, W))
            if not no_flush:
                flush.add_(1.0)
            ll.sum().backward()
            if not no_flush:
                flush.add_(1.0)
        if only == 'tmpl':
            for t in (templates, pose, presence, alpha, bg_value, bg_logit):
                t.grad = None
            return
        res = dict(zip(ops.CAPS_RETURNS, ops.CapsuleVoteLikelihood.apply(
            all_param, cpr_static, *biases, dummy, px, ppres, noise_caps, noise_vote, flags)))
        if not no_flush:
            flush.add_(1.0)
        (res['ll_per_example'].sum() + res['reg_per_example'].sum() + (res['posterior_mixing_prob'] * up_post).sum()
         + (res['caps_presence'] * up_cp).sum()).backward()
        for t in (templates, pose, presence, alpha, bg_value, bg_logit, all_param, cpr_static, *biases):
            t.grad = None
        if not no_flush:
            flush.add_(1.0)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    with ops.KernelTimer() as timer:
        for _ in range(iters):
            once()
        torch.cuda.synchronize()
    peak, _ = measured_peaks()
    bpi = algorithmic_bytes(M=M, C=C, h=h, w=w, H=H, W=W, O=O)
    out = dict(config=name, B=B, l2_flush=not no_flush, **cfg)
    for k, (calls, _, total_ms) in timer.summary().items():
        ms = total_ms / calls
        gbs = bpi[k] * B / (ms * 1e-3) / 1e9
        out[k] = dict(ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 4),
                      mimg_s=round(B / ms / 1e3, 3))
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='mnist32,mnist10,stress,color')
    ap.add_argument('--batches', default='1024,8192')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--only', default='all', choices=('all', 'caps', 'tmpl'))
    args = ap.parse_args()
    flush = torch.empty(128 * 1024 * 1024, device='cuda')    # 512 MB > 126 MB L2
    for name in args.configs.split(','):
        for B in (int(b) for b in args.batches.split(',')):
            run(name, CONFIGS[name], B, args.iters, flush, args.no_flush, args.only)


if __name__ == '__main__':
    main()
