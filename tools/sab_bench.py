"""Times (or, under ncu, exposes) the fused set-attention-block kernels at the train-step shape (B, 40, 16)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_scae_b200 import ops, set_transformer as st  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--tokens', type=int, default=40)
ap.add_argument('--iters', type=int, default=20)
args = ap.parse_args()
torch.manual_seed(0)
sab = st.SAB(d=16, n_heads=1, layer_norm=True).cuda()
x = torch.randn(args.batch, args.tokens, 16, device='cuda', requires_grad=True)
presence = torch.rand(args.batch, args.tokens, device='cuda')
up = torch.randn_like(x)


def once():
    y = sab(x, presence)
    torch.autograd.grad((y * up).sum(), [x] + list(sab.parameters()))


for _ in range(3):
    once()
torch.cuda.synchronize()
with ops.KernelTimer() as t:
    for _ in range(args.iters):
        once()
    torch.cuda.synchronize()
for k, (calls, launches, ms) in t.summary().items():
    print(f'{k}: {1000 * ms / calls:.1f} us per call ({launches // calls} launches)')
