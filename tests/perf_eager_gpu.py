"""GPU "before" number: the reference's op sequence (oracle port: per-capsule MLP loops,
affine_grid + grid_sample, materialised B x K x H x W tensors, Normal.log_prob chain) run with stock PyTorch CUDA ops on
the same B200, next to the fused path.  Not a pytest module (the oracle may only be executed from tests/); prints one
JSON line.

    python tests/perf_eager_gpu.py [--batch 1024] [--steps 10] [--tf32]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import model_params  # noqa: E402
from oracle import scae_model  # noqa: E402
from torch_scae_b200 import factory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--n-obj-caps', type=int, default=32)
ap.add_argument('--tf32', action='store_true', help="PyTorch's stock defaults (TF32 convolutions) instead of strict fp32")
args = ap.parse_args()

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = bool(args.tf32)
torch.backends.cudnn.benchmark = True
dev = torch.device('cuda', 0)
torch.manual_seed(42)
B = args.batch
params = model_params(args.n_obj_caps)
cfg = factory.prepare_model_params(**params)
sd = {k: v.detach().clone().to(dev).requires_grad_(v.is_floating_point())
      for k, v in factory.make_scae(params).state_dict().items()}
leaves = [v for v in sd.values() if v.requires_grad]
opt = torch.optim.RMSprop(leaves, lr=3e-5, momentum=0.9, eps=1e-2 / float(B) ** 2, foreach=True)
image = torch.rand(B, 1, 40, 40, device=dev)
label = torch.randint(0, 10, (B,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    res = scae_model.scae_forward(sd, cfg, image, None, training=True)
    loss, _ = scae_model.scae_loss(res, cfg, image, label)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.steps):
    step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / args.steps
print(json.dumps(dict(what='reference op sequence (oracle port), stock PyTorch CUDA ops, eager', batch=B,
                      n_obj_caps=args.n_obj_caps, cudnn_tf32=bool(args.tf32), ms_per_step=round(ms, 3),
                      images_per_s=round(B * 1000.0 / ms, 1),
                      peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))))
