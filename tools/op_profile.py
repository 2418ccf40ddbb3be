"""Operator-level view of one SCAE train step (torch.profiler): which ATen ops, with which shapes, own the device time
outside the four fused entry points.  Complements the ncu launch list (kernel names only).

    python tools/op_profile.py [--batch 1024] > gpurun_out/op_profile.txt
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import model_params  # noqa: E402
from torch_scae_b200 import ddp, factory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--n-obj-caps', type=int, default=32)
ap.add_argument('--rows', type=int, default=60)
args = ap.parse_args()

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device('cuda', 0)
torch.manual_seed(42)
model = factory.make_scae(model_params(args.n_obj_caps)).to(dev).train()
bucket = ddp.FlatGradBucket(model, assign=True, flat_params=True)
opt = ddp.FlatRMSprop(bucket, lr=3e-5, momentum=0.9, eps=1e-2 / float(args.batch) ** 2)
image = torch.rand(args.batch, 1, 40, 40, device=dev)
label = torch.randint(0, 10, (args.batch,), device=dev)


def step():
    bucket.zero()
    res = model(image)
    loss, _ = model.loss(res, image, label)
    loss.backward()
    bucket.collect()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by='self_cuda_time_total', row_limit=args.rows,
                                                         max_name_column_width=48, max_shapes_column_width=70))
print(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=40, max_name_column_width=60))
