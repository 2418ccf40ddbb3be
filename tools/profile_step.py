"""One SCAE train step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (B200_PROFILING.md).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'tmpl_ll|caps_ll|caps_bwd|reduce_rows' -o gpurun_out/prof python tools/profile_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import model_params  # noqa: E402
from torch_scae_b200 import ddp, factory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--n-obj-caps', type=int, default=32)
ap.add_argument('--steps', type=int, default=1)
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
torch.cuda.profiler.start()
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
