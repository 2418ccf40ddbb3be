"""A/B timing of the SCAE train step under the package's environment switches, in ONE process (one model, one cuDNN
autotune): for every variant the step is re-captured as a CUDA graph and replayed.

    python tools/ab_step.py [--batch 1024] [--steps 40] > gpurun_out/ab_step.jsonl

Switches (read at call time by the modules): SCAE_B200_LOSS_HEAD (csrc/loss_head.cu vs the PyTorch loss tail),
SCAE_B200_ATT_GEMM (1x1 attention convolution as a GEMM + channels-last pooling vs cuDNN + NCHW pooling),
SCAE_B200_CUDNN_FUSED_RELU (cuDNN's fused conv+bias+ReLU forward vs conv + scae_bias_act_fwd), SCAE_B200_ATT_SPLITK
(row chunks of the attention GEMM's weight gradient), SCAE_B200_CONV_GEMM (0 | dgrad | full | auto: GEMM-form passes of
the 3x3 convolutions, csrc/conv_cols.cu).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import model_params  # noqa: E402
from torch_scae_b200 import _lib, ddp, factory, graph  # noqa: E402

VARIANTS = [
    ('default', {}),
    ('loss_head_off', {'SCAE_B200_LOSS_HEAD': '0'}),
    ('att_gemm_off', {'SCAE_B200_ATT_GEMM': '0'}),
    ('both_off', {'SCAE_B200_LOSS_HEAD': '0', 'SCAE_B200_ATT_GEMM': '0'}),
    ('cudnn_fused_relu', {'SCAE_B200_CUDNN_FUSED_RELU': '1'}),
    ('conv_gemm_off', {'SCAE_B200_CONV_GEMM': '0'}),
    ('conv_gemm_dgrad_only', {'SCAE_B200_CONV_GEMM': 'dgrad'}),
    ('conv_gemm_full_everywhere', {'SCAE_B200_CONV_GEMM': 'full'}),
    ('conv_gemm_off_fused_relu', {'SCAE_B200_CONV_GEMM': '0', 'SCAE_B200_CUDNN_FUSED_RELU': '1'}),
    ('default_again', {}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--n-obj-caps', type=int, default=32)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    dev = torch.device('cuda', 0)
    torch.manual_seed(42)
    B = args.batch
    model = factory.make_scae(model_params(args.n_obj_caps)).to(dev).train()
    bucket = ddp.FlatGradBucket(model, assign=True, flat_params=True)
    opt = ddp.FlatRMSprop(bucket, lr=3e-5, momentum=0.9, eps=1e-2 / float(B) ** 2)
    image = torch.rand(B, 1, 40, 40, device=dev)
    label = torch.randint(0, 10, (B,), device=dev)
    lib = _lib.load()
    for name, env in VARIANTS:
        for k in ('SCAE_B200_LOSS_HEAD', 'SCAE_B200_ATT_GEMM', 'SCAE_B200_CUDNN_FUSED_RELU', 'SCAE_B200_ATT_SPLITK',
                  'SCAE_B200_CONV_GEMM'):
            os.environ.pop(k, None)
        os.environ.update(env)
        out = dict(variant=name, env=env, batch=B)
        try:
            before = lib.scae_launch_count()
            step = graph.GraphedTrainStep(model, opt, bucket, image, label)
            out['library_launches_per_capture_pass'] = (lib.scae_launch_count() - before) // 4   # 3 warm-ups + capture
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
            out.update(ms_per_step=round(ms, 4), images_per_s=round(B * 1000.0 / ms, 1), loss=float(step.loss))
            del step
        except Exception as exc:                                # noqa: BLE001 - report the variant as failed, go on
            out['error'] = f'{type(exc).__name__}: {exc}'[:400]
            torch.cuda.synchronize()
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
