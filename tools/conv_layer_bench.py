"""Forward + backward time of ops.conv_bias_act on the part encoder's 128 -> 128 layers (B = 1024, strict fp32) per pass
formulation (SCAE_B200_CONV_GEMM = 0 | dgrad | full | auto), with the im2col / col2im kernel times.
    python tools/conv_layer_bench.py > gpurun_out/conv_layers.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_scae_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
B = 1024
for name, hin, stride in (('L2', 19, 2), ('L3', 9, 1), ('L4', 7, 1)):
    conv = torch.nn.Conv2d(128, 128, 3, stride).cuda()
    x = torch.randn(B, 128, hin, hin, device='cuda', requires_grad=True)
    for mode in ('0', 'dgrad', 'full', 'auto'):
        os.environ['SCAE_B200_CONV_GEMM'] = mode

        def step():
            y = ops.conv_bias_act(x, conv, True)
            y.backward(y)
            x.grad = conv.weight.grad = conv.bias.grad = None
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        with ops.KernelTimer() as t:
            step()
            torch.cuda.synchronize()
        ks = {k: round(v[2] / v[0], 4) for k, v in t.summary().items()}
        print(f'{name} {hin}x{hin} s{stride} mode={mode}: {ms:.3f} ms fwd+bwd  kernels {ks}', flush=True)
