"""Times the CNN part encoder (fwd+bwd, strict fp32) in NCHW vs channels_last to pick the faster layout."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_scae_b200.part_encoder import CNNEncoder  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = 'cuda'
for fmt in (torch.contiguous_format, torch.channels_last):
    enc = CNNEncoder((1, 40, 40), [128] * 4, [3] * 4, [2, 2, 1, 1]).to(dev).to(memory_format=fmt)
    x = torch.rand(1024, 1, 40, 40, device=dev).contiguous(memory_format=fmt)
    for _ in range(5):
        enc(x).sum().backward()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        enc(x).sum().backward()
    e.record()
    torch.cuda.synchronize()
    print(fmt, f'{s.elapsed_time(e) / 20:.3f} ms per fwd+bwd')
