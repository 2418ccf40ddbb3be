"""N>1 host logic on CPU: two gloo ranks, FlatGradBucket all-reduce == averaging the per-shard gradients, parameters
stay identical across ranks after an optimizer step, and gradient views keep aliasing the flat bucket."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from torch_scae_b200 import ddp
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(1234 + rank)                       # deliberately different initial weights per rank
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
    ddp.broadcast_parameters(model)                      # ... made identical by the broadcast
    bucket = ddp.FlatGradBucket(model)
    opt = torch.optim.RMSprop(model.parameters(), lr=1e-2, momentum=0.9)
    torch.manual_seed(99)
    data = torch.randn(8, 6)
    target = torch.randn(8, 3)
    shard = slice(rank * 4, rank * 4 + 4)                # batch sharded by image, no data-path collective
    for _ in range(3):
        bucket.zero()
        loss = ((model(data[shard]) - target[shard]) ** 2).mean()
        loss.backward()
        assert bucket.check_views()
        bucket.all_reduce_mean()
        opt.step()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        # single-process reference on the full batch (mean over shards of shard-mean losses == full-batch mean here)
        torch.manual_seed(1234)
        ref = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
        ropt = torch.optim.RMSprop(ref.parameters(), lr=1e-2, momentum=0.9)
        for _ in range(3):
            ropt.zero_grad()
            ((ref(data) - target) ** 2).mean().backward()
            ropt.step()
        ref_flat = torch.cat([p.detach().flatten() for p in ref.parameters()])
        torch.save(dict(same=bool(torch.equal(gathered[0], gathered[1])),
                        err=float((gathered[0] - ref_flat).abs().max())), out)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2(tmp_path):
    out = str(tmp_path / 'result.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res['same'], 'ranks diverged'
    assert res['err'] < 1e-6, res
