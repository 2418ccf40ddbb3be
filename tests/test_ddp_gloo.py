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


# ---- batch-global sparsity statistics: sync_batch_stats=True == single process, global batch ----
def _sparsity_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from torch_scae_b200 import object_decoder as od
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(5)
    full = torch.rand(12, 7, dtype=torch.float64)                  # caps_presence of the global batch (B=12, O=7)
    weights = torch.rand(4, dtype=torch.float64)
    results = {}
    for kind in ('l2', 'entropy', 'kl'):
        # sharded: each rank sees 6 examples; losses weighted like SCAE.loss, gradients averaged like FlatGradBucket
        local = full[rank * 6:(rank + 1) * 6].clone().requires_grad_(True)
        within, between = od.sparsity_loss(kind, local, n_classes=3, sync_batch_stats=True)
        (weights[0] * within + weights[1] * between).backward()
        grad = local.grad / world                                  # all_reduce_mean of parameter gradients
        parts = [torch.zeros_like(grad) for _ in range(world)]
        dist.all_gather(parts, grad)
        stats = torch.stack([within.detach(), between.detach()])
        dist.all_reduce(stats)                                     # within: mean of shard means; between: same on all ranks
        if rank == 0:
            ref_in = full.clone().requires_grad_(True)
            rw, rb = od.sparsity_loss(kind, ref_in, n_classes=3)
            (weights[0] * rw + weights[1] * rb).backward()
            results[kind] = dict(
                grad_err=float((torch.cat(parts) - ref_in.grad).abs().max() / ref_in.grad.abs().max()),
                within_err=float((stats[0] / world - rw).abs()), between_err=float((stats[1] / world - rb).abs()))
    # and the default (per-shard statistics, the reference's behaviour under Lightning DDP) stays collective-free
    local = full[rank * 6:(rank + 1) * 6].clone()
    _, b_local = od.sparsity_loss('l2', local, n_classes=3)
    if rank == 0:
        results['local_between'] = float(b_local)
        results['expected_local_between'] = float(torch.mean((local.sum(0) - 6 / 3) ** 2))
        torch.save(results, out)
    dist.destroy_process_group()


def test_sync_batch_stats_world2(tmp_path):
    out = str(tmp_path / 'sparsity.pt')
    mp.spawn(_sparsity_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    for kind in ('l2', 'entropy', 'kl'):
        assert res[kind]['grad_err'] < 1e-12, (kind, res[kind])
        assert res[kind]['within_err'] < 1e-12 and res[kind]['between_err'] < 1e-12, (kind, res[kind])
    assert abs(res['local_between'] - res['expected_local_between']) < 1e-12


# ---- flat parameter / gradient / optimizer-state buffers (single process, CPU arithmetic path) --------------------------
def test_flat_rmsprop_assign_mode_matches_torch_rmsprop():
    from torch_scae_b200 import ddp
    torch.manual_seed(3)
    ref = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3), nn.Linear(3, 2))
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3), nn.Linear(3, 2))
    model.load_state_dict(ref.state_dict())
    for m in (ref, model):
        m[3].weight.requires_grad_(True)
    ropt = torch.optim.RMSprop(ref.parameters(), lr=1e-2, momentum=0.9, eps=1e-3)
    bucket = ddp.FlatGradBucket(model, assign=True, flat_params=True)
    opt = ddp.FlatRMSprop(bucket, lr=1e-2, momentum=0.9, eps=1e-3)
    data, target = torch.randn(8, 6), torch.randn(8, 3)
    for _ in range(4):
        ropt.zero_grad()
        ((ref[:3](data) - target) ** 2).mean().backward()        # the last Linear gets no gradient at all
        ropt.step()
        bucket.zero()
        ((model[:3](data) - target) ** 2).mean().backward()
        bucket.collect()
        assert bucket.check_views()
        bucket.all_reduce_mean()
        opt.step()
    for (k, a), b in zip(ref.state_dict().items(), model.state_dict().values()):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), k
    # parameters are views of one flat buffer, state_dict round-trips through them
    base = bucket.flat_param.untyped_storage().data_ptr()
    assert all(p.untyped_storage().data_ptr() == base for p in model.parameters())
    model.load_state_dict(ref.state_dict())
    assert all(p.untyped_storage().data_ptr() == base for p in model.parameters())


def test_flat_bucket_parameter_unused_in_a_later_step():
    """A parameter with a gradient in step 1 and none in step 2 must not see its old gradient again (torch.optim skips
    p.grad None: parameter, square average and momentum stay put)."""
    from torch_scae_b200 import ddp
    torch.manual_seed(5)
    ref = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3), nn.Linear(3, 2))
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3), nn.Linear(3, 2))
    model.load_state_dict(ref.state_dict())
    ropt = torch.optim.RMSprop(ref.parameters(), lr=1e-2, momentum=0.9, eps=1e-3)
    bucket = ddp.FlatGradBucket(model, assign=True, flat_params=True)
    opt = ddp.FlatRMSprop(bucket, lr=1e-2, momentum=0.9, eps=1e-3)
    data, t3, t2 = torch.randn(8, 6), torch.randn(8, 3), torch.randn(8, 2)
    for use_last in (True, False, False, True):
        ropt.zero_grad()
        out = ref(data) if use_last else ref[:3](data)
        ((out - (t2 if use_last else t3)) ** 2).mean().backward()
        ropt.step()
        bucket.zero()
        out = model(data) if use_last else model[:3](data)
        ((out - (t2 if use_last else t3)) ** 2).mean().backward()
        bucket.collect()
        bucket.all_reduce_mean()
        opt.step()
        for (k, a), b in zip(ref.state_dict().items(), model.state_dict().values()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), (k, use_last)


# ---- the same through the loss-head KERNELS on two GPUs: NCCL, skipped with fewer devices ---------
def _loss_head_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from torch_scae_b200 import ops
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(7)
    B, O, V, K = 16, 10, 8, 5
    cp_full, post_full = torch.rand(B, O), torch.rand(B, O, V)
    label_full = torch.randint(0, K, (B,))
    head = nn.Linear(O, K)
    weights = (0.7, 0.9, 1.3, 0.4)
    results = {}
    for pt, qt in (('l2', 'entropy'), ('entropy', 'kl'), ('kl', 'l2')):
        sl = slice(rank * B // world, (rank + 1) * B // world)
        cp = cp_full[sl].to(dev).requires_grad_(True)
        post = post_full[sl].to(dev).requires_grad_(True)
        h = nn.Linear(O, K).to(dev)
        h.load_state_dict(head.state_dict())
        total, terms, _ = ops.loss_head(cp, post, label_full[sl].to(dev), h, K, pt, qt, weights, None, True,
                                        sync_batch_stats=True)
        total.backward()
        grads = torch.cat([cp.grad.flatten(), post.grad.flatten()]) / world      # FlatGradBucket.all_reduce_mean
        parts = [torch.zeros_like(grads) for _ in range(world)]
        dist.all_gather(parts, grads)
        tsum = total.detach().clone()
        dist.all_reduce(tsum)
        if rank == 0:
            cpr = cp_full.to(dev).requires_grad_(True)
            postr = post_full.to(dev).requires_grad_(True)
            tot_ref, terms_ref, _ = ops.loss_head(cpr, postr, label_full.to(dev), h, K, pt, qt, weights, None, True)
            tot_ref.backward()
            n_cp = cp.numel()
            g_cp = torch.cat([p[:n_cp].view(-1, O) for p in parts])
            g_post = torch.cat([p[n_cp:].view(-1, O, V) for p in parts])
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            results[pt + '/' + qt] = dict(total=rel(tsum / world, tot_ref.detach()), g_cp=rel(g_cp, cpr.grad),
                                          g_post=rel(g_post, postr.grad),
                                          between=rel(terms[[1, 3]], terms_ref[[1, 3]]))
    if rank == 0:
        torch.save(results, out)
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices')
def test_sync_batch_stats_loss_head_kernels_world2(tmp_path):
    """sync_batch_stats=True through scae_loss_head_fwd_rows -> all-reduce -> _finish and the world-scaled backward: two
    ranks on half the batch each reproduce the single-process loss head on the whole batch."""
    out = str(tmp_path / 'loss_head.pt')
    mp.spawn(_loss_head_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    for kind, r in res.items():
        assert r['between'] < 1e-5 and r['g_cp'] < 1e-4 and r['g_post'] < 1e-4, (kind, r)
        # the within terms and cross-entropies are means over the shard: their rank average is the global mean
        assert r['total'] < 1e-5, (kind, r)
