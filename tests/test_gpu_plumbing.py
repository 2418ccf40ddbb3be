"""GPU tests of the small kernels that serve the callers of the hot paths (csrc/api.cu)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('rows,cols', [(40960, 16), (40960, 256), (32768, 256), (777, 32), (1000, 512), (5, 1),
                                       (3, 256), (100, 48)])
def test_colsum_matches_fp64_sum(rows, cols):
    """scae_colsum (bias gradients of the set transformer's tall-skinny linears) vs an fp64 column sum; (100, 48) is a
    shape the kernel does not cover and goes through torch.sum."""
    from torch_scae_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g)
    ref = x.double().sum(0)
    got = ops.colsum(x.cuda())
    assert got.shape == (cols,)
    assert rel_err(got, ref) < 1e-5
    again = ops.colsum(x.cuda())
    assert torch.equal(got, again)                           # fixed summation order


def test_skinny_linear_gradients():
    """split-K weight gradient + colsum bias gradient of skinny.linear vs autograd on nn.Linear in fp64."""
    from torch_scae_b200 import skinny
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    lin = torch.nn.Linear(16, 16).cuda()
    x = torch.randn(1024, 40, 16, device='cuda', requires_grad=True)
    w = torch.randn(1024, 40, 16, device='cuda')
    y = skinny.linear(x, lin)
    gx, gw, gb = torch.autograd.grad((y * w).sum(), [x, lin.weight, lin.bias])
    lin64 = torch.nn.Linear(16, 16).double()
    lin64.load_state_dict({k: v.double().cpu() for k, v in lin.state_dict().items()})
    x64 = x.detach().double().cpu().requires_grad_(True)
    rx, rw, rb = torch.autograd.grad((lin64(x64) * w.double().cpu()).sum(), [x64, lin64.weight, lin64.bias])
    assert rel_err(gx, rx) < 1e-5 and rel_err(gw, rw) < 1e-5 and rel_err(gb, rb) < 1e-5


# ---- fused plumbing kernels (csrc/support.cu) vs the plain PyTorch ops in fp64 -------------------------------------------
@pytest.mark.parametrize('shape', [(1024, 40, 16), (7, 3, 16), (129, 32), (5, 64), (1000, 8)])
@pytest.mark.parametrize('affine', [True, False])
def test_layer_norm_matches_torch(shape, affine):
    from torch_scae_b200 import ops
    d = shape[-1]
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 3 + 1).cuda().requires_grad_(True)
    w = (torch.rand(d, generator=g) + 0.5).cuda().requires_grad_(True) if affine else None
    b = torch.randn(d, generator=g).cuda().requires_grad_(True) if affine else None
    up = torch.randn(*shape, generator=g).cuda()
    y = ops.layer_norm(x, w, b, 1e-5)
    grads = torch.autograd.grad((y * up).sum(), [x] + ([w, b] if affine else []))
    x64 = x.detach().double().cpu().requires_grad_(True)
    w64 = w.detach().double().cpu().requires_grad_(True) if affine else None
    b64 = b.detach().double().cpu().requires_grad_(True) if affine else None
    y64 = torch.nn.functional.layer_norm(x64, (d,), w64, b64, 1e-5)
    ref = torch.autograd.grad((y64 * up.double().cpu()).sum(), [x64] + ([w64, b64] if affine else []))
    assert rel_err(y, y64) < 1e-5
    for got, want in zip(grads, ref):
        assert rel_err(got, want) < 1e-5
    y2 = ops.layer_norm(x, w, b, 1e-5)
    assert torch.equal(y, y2)


@pytest.mark.parametrize('N,cin,cout,hw,k,stride,relu', [(64, 1, 128, 40, 3, 2, True), (64, 128, 128, 19, 3, 2, True),
                                                         (33, 16, 24, 7, 3, 1, True), (64, 128, 960, 5, 1, 1, False),
                                                         (3, 4, 5, 6, 3, 1, False)])
def test_conv_bias_act_matches_torch(N, cin, cout, hw, k, stride, relu):
    from torch_scae_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(N + cout)
    conv = torch.nn.Conv2d(cin, cout, k, stride).cuda()
    x = torch.randn(N, cin, hw, hw, device='cuda', requires_grad=True)
    y = ops.conv_bias_act(x, conv, relu)
    up = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad((y * up).sum(), [x, conv.weight, conv.bias])
    conv64 = torch.nn.Conv2d(cin, cout, k, stride).double()
    conv64.load_state_dict({n: v.double().cpu() for n, v in conv.state_dict().items()})
    x64 = x.detach().double().cpu().requires_grad_(True)
    y64 = conv64(x64)
    if relu:
        y64 = torch.relu(y64)
    rx, rw, rb = torch.autograd.grad((y64 * up.double().cpu()).sum(), [x64, conv64.weight, conv64.bias])
    assert rel_err(y, y64) < 1e-5
    # a ReLU mask can flip where the fp32 pre-activation is within rounding of zero: compare in the l2 norm
    from conftest import l2_rel_err
    assert l2_rel_err(gx, rx) < 1e-4 and l2_rel_err(gw, rw) < 1e-4 and l2_rel_err(gb, rb) < 1e-4


@pytest.mark.parametrize('B,n,D,G', [(64, 40, 24, 5), (5, 3, 7, 8), (2, 1, 1, 1), (3, 2, 40, 6)])
def test_attention_pool_matches_reference_formula(B, n, D, G):
    from torch_scae_b200 import nn_ext
    g = torch.Generator().manual_seed(B * n + D)
    h = torch.randn(B, n * (D + 1), G, G, generator=g).cuda().requires_grad_(True)
    out = nn_ext.multiple_attention_pooling_2d(h, n)
    up = torch.randn(out.shape, generator=g).cuda()
    (gh,) = torch.autograd.grad((out * up).sum(), [h])
    h64 = h.detach().double().cpu().requires_grad_(True)
    grouped = h64.view(B, n, D + 1, G * G)                      # the reference formula (nn_ext.py:76-101)
    ref = (grouped[:, :, :-1] * torch.softmax(grouped[:, :, -1:], -1)).sum(-1).reshape(B, n * D, 1, 1)
    (rh,) = torch.autograd.grad((ref * up.double().cpu()).sum(), [h64])
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 1e-5 and rel_err(gh, rh) < 1e-5


def test_flat_rmsprop_kernel_matches_torch_rmsprop():
    from torch_scae_b200 import ddp
    torch.manual_seed(5)
    ref = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Tanh(), torch.nn.Linear(17, 5)).cuda()
    model = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Tanh(), torch.nn.Linear(17, 5)).cuda()
    model.load_state_dict(ref.state_dict())
    for momentum in (0.9, 0.0):
        ropt = torch.optim.RMSprop(ref.parameters(), lr=3e-3, momentum=momentum, eps=1e-4)
        bucket = ddp.FlatGradBucket(model, assign=True, flat_params=True)
        opt = ddp.FlatRMSprop(bucket, lr=3e-3, momentum=momentum, eps=1e-4)
        data, target = torch.randn(64, 33, device='cuda'), torch.randn(64, 5, device='cuda')
        for _ in range(5):
            ropt.zero_grad()
            ((ref(data) - target) ** 2).mean().backward()
            ropt.step()
            bucket.zero()
            ((model(data) - target) ** 2).mean().backward()
            bucket.collect()
            opt.step()
        for (k, a), b in zip(ref.state_dict().items(), model.state_dict().values()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (momentum, k)


def test_shared_query_attention_matches_expanded_queries():
    """MultiHeadQKVAttention with batch-shared queries (seeds.expand) projects them once; same result and gradients as
    the reference's per-image evaluation (set_transformer.py:24-71)."""
    from torch_scae_b200 import set_transformer as st
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    att = st.MultiHeadQKVAttention(d_k=32, d_v=32, n_heads=1).cuda()
    seeds = torch.randn(1, 6, 32, device='cuda', requires_grad=True)
    z = torch.randn(9, 11, 32, device='cuda', requires_grad=True)
    presence = torch.rand(9, 11, device='cuda')
    up = torch.randn(9, 6, 32, device='cuda')
    params = [seeds, z] + list(att.parameters())
    out = att(seeds.expand(9, -1, -1), z, z, presence)
    got = torch.autograd.grad((out * up).sum(), params)
    out_ref = att(seeds.expand(9, -1, -1).contiguous(), z, z, presence)
    ref = torch.autograd.grad((out_ref * up).sum(), params)
    assert rel_err(out, out_ref) < 1e-5
    for a, b in zip(got, ref):
        assert rel_err(a, b) < 1e-5


def test_set_transformer_pooled_output_matches_unfused_attention():
    """SetTransformer's re-associated output attention (no z / keys / values) vs the reference's op order
    (fc2 -> k/v projections -> attention -> output projection) on the same module: values and every gradient."""
    from torch_scae_b200 import set_transformer as st
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(4)
    net = st.SetTransformer(dim_in=24, dim_hidden=16, dim_out=64, n_outputs=8, n_layers=2, n_heads=1,
                            layer_norm=True).cuda()
    x = torch.randn(12, 10, 24, device='cuda', requires_grad=True)
    presence = torch.rand(12, 10, device='cuda')
    up = torch.randn(12, 8, 64, device='cuda')
    params = [x] + list(net.parameters())
    out = net(x, presence)
    got = torch.autograd.grad((out * up).sum(), params)

    def unfused(x, presence):
        h = net.fc1(x)
        for block in net.sabs:
            h = block(h, presence)
        z = net.fc2(h)
        return net.multi_head_attention(net.seeds.expand(x.shape[0], -1, -1).contiguous(), z, z, presence)
    out_ref = unfused(x, presence)
    ref = torch.autograd.grad((out_ref * up).sum(), params)
    assert rel_err(out, out_ref) < 1e-5
    for (name, _), a, b in zip([('x', None)] + list(net.named_parameters()), got, ref):
        assert rel_err(a, b) < 5e-5, name


@pytest.mark.parametrize('B,N,pres', [(5, 40, 'rand'), (3, 24, 'ones'), (2, 64, 'none'), (4, 7, 'binary'), (300, 40, 'rand')])
def test_fused_set_attention_block_matches_pytorch_ops(B, N, pres):
    """csrc/sab.cu (one kernel per direction) vs the same SAB evaluated with PyTorch ops in fp64 on the CPU: output, input
    gradient and all 14 parameter gradients.  'rand' presences exercise the reference's -(1-p) 1e32 mask quirk."""
    import copy
    from torch_scae_b200 import ops, set_transformer as st
    torch.manual_seed(B * 100 + N)
    sab = st.SAB(d=16, n_heads=1, layer_norm=True)
    with torch.no_grad():
        for p in sab.parameters():                      # non-trivial LayerNorm affine parameters and biases
            p.add_(torch.randn_like(p) * 0.3)
    x0 = torch.randn(B, N, 16)
    presence = dict(rand=torch.rand(B, N), ones=torch.ones(B, N), none=None,
                    binary=(torch.rand(B, N) > 0.4).float())[pres]
    up = torch.randn(B, N, 16)

    ref_mod = copy.deepcopy(sab).double()
    x64 = x0.double().requires_grad_(True)
    y64 = ref_mod.mab(x64, x64, presence.double() if presence is not None else None)
    ref = torch.autograd.grad((y64 * up.double()).sum(), [x64] + list(ref_mod.parameters()))

    gpu = copy.deepcopy(sab).cuda()
    x = x0.cuda().requires_grad_(True)
    pr = presence.cuda() if presence is not None else None
    y = ops.set_attention_block(x, pr, gpu.mab)
    assert y is not None, 'fused path not taken'
    got = torch.autograd.grad((y * up.cuda()).sum(), [x] + list(gpu.parameters()))
    assert rel_err(y, y64) < 2e-5
    # the key-bias gradient is identically zero (softmax is invariant to a per-row constant): measure every error against
    # the gradient's own scale, floored at a fraction of the largest gradient in the block
    floor = 1e-2 * max(float(b.abs().max()) for b in ref)
    for (name, _), a, b in zip([('x', None)] + list(gpu.named_parameters()), got, ref):
        err = float((a.double().cpu() - b).abs().max()) / max(float(b.abs().max()), floor)
        assert err < 1e-4, (name, err)
    y2 = gpu(x, pr)                                     # the module routes through the fused path, bit-reproducibly
    assert torch.equal(y, y2)


@pytest.mark.parametrize('similarity', [False, True])
def test_pose_transform_matches_reference_formula(similarity):
    """scae_pose_transform vs cv_ops.geometric_transform's PyTorch formula (cv_ops.py:20-76) in fp64."""
    from torch_scae_b200 import cv_ops
    g = torch.Generator().manual_seed(9)
    t0 = torch.randn(37, 40, 6, generator=g) * 0.7
    up = torch.randn(37, 40, 6, generator=g)
    t = t0.cuda().requires_grad_(True)
    out = cv_ops.geometric_transform(t, similarity)
    (gt,) = torch.autograd.grad((out * up.cuda()).sum(), [t])
    t64 = t0.double().requires_grad_(True)
    ref = cv_ops.geometric_transform(t64, similarity)              # CPU tensor: the PyTorch formula
    (rt,) = torch.autograd.grad((ref * up.double()).sum(), [t64])
    assert rel_err(out, ref) < 1e-5 and rel_err(gt, rt) < 1e-5
