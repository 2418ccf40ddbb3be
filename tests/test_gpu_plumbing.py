"""GPU tests of the small kernels that serve the callers of the hot paths (csrc/api.cu)."""
import pytest
import torch

from conftest import rel_err
from gpu_util import DEV

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
                                                         (3, 4, 5, 6, 3, 1, False), (37, 128, 128, 9, 3, 1, True),
                                                         (16, 128, 128, 7, 3, 1, True), (5, 40, 33, 8, 3, 2, False)])
@pytest.mark.parametrize('gemm', ['auto', '0', 'dgrad', 'full'])
def test_conv_bias_act_matches_torch(N, cin, cout, hw, k, stride, relu, gemm, monkeypatch):
    """ops.conv_bias_act vs nn.Conv2d (+ ReLU) in fp64 for every pass formulation: cuDNN (`0`), GEMM-form data gradient
    (`dgrad`), GEMM-form everything (`full`: im2col + SGEMM forward, split-K weight gradient) and the shipped rule."""
    from torch_scae_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    monkeypatch.setenv('SCAE_B200_CONV_GEMM', gemm)
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


def _loss_head_reference(cp, post, label, weight, bias, K, prior_type, posterior_type, ws, sparsity):
    """The PyTorch ops of SCAE.loss (stacked_capsule_auto_encoder.py:243-285) in fp64 on the CPU, with autograd."""
    import torch.nn.functional as F
    from torch_scae_b200.object_decoder import sparsity_loss
    leaves = [t.double().clone().requires_grad_(True) for t in (cp, post, weight, bias)]
    a, b, w_, b_ = leaves
    V = post.shape[-1]
    total, terms = 0.0, [torch.zeros((), dtype=torch.float64)] * 6
    if sparsity:
        pw, pb = sparsity_loss(prior_type, a, n_classes=K, within_example_constant=None)
        qw, qb = sparsity_loss(posterior_type, b.sum(-1) / V, n_classes=K)
        total = ws[0] * pw + ws[1] * pb + ws[2] * qw + ws[3] * qb
        terms[:4] = [pw, pb, qw, qb]
    probs = None
    if label is not None:
        p1 = torch.softmax(F.linear(a.detach(), w_, b_), -1)
        p2 = torch.softmax(F.linear(b.sum(-1).detach(), w_, b_), -1)
        terms[4:] = [F.cross_entropy(p1, label), F.cross_entropy(p2, label)]
        total = total + terms[4] + terms[5]
        probs = torch.stack([p1, p2])
    grads = torch.autograd.grad(total * 1.7, leaves, allow_unused=True)
    return total, torch.stack([t.detach() for t in terms]), probs, grads


@pytest.mark.parametrize('B,O,V,K', [(37, 10, 40, 10), (1024, 32, 40, 10), (64, 40, 24, 16), (5, 64, 7, 3),
                                      (300, 33, 6, 2)])
@pytest.mark.parametrize('prior_type,posterior_type', [('l2', 'entropy'), ('entropy', 'kl'), ('kl', 'l2')])
@pytest.mark.parametrize('with_label', [True, False])
def test_loss_head_matches_pytorch_ops(B, O, V, K, prior_type, posterior_type, with_label):
    """scae_loss_head_fwd/bwd vs the stock ops they replace (object_decoder.py:431-493 sparsity losses, classifier heads
    and cross-entropies on softmax outputs, stacked_capsule_auto_encoder.py:203-213,:279-285), values and gradients."""
    from torch_scae_b200 import ops
    g = torch.Generator().manual_seed(B + O + V)
    cp = torch.rand(B, O, generator=g)
    post = torch.rand(B, O, V, generator=g) / O
    label = torch.randint(0, K, (B,), generator=g) if with_label else None
    lin = torch.nn.Linear(O, K)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(K, O, generator=g) * 0.3)
        lin.bias.copy_(torch.randn(K, generator=g) * 0.3)
    ws = (2.0, 0.35, 0.7, 0.2)
    ref_total, ref_terms, ref_probs, ref_grads = _loss_head_reference(
        cp, post, label, lin.weight.detach(), lin.bias.detach(), K, prior_type, posterior_type, ws, True)

    lin = lin.cuda()
    cpd, postd = cp.cuda().requires_grad_(True), post.cuda().requires_grad_(True)
    out = ops.loss_head(cpd, postd, label.cuda() if with_label else None, lin if with_label else None, K, prior_type,
                        posterior_type, ws)
    assert out is not None
    total, terms, probs = out
    assert rel_err(total, ref_total) < 1e-5
    # absolute floor: the kl between-example term is a near-cancelling sum (~2e-4 left of O(1) summands), so fp32 leaves
    # ~1e-7 of absolute error on it whatever the implementation
    for i in range(6 if with_label else 4):
        assert abs(float(terms[i]) - float(ref_terms[i])) <= 1e-5 * abs(float(ref_terms[i])) + 1e-6, i
    assert rel_err(terms[6], ref_total) < 1e-5
    wanted = [cpd, postd] + ([lin.weight, lin.bias] if with_label else [])
    grads = torch.autograd.grad(total * 1.7, wanted)
    for name, got, ref in zip(('caps_presence', 'posterior', 'weight', 'bias'), grads, ref_grads):
        assert rel_err(got, ref) < 1e-4, name
    if with_label:
        assert rel_err(probs, ref_probs) < 1e-5
    # bit-reproducible (fixed summation orders)
    total2, terms2, _ = ops.loss_head(cpd, postd, label.cuda() if with_label else None, lin if with_label else None, K,
                                      prior_type, posterior_type, ws)
    assert torch.equal(terms, terms2) and torch.equal(total, total2)


def test_loss_head_log_safe_floor_and_classifier_only():
    """Zero capsule presences hit log_safe's -1e8 floor in the entropy losses (math_ops.py:18-22); and with both prior
    weights 0 the reference skips the sparsity terms (stacked_capsule_auto_encoder.py:243) but keeps the classifiers."""
    from torch_scae_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, O, V, K = 19, 12, 8, 10
    cp = torch.rand(B, O, generator=g)
    cp[3, 5] = 0.0
    post = torch.rand(B, O, V, generator=g) / O
    post[7, 2] = 0.0
    label = torch.randint(0, K, (B,), generator=g)
    lin = torch.nn.Linear(O, K)
    ws = (1.0, 1.0, 1.0, 1.0)
    for sparsity in (True, False):
        ref_total, ref_terms, _, ref_grads = _loss_head_reference(cp, post, label, lin.weight.detach(),
                                                                  lin.bias.detach(), K, 'entropy', 'entropy', ws, sparsity)
        lin_d = torch.nn.Linear(O, K).cuda()
        lin_d.load_state_dict(lin.state_dict())
        cpd, postd = cp.cuda().requires_grad_(True), post.cuda().requires_grad_(True)
        total, terms, _ = ops.loss_head(cpd, postd, label.cuda(), lin_d, K, 'entropy', 'entropy', ws, sparsity=sparsity)
        assert rel_err(total, ref_total) < 1e-5
        for i in range(6):
            assert abs(float(terms[i]) - float(ref_terms[i])) <= 1e-5 * abs(float(ref_terms[i])) + 1e-6, i
        grads = torch.autograd.grad(total * 1.7, [cpd, postd, lin_d.weight, lin_d.bias], allow_unused=True)
        for name, got, ref in zip(('caps_presence', 'posterior', 'weight', 'bias'), grads, ref_grads):
            if ref is None:
                assert got is None or float(got.abs().max()) == 0.0, name
            else:
                assert rel_err(got, ref) < 1e-4, name


def test_scae_loss_uses_the_loss_head_and_matches_the_pytorch_tail():
    """SCAE.loss through the fused loss head vs the same model with the head disabled (stock PyTorch ops): loss, log
    entries, class probabilities and every parameter gradient."""
    from golden.cases import tiny_model_params
    from torch_scae_b200 import factory
    from torch_scae_b200.stacked_capsule_auto_encoder import SCAE
    torch.manual_seed(5)
    model = factory.make_scae(tiny_model_params()).cuda().train()
    image = torch.rand(6, 1, 20, 20, device='cuda')
    label = torch.randint(0, 10, (6,), device='cuda')
    M, O = model.part_encoder.n_caps, model.obj_decoder.n_obj_capsules
    noise = dict(part_presence=(torch.rand(6, M, device='cuda') - .5) * 4,
                 caps=(torch.rand(6, O, 1, device='cuda') - .5) * 4, vote=(torch.rand(6, O, M, device='cuda') - .5) * 4)

    def run(fused):
        model.zero_grad(set_to_none=True)
        res = model(image, noise=noise)
        if fused:
            loss, log = model.loss(res, image, label)
        else:
            orig = SCAE._fused_loss_head
            SCAE._fused_loss_head = lambda self, res, label: None
            try:
                loss, log = model.loss(res, image, label)
            finally:
                SCAE._fused_loss_head = orig
        loss.backward()
        acc = model.calculate_accuracy(res, label)
        return loss, log, res, acc, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    from torch_scae_b200 import ops
    with ops.KernelTimer() as timer:
        loss_f, log_f, res_f, acc_f, grads_f = run(True)
    torch.cuda.synchronize()
    assert 'scae_loss_head_fwd' in timer.summary() and 'scae_loss_head_bwd' in timer.summary()
    loss_e, log_e, res_e, acc_e, grads_e = run(False)
    assert rel_err(loss_f, loss_e) < 1e-6
    assert set(log_f) == set(log_e)
    for k in log_e:
        assert rel_err(log_f[k], log_e[k]) < 1e-5, k
    assert rel_err(res_f.prior_cls_prob, res_e.prior_cls_prob) < 1e-5
    assert rel_err(res_f.posterior_cls_prob, res_e.posterior_cls_prob) < 1e-5
    assert float(acc_f) == float(acc_e)
    assert set(grads_f) == set(grads_e)
    for k in grads_e:
        assert rel_err(grads_f[k], grads_e[k]) < 1e-4, k


@pytest.mark.parametrize('B,Cin,G_,n,D', [(37, 128, 5, 40, 23), (8, 16, 3, 24, 23), (5, 8, 7, 3, 8), (1024, 128, 5, 40, 23)])
def test_attention_conv_pool_matches_conv_then_pooling(B, Cin, G_, n, D):
    """ops.attention_conv_pool (GEMM over positions + csrc/attnpool_cl.cu + bias after pooling) vs the reference
    formulation conv2d -> multiple_attention_pooling_2d (part_encoder.py:95-101, nn_ext.py:76-101) in fp64."""
    import torch.nn.functional as F
    from torch_scae_b200 import nn_ext, ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(B + n)
    conv = torch.nn.Conv2d(Cin, n * (D + 1), 1)
    x = torch.randn(B, Cin, G_, G_, generator=g)
    up = torch.randn(B, n * D, 1, 1, generator=g)
    x64 = x.double().requires_grad_(True)
    w64, b64 = conv.weight.detach().double().requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
    ref = nn_ext.multiple_attention_pooling_2d(F.conv2d(x64, w64, b64), n)
    g_ref = torch.autograd.grad((ref * up.double()).sum(), [x64, w64, b64])
    conv = conv.cuda()
    xd = x.cuda().requires_grad_(True)
    got = ops.attention_conv_pool(xd, conv, n)
    assert got is not None and got.shape == ref.shape
    assert rel_err(got, ref) < 1e-5
    g_got = torch.autograd.grad((got * up.cuda()).sum(), [xd, conv.weight, conv.bias])
    for name, a, r in zip(('x', 'weight', 'bias'), g_got, g_ref):
        assert rel_err(a, r) < 1e-4, name


def test_part_encoder_head_paths_agree():
    """CapsuleImageEncoder with the GEMM head (default) vs the cuDNN convolution + NCHW pooling (SCAE_B200_ATT_GEMM=0)."""
    import os
    from torch_scae_b200.part_encoder import CapsuleImageEncoder, CNNEncoder
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    enc = CapsuleImageEncoder((1, 40, 40), CNNEncoder((1, 40, 40), [32] * 4, [3] * 4, [2, 2, 1, 1]), n_caps=40,
                              n_poses=6, n_special_features=16).cuda().train()
    image = torch.rand(9, 1, 40, 40, device='cuda')
    noise = (torch.rand(9, 40, device='cuda') - .5) * 4

    def run():
        enc.zero_grad(set_to_none=True)
        r = enc(image, presence_noise=noise)
        (r.pose.sum() * 0.3 + (r.presence * r.presence).sum() + (r.feature ** 2).sum()).backward()
        return r, {k: p.grad.clone() for k, p in enc.named_parameters()}
    from torch_scae_b200 import ops
    with ops.KernelTimer() as timer:
        r1, g1 = run()
    torch.cuda.synchronize()
    assert 'scae_attnpool_cl_fwd' in timer.summary() and 'scae_attnpool_cl_bwd' in timer.summary()
    os.environ['SCAE_B200_ATT_GEMM'] = '0'
    try:
        r0, g0 = run()
    finally:
        del os.environ['SCAE_B200_ATT_GEMM']
    for k in ('pose', 'presence', 'feature'):
        assert rel_err(r1[k], r0[k]) < 1e-5, k
    for k in g0:
        assert rel_err(g1[k], g0[k]) < 1e-4, k


def test_conv_gemm_paths_are_taken():
    """The shipped rule: 128 -> 128 stride-2 layer fully in GEMM form, stride-1 layers with the GEMM data gradient, the
    1-channel input layer with cuDNN (guards against a silent fall-back to the slower formulation)."""
    from torch_scae_b200 import ops
    for cin, hw, stride, expect in ((128, 19, 2, {'scae_im2col3x3', 'scae_col2im3x3'}), (128, 9, 1, {'scae_col2im3x3'}),
                                    (1, 40, 2, set())):
        conv = torch.nn.Conv2d(cin, 128, 3, stride).cuda()
        x = torch.randn(8, cin, hw, hw, device='cuda', requires_grad=True)
        with ops.KernelTimer() as timer:
            ops.conv_bias_act(x, conv, True).sum().backward()
        torch.cuda.synchronize()
        got = {k for k in timer.summary() if k in ('scae_im2col3x3', 'scae_col2im3x3')}
        assert got == expect, (cin, hw, stride, got)


@pytest.mark.parametrize('B,C,H,W', [(64, 128, 9, 9), (3, 5, 3, 11), (1024, 128, 5, 5)])
def test_nchw_rows_transposes(B, C, H, W):
    """scae_transpose_batched behind ops._NchwToRows (the GEMM operands' layout): exact data movement both ways."""
    from torch_scae_b200 import ops
    x = torch.randn(B, C, H, W, device='cuda', requires_grad=True)
    rows = ops._NchwToRows.apply(x)
    assert torch.equal(rows, x.permute(0, 2, 3, 1).reshape(B * H * W, C))
    up = torch.randn(B * H * W, C, device='cuda')
    (gx,) = torch.autograd.grad((rows * up).sum(), [x])
    assert torch.equal(gx, up.view(B, H, W, C).permute(0, 3, 1, 2))


def test_loss_head_split_forward_equals_the_fused_one():
    """scae_loss_head_fwd_rows + scae_loss_head_fwd_finish (the halves a data-parallel caller all-reduces between) give the
    terms and statistics of scae_loss_head_fwd bit for bit on one rank."""
    import ctypes
    from torch_scae_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(3)
    B, O, V, K = 37, 32, 40, 10
    cp, post = torch.rand(B, O, device=DEV), torch.rand(B, O, V, device=DEV)
    label = torch.randint(0, K, (B,), device=DEV)
    w, b = torch.randn(K, O, device=DEV), torch.randn(K, device=DEV)
    args = _lib.LossHeadArgs(_lib.ptr(cp), _lib.ptr(post), _lib.ptr(label), _lib.ptr(w), _lib.ptr(b), B, O, V, K,
                             1, _lib.LOSS_TYPES['l2'], _lib.LOSS_TYPES['entropy'], 0.7, 0.9, 1.3, 0.4, 3.2, 3.2, B / K)
    ws_bytes = lib.scae_loss_head_workspace_bytes(ctypes.byref(args))
    ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)
    terms = [torch.zeros(8, device=DEV) for _ in range(2)]
    stats = [torch.zeros(128, device=DEV) for _ in range(2)]
    probs = [torch.zeros(2, B, K, device=DEV) for _ in range(2)]
    _lib.check(lib.scae_loss_head_fwd(ctypes.byref(args), _lib.ptr(terms[0]), _lib.ptr(probs[0]), _lib.ptr(stats[0]),
                                      _lib.ptr(ws), ws_bytes, ops._stream()), 'fwd')
    colsums = torch.zeros(132, device=DEV)
    _lib.check(lib.scae_loss_head_fwd_rows(ctypes.byref(args), _lib.ptr(probs[1]), _lib.ptr(colsums), _lib.ptr(ws),
                                           ws_bytes, ops._stream()), 'rows')
    _lib.check(lib.scae_loss_head_fwd_finish(ctypes.byref(args), _lib.ptr(colsums), _lib.ptr(terms[1]),
                                             _lib.ptr(stats[1]), ops._stream()), 'finish')
    assert torch.equal(terms[0], terms[1]) and torch.equal(stats[0], stats[1]) and torch.equal(probs[0], probs[1])
    assert rel_err(colsums[:O], cp.sum(0)) < 1e-6 and rel_err(colsums[64:64 + O], post.sum(-1).sum(0) / V) < 1e-6
