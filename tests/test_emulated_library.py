"""The whole C-ABI library on the CPU: torch_scae_b200/csrc/*.cu -- kernels AND host code (validation, planning, launch
sequences) -- compiled for the host by tests/emu/build_lib.py and driven through the product's own ctypes binding and
autograd Functions (torch_scae_b200/ops.py), with tensors in host memory.  Compared with the fp64 oracle exactly like
the GPU parity tests (tests/test_gpu_template.py, tests/test_gpu_capsule.py) do, at shapes the emulation finishes in
seconds.

This is test infrastructure: the binding is redirected by monkeypatching INSIDE this test module only.  The product has
no such switch -- torch_scae_b200 still refuses CPU tensors and a missing CUDA library (tests/test_lib_abi.py)."""
import ctypes

import pytest
import torch

import gpu_util
from conftest import l2_rel_err, rel_err


@pytest.fixture(scope='module')
def emulated_library(tmp_path_factory):
    from emu.build_lib import build
    return build(str(tmp_path_factory.mktemp('scae_emu_lib')))


@pytest.fixture
def emu(emulated_library, monkeypatch):
    """Redirects the ctypes binding to the emulated library and lets host tensors through the pointer helpers."""
    from torch_scae_b200 import _lib, ops
    lib = ctypes.CDLL(emulated_library)
    for name, (restype, argtypes) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    assert lib.scae_abi_version() == _lib.ABI_VERSION

    def host_ptr(t):
        if t is None:
            return None
        assert t.is_contiguous() and t.device.type == 'cpu'
        return t.data_ptr()
    monkeypatch.setattr(_lib, '_lib', lib)
    monkeypatch.setattr(_lib, 'ptr', host_ptr)
    monkeypatch.setattr(ops, 'ptr', host_ptr)
    monkeypatch.setattr(ops, '_stream', lambda: None)
    monkeypatch.setattr(ops, '_f32c', lambda t: None if t is None else t.float().contiguous())
    monkeypatch.setattr(gpu_util, 'DEV', 'cpu')
    return lib


def f32_inputs(d):
    """fp32-representable inputs, upcast: the oracle (fp64) and the kernels (fp32) then see identical values."""
    def r(t):
        if isinstance(t, torch.Tensor):
            return t.float().double()
        if isinstance(t, list):
            return [x.float().double() for x in t]
        if isinstance(t, dict):
            return {k: v.float().double() for k, v in t.items()}
        return t
    return {k: r(v) for k, v in d.items()}


TEMPLATE_CASES = [
    # B, M, C, h, w, H, W, alpha mode, presence, bg_image, learnt sigma
    (3, 5, 1, 5, 5, 12, 12, True, True, False, False),
    (2, 40, 1, 11, 11, 40, 40, True, True, False, False),      # the MNIST config of the train step
    (2, 4, 3, 7, 9, 12, 10, False, True, True, True),          # temperature mode, colour, non-square, bg image, sigma
    (2, 3, 2, 4, 4, 9, 7, True, False, False, True),
]


@pytest.mark.parametrize('B,M,C,h,w,H,W,alpha,presence,bg_image,learn_scale', TEMPLATE_CASES)
def test_emulated_template_path_matches_the_oracle(emu, B, M, C, h, w, H, W, alpha, presence, bg_image, learn_scale):
    """Hot path 1: scae_tmpl_ll_fwd / scae_tmpl_ll_bwd (csrc/tmpl_fwd.cu, tmpl_bwd.cu incl. their host-side planning)."""
    d = f32_inputs(gpu_util.make_template_inputs(B, M, C, h, w, H, W, alpha=alpha, presence=presence, bg_image=bg_image,
                                                 learn_scale=learn_scale, seed=B + M + C))
    got, ref = gpu_util.template_cuda(d), gpu_util.template_oracle(d)
    assert rel_err(got['log_prob'], ref['log_prob']) < 1e-5
    for k in ref:
        if not k.startswith('g_'):
            continue
        if k == 'g_pose':
            # bilinear cell decisions make pose gradients discontinuous: norm-wise
            assert l2_rel_err(got[k], ref[k]) < 2e-3, k
        else:
            assert rel_err(got[k], ref[k]) < 1e-4, k


def test_emulated_explicit_vote_likelihood_vs_reference_golden(emu):
    """scae_caps_explicit_fwd / _bwd (csrc/caps_explicit.cu): the standalone CapsuleLikelihood on explicit votes, every
    output and every input gradient against the values recorded from the reference (object_decoder.py:243-372)"""
    from conftest import load_golden, sub
    from torch_scae_b200 import ops
    g = load_golden('capsule_likelihood_explicit')
    leaf = {k: g[k].clone().requires_grad_(True) for k in ('vote', 'scale', 'vote_presence', 'dummy_vote', 'x', 'presence')}
    res = dict(zip(ops.EXPLICIT_RETURNS, ops.CapsuleExplicitLikelihood.apply(
        leaf['vote'], leaf['scale'], leaf['vote_presence'], leaf['dummy_vote'], leaf['x'], leaf['presence'])))
    res['log_prob'] = res.pop('ll_per_example').mean()
    out = sub(g, 'out.')
    assert set(out) == set(res)
    for k, ref in out.items():
        if ref.dtype == torch.int64:
            assert torch.equal(res[k], ref), k
        else:
            assert rel_err(res[k], ref) < 1e-5, k
    loss = 1.3 * res['log_prob']
    for k, w in sub(g, 'weight.').items():
        loss = loss + 0.4 * (res[k] * w).sum()
    loss.backward()
    for k, t in leaf.items():
        assert rel_err(t.grad, g['g_' + k]) < 1e-4, k
    # presence=None is presence = ones; only the likelihood differentiated
    ones = torch.ones_like(leaf['presence'])
    a = ops.CapsuleExplicitLikelihood.apply(leaf['vote'], leaf['scale'], leaf['vote_presence'], leaf['dummy_vote'],
                                            leaf['x'], None)[0]
    b = ops.CapsuleExplicitLikelihood.apply(leaf['vote'], leaf['scale'], leaf['vote_presence'], leaf['dummy_vote'],
                                            leaf['x'], ones)[0]
    assert torch.equal(a, b)
    for t in leaf.values():
        t.grad = None
    a.sum().backward()
    ga = leaf['vote'].grad.clone()
    leaf['vote'].grad = None
    b.sum().backward()
    assert torch.equal(ga, leaf['vote'].grad)


@pytest.mark.parametrize('B,O,V,part_grads', [(3, 4, 5, True), (3, 4, 5, False), (2, 32, 40, False), (4, 10, 40, False)])
def test_emulated_capsule_path_matches_the_oracle(emu, B, O, V, part_grads):
    """Hot path 2 through scae_caps_ll_fwd / _bwd: with gradients for the part poses the general kernels (csrc/caps_ll.cu)
    run, without them (what training does) the TMA-staged fast path (csrc/caps_ll2.cu) -- counted by the library."""
    flags = dict(similarity=False, learn_vote_scale=True, allow_deformations=True)
    d = f32_inputs(gpu_util.make_capsule_inputs(B, O, V, seed=B + O))
    which = gpu_util.CAPS_UP if part_grads else ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob',
                                                 'caps_presence')
    before = emu.scae_caps_fast_path_count()
    got = gpu_util.capsule_cuda(d, flags, which=which, part_grads=part_grads)
    fast_calls = emu.scae_caps_fast_path_count() - before
    assert fast_calls == (1 if part_grads else 2)          # forward always qualifies; the backward only without g_x
    ref = gpu_util.capsule_oracle(d, flags, which=which)
    for k in ('vote', 'scale', 'vote_presence', 'caps_presence', 'll_per_example', 'reg_per_example', 'winner',
              'soft_winner', 'soft_winner_presence', 'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit'):
        assert rel_err(got[k], ref[k]) < 1e-5, k
    assert torch.equal(got['is_from_capsule'], ref['is_from_capsule'])
    for k in ('g_all_param', 'g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3') + (('g_x', 'g_presence') if part_grads else ()):
        assert rel_err(got[k], ref[k]) < 1e-4, k


PERSISTENT_CASES = [
    # B, O, V, development switches of the host-side planner (pairs per thread, ring depth)
    (5, 4, 5, {}), (5, 4, 5, {'SCAE_CAPS3_NP': '4', 'SCAE_CAPS3_BWD_NP': '4'}),
    (5, 4, 5, {'SCAE_CAPS3_NP': '1', 'SCAE_CAPS3_BWD_NP': '1'}),
    (7, 3, 6, {'SCAE_CAPS3_STAGES': '2', 'SCAE_CAPS3_BWD_STAGES': '2'}),      # O * A odd: edge floats through registers
    (3, 10, 7, {}), (1, 1, 1, {}),
    (2, 32, 40, {}),                                                         # the MNIST shape with its launch configuration
]


@pytest.mark.parametrize('B,O,V,env', PERSISTENT_CASES)
@pytest.mark.parametrize('flags', [dict(similarity=False, learn_vote_scale=True, allow_deformations=True),
                                   dict(similarity=True, learn_vote_scale=False, allow_deformations=False)])
def test_emulated_persistent_capsule_kernels(emu, monkeypatch, B, O, V, env, flags):
    """csrc/caps_ll3.cu / caps_ll3_bwd.cu (persistent CTAs, TMA stage ring on mbarriers, named barriers, REDUX, cp.async,
    bulk-store gradient rows) executed on the CPU: every forward tensor and the training-set gradients vs the fp64 oracle,
    and both calls must have been served by these kernels."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    full = flags['similarity'] is False
    d = f32_inputs(gpu_util.make_capsule_inputs(B, O, V, presence=full, noise=full, seed=B + O + V))
    which = ('ll_per_example', 'reg_per_example', 'posterior_mixing_prob', 'caps_presence')
    before = emu.scae_caps_persistent_path_count()
    got = gpu_util.capsule_cuda(d, flags, which=which, part_grads=False)
    assert emu.scae_caps_persistent_path_count() - before == 2
    ref = gpu_util.capsule_oracle(d, flags, which=which)
    for k in ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote', 'caps_presence',
              'vote_presence_binary', 'winner', 'winner_presence', 'soft_winner', 'soft_winner_presence',
              'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit', 'll_per_example', 'reg_per_example'):
        assert rel_err(got[k].reshape(ref[k].shape), ref[k]) < 1e-5, k
    assert torch.equal(got['is_from_capsule'], ref['is_from_capsule'])
    for k in ('g_all_param', 'g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3'):
        assert rel_err(got[k].reshape(ref[k].shape), ref[k]) < 1e-4, k


@pytest.mark.parametrize('B,O,V', [(5, 4, 5), (3, 10, 7), (2, 32, 40)])
def test_emulated_persistent_capsule_backward_with_winner_gradients(emu, B, O, V):
    """The winner-gradient variant of the persistent backward: upstream gradients of the soft and hard winners, g_x,
    g_presence and the dummy vote's gradient (what vote_type / presence_type 'soft' or 'hard' and
    stop_grad_caps_target=False produce) stay on the persistent kernels."""
    flags = dict(similarity=False, learn_vote_scale=True, allow_deformations=True)
    d = f32_inputs(gpu_util.make_capsule_inputs(B, O, V, seed=B + O + V))
    which = tuple(k for k in gpu_util.CAPS_UP if k != 'mixing_log_prob')
    before = emu.scae_caps_persistent_path_count()
    got = gpu_util.capsule_cuda(d, flags, which=which, part_grads=True)
    assert emu.scae_caps_persistent_path_count() - before == 2
    ref = gpu_util.capsule_oracle(d, flags, which=which)
    for k in ('g_all_param', 'g_cpr_static', 'g_b0', 'g_b1', 'g_b2', 'g_b3', 'g_x', 'g_presence', 'g_dummy_vote'):
        assert rel_err(got[k].reshape(ref[k].shape), ref[k]) < 1e-4, k


def test_emulated_plumbing_kernels(emu):
    """A pass over the small kernels around the hot paths (csrc/support.cu, sab.cu, api.cu) against stock PyTorch ops."""
    from torch_scae_b200 import ops
    g = torch.Generator().manual_seed(0)
    # column sums
    x = torch.randn(300, 16, generator=g)
    assert rel_err(ops.colsum.__wrapped__(x) if hasattr(ops.colsum, '__wrapped__') else _colsum(ops, x), x.double().sum(0)) < 1e-5
    # LayerNorm(16)
    x = torch.randn(70, 16, generator=g, requires_grad=True)
    wt, bs = torch.randn(16, generator=g, requires_grad=True), torch.randn(16, generator=g, requires_grad=True)
    up = torch.randn(70, 16, generator=g)
    y = ops._LayerNorm.apply(x, wt, bs, 1e-5)
    got = torch.autograd.grad((y * up).sum(), [x, wt, bs])
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, wt, bs))
    y64 = torch.nn.functional.layer_norm(x64, (16,), w64, b64, 1e-5)
    ref = torch.autograd.grad((y64 * up.double()).sum(), [x64, w64, b64])
    assert rel_err(y, y64) < 1e-5
    for a, r in zip(got, ref):
        assert rel_err(a, r) < 1e-4
    # pose transform
    t = (torch.randn(33, 6, generator=g) * 0.7).requires_grad_(True)
    up = torch.randn(33, 6, generator=g)
    out = ops._PoseTransform.apply(t, False)
    (gt,) = torch.autograd.grad((out * up).sum(), [t])
    from torch_scae_b200 import cv_ops
    t64 = t.detach().double().requires_grad_(True)
    ref = cv_ops.geometric_transform(t64, False)
    (rt,) = torch.autograd.grad((ref * up.double()).sum(), [t64])
    assert rel_err(out, ref) < 1e-5 and rel_err(gt, rt) < 1e-5


def _colsum(ops, x):
    from torch_scae_b200 import _lib
    lib = _lib.load()
    rows, cols = x.shape
    ws_bytes = lib.scae_colsum_workspace_bytes(rows, cols)
    assert ws_bytes > 0
    out = torch.empty(cols)
    ws = torch.empty(ws_bytes, dtype=torch.uint8)
    _lib.check(lib.scae_colsum(x.data_ptr(), rows, cols, out.data_ptr(), ws.data_ptr(), ws_bytes, None), 'scae_colsum')
    return out


@pytest.mark.parametrize('case', ['enc', 'soft', 'hard', 'sparse', 'color_temp'])
def test_emulated_whole_model_vs_reference_golden(emu, case):
    """The whole SCAE (PyTorch modules on the CPU + both likelihood paths through the emulated library) against the
    golden vectors recorded from the reference itself (tests/golden/make_golden.py): loss, every log entry, outputs and
    every parameter gradient.  Same checks as tests/test_gpu_model.py::test_scae_vs_reference_golden."""
    from conftest import load_golden, sub
    from golden.cases import scae_case_params
    from torch_scae_b200 import factory
    g = load_golden('scae_' + case)
    model = factory.make_scae(scae_case_params(case))
    model.load_state_dict(sub(g, 'param.'), strict=True)
    model.train()
    image, label = g['image'], g['label']
    noise = dict(part_presence=g['noise_part_presence'], caps=g['noise_caps'], vote=g['noise_vote'])
    res = model(image, noise=noise)
    loss, log = model.loss(res, image, label)
    assert rel_err(loss, g['loss']) < 1e-5
    for k, ref in sub(g, 'log.').items():
        assert rel_err(log[k], ref) < 1e-5, k
    for k, ref in sub(g, 'out.').items():
        if k == 'rec_log_prob':
            got = res.rec.pdf.log_prob(image)
        elif k == 'rec_mixing_logits':
            with torch.no_grad():
                got = res.rec.mixing_logits
        else:
            got = res[k]
        if ref.dtype == torch.int64:
            assert torch.equal(got, ref), k
        else:
            assert rel_err(got, ref) < 2e-5, k
    assert float(model.calculate_accuracy(res, label)) == float(g['accuracy'])
    loss.backward()
    grads = {name: p.grad for name, p in model.named_parameters()}
    layer = model.obj_decoder.capsule_layer
    for mod_name, mod in (('mlps', layer.mlps), ('caps_mlps', layer.caps_mlps)):
        for suffix, p in mod._named():
            for i in range(mod.n):
                grads[f'obj_decoder.capsule_layer.{mod_name}.{i}.{suffix}'] = p.grad[i]
    for k, ref in sub(g, 'g_param.').items():
        got = grads[k]
        if float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        e = l2_rel_err(got, ref) if k.startswith('part_encoder.') else rel_err(got, ref)
        assert e < (2e-3 if k.startswith('part_encoder.') else 1e-4), (k, e)


def test_emulated_loss_head_attention_head_and_gemm_convolutions(emu):
    """The round's newer kernels through their C entry points (host-side grid / workspace logic included): loss head vs
    the analytic oracle, channels-last attention pooling vs nn_ext's formula, GEMM-form convolution vs F.conv2d."""
    import torch.nn.functional as F
    from oracle import manual_backward as mb
    from torch_scae_b200 import nn_ext, ops
    g = torch.Generator().manual_seed(5)
    # ---- loss head (csrc/loss_head.cu): 70 rows -> several rows per warp on the 2-SM emulated device
    B, O, V, K = 70, 10, 8, 10
    cp = torch.rand(B, O, generator=g).requires_grad_(True)
    post = (torch.rand(B, O, V, generator=g) / O).requires_grad_(True)
    label = torch.randint(0, K, (B,), generator=g)
    lin = torch.nn.Linear(O, K)
    ws = (2.0, 0.35, 0.7, 0.2)
    cfg = (1, 0, 1, *ws, float(O) / K, float(O) / K, float(B) / K)          # l2 prior, entropy posterior
    total, terms, probs = ops._LossHead.apply(cp, post, label, lin.weight, lin.bias, cfg)
    got = torch.autograd.grad(total * 1.3, [cp, post, lin.weight, lin.bias])
    ref = mb.loss_head_forward_backward(cp.detach().double(), post.detach().double(), label, lin.weight.detach().double(),
                                        lin.bias.detach().double(), K, 'l2', 'entropy', ws)
    assert rel_err(total, ref['total']) < 1e-5 and rel_err(probs[0], ref['prior_cls_prob']) < 1e-5
    for a, k in zip(got, ('g_caps_presence', 'g_posterior', 'g_weight', 'g_bias')):
        assert rel_err(a, 1.3 * ref[k]) < 1e-4, k
    # ---- channels-last attention pooling (csrc/attnpool_cl.cu)
    Bq, n, D, S = 5, 6, 7, 9
    y = torch.randn(Bq, S, n * (D + 1), generator=g, requires_grad=True)
    up = torch.randn(Bq * n, D, generator=g)
    out = ops._AttentionPoolCL.apply(y, n, D)
    (gy,) = torch.autograd.grad((out * up).sum(), [y])
    y64 = y.detach().double().requires_grad_(True)
    nchw = y64.view(Bq, 3, 3, n * (D + 1)).permute(0, 3, 1, 2)
    ref_out = nn_ext.multiple_attention_pooling_2d(nchw, n).reshape(Bq * n, D)
    (ref_gy,) = torch.autograd.grad((ref_out * up.double()).sum(), [y64])
    assert rel_err(out, ref_out) < 1e-5 and rel_err(gy, ref_gy) < 1e-5
    # ---- GEMM-form 3x3 convolution (csrc/conv_cols.cu + ops._Conv3x3Gemm), both variants, both strides
    # (the 31x31 / 15x15 maps are those of the 64x64 likelihood-stress config: 130 KB / 98 KB shared-memory tiles)
    for stride, full, cin, hw in ((2, True, 40, (9, 8)), (1, False, 40, (9, 8)), (1, True, 40, (9, 8)),
                                  (2, True, 32, (31, 31)), (1, False, 32, (15, 15))):
        assert emu.scae_conv_cols_supported(3, cin, hw[0], hw[1], stride) == 1
        conv = torch.nn.Conv2d(cin, 33, 3, stride)
        x = torch.randn(3, cin, *hw, generator=g, requires_grad=True)
        yv = ops._Conv3x3Gemm.apply(x, conv.weight, conv.bias, stride, True, full)
        upc = torch.randn(yv.shape, generator=g)
        got = torch.autograd.grad((yv * upc).sum(), [x, conv.weight, conv.bias])
        x64 = x.detach().double().requires_grad_(True)
        w64, b64 = conv.weight.detach().double().requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
        y64 = torch.relu(F.conv2d(x64, w64, b64, stride))
        ref = torch.autograd.grad((y64 * upc.double()).sum(), [x64, w64, b64])
        assert rel_err(yv, y64) < 1e-5
        for a, r in zip(got, ref):
            assert l2_rel_err(a, r) < 1e-4, (stride, full)


def test_emulated_set_attention_block_and_optimizer(emu):
    """csrc/sab.cu (one kernel per direction for a whole set-attention block) vs the module's PyTorch ops; the flat
    RMSprop kernel vs torch.optim.RMSprop."""
    from torch_scae_b200 import ops, set_transformer
    torch.manual_seed(3)
    import copy
    mab = set_transformer.MAB(d=16, n_heads=1, layer_norm=True)
    with torch.no_grad():
        for p in mab.parameters():                          # non-trivial LayerNorm affine parameters and biases
            p.add_(torch.randn_like(p) * 0.3)
    for presence in (torch.rand(5, 11), (torch.rand(5, 11) > 0.4).float(), None):
        x = torch.randn(5, 11, 16, requires_grad=True)
        att = mab.mqkv
        params = (att.q_projector.weight, att.q_projector.bias, att.k_projector.weight, att.k_projector.bias,
                  att.v_projector.weight, att.v_projector.bias, att.o_projector.weight, att.o_projector.bias,
                  mab.fc.weight, mab.fc.bias, mab.ln0.weight, mab.ln0.bias, mab.ln1.weight, mab.ln1.bias)
        up = torch.randn(5, 11, 16)
        y = ops._SetAttentionBlock.apply(x, presence, float(mab.ln0.eps), float(mab.ln1.eps), *params)
        got = torch.autograd.grad((y * up).sum(), [x, *params])
        ref_mod = copy.deepcopy(mab).double()                  # the module's PyTorch ops, in fp64
        x64 = x.detach().double().requires_grad_(True)
        ref_y = ref_mod(x64, x64, presence.double() if presence is not None else None)
        rm = ref_mod.mqkv
        ref_params = (rm.q_projector.weight, rm.q_projector.bias, rm.k_projector.weight, rm.k_projector.bias,
                      rm.v_projector.weight, rm.v_projector.bias, rm.o_projector.weight, rm.o_projector.bias,
                      ref_mod.fc.weight, ref_mod.fc.bias, ref_mod.ln0.weight, ref_mod.ln0.bias, ref_mod.ln1.weight,
                      ref_mod.ln1.bias)
        ref = torch.autograd.grad((ref_y * up.double()).sum(), [x64, *ref_params])
        assert rel_err(y, ref_y) < 2e-5
        # the key-bias gradient is identically zero (softmax is invariant to a per-row constant): errors are measured
        # against the gradient's own scale, floored at a fraction of the largest gradient of the block
        floor = 1e-2 * max(float(r.abs().max()) for r in ref)
        for a, r in zip(got, ref):
            assert float((a.double() - r).abs().max()) / max(float(r.abs().max()), floor) < 1e-4
    # RMSprop with momentum on a flat buffer
    from torch_scae_b200 import _lib
    lib = _lib.load()
    p = torch.randn(1000)
    grad = torch.randn(1000)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.RMSprop([ref_p], lr=3e-3, momentum=0.9, eps=1e-4)
    sq, buf = torch.zeros(1000), torch.zeros(1000)
    for _ in range(3):
        ref_p.grad = grad.clone()
        opt.step()
        _lib.check(lib.scae_rmsprop_step(p.data_ptr(), grad.data_ptr(), sq.data_ptr(), buf.data_ptr(), 1000, 3e-3, 0.99,
                                         1e-4, 0.9, None), 'scae_rmsprop_step')
    assert rel_err(p, ref_p) < 1e-6


BASELINE_SHAPES = {
    # BASELINE.json configs[4]: SVHN/CIFAR-shaped colour templates, 24 part capsules
    'color': dict(image_shape=(3, 32, 32), n_classes=10, n_part_caps=24, n_obj_caps=32,
                  scae_params=dict(reconstruct_alternatives=False)),
    # BASELINE.json configs[3]: likelihood-stress, 64x64 images, 21x21 templates, 64 part / 32 object capsules
    'stress': dict(image_shape=(1, 64, 64), n_classes=10, n_part_caps=64, n_obj_caps=32,
                   pcae_template_generator_params=dict(template_size=(21, 21)),
                   scae_params=dict(reconstruct_alternatives=False)),
}


@pytest.mark.parametrize('name', ['color', 'stress'])
def test_emulated_whole_model_at_the_other_baseline_shapes(emu, name):
    """One train step of the whole SCAE at the colour and likelihood-stress shapes of BASELINE.json (B = 2 / 1): loss, log
    entries and gradients against the CPU oracle model (oracle/scae_model.py, the reference's op sequence).  These
    shapes take the chunked template kernels (64 templates of 21x21 do not fit shared memory at once), three-channel
    texels, and the single-stage capsule backward."""
    import numpy as np
    from oracle import scae_model
    from torch_scae_b200 import factory
    params = BASELINE_SHAPES[name]
    torch.manual_seed(1)
    np.random.seed(1)
    model = factory.make_scae(params)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if 'templates_alpha' in pname or 'cpr_static' in pname or 'caps_bias_list' in pname:
                p.copy_(0.1 * torch.randn_like(p))
    cfg = factory.prepare_model_params(**params)
    B = 1 if name == 'stress' else 2          # (the 64x64 / 64-template step takes ~30 s per image under the emulation)
    C, H, W = params['image_shape']
    M, O = params['n_part_caps'], params['n_obj_caps']
    image = torch.rand(B, C, H, W)
    label = torch.randint(0, 10, (B,))
    noise = dict(part_presence=(torch.rand(B, M) - .5) * 4, caps=(torch.rand(B, O, 1) - .5) * 4,
                 vote=(torch.rand(B, O, M) - .5) * 4)
    # the oracle model in fp64 on the same fp32 weights and inputs: the arbiter (its own fp32 run deviates from this by up to
    # 2e-4 on the object encoder's gradients at some seeds, measured round 2)
    sd = {k: (v.detach().clone().double().requires_grad_(True) if v.is_floating_point() else v.clone())
          for k, v in model.state_dict().items()}
    ref_res = scae_model.scae_forward(sd, cfg, image.double(), {k: v.double() for k, v in noise.items()}, training=True)
    ref_loss, ref_log = scae_model.scae_loss(ref_res, cfg, image.double(), label)
    ref_loss.backward()

    model.train()
    res = model(image, noise=noise)
    loss, log = model.loss(res, image, label)
    loss.backward()
    assert rel_err(loss, ref_loss) < 1e-5
    for k in ref_log:
        assert rel_err(log[k], ref_log[k]) < 1e-5, k
    named = dict(model.named_parameters())
    for k in ('part_decoder.templates_alpha', 'part_decoder.bg_value', 'part_decoder.bg_mixing_logit',
              'template_generator.template_logits', 'obj_decoder.capsule_layer.cpr_static', 'obj_encoder.fc1.weight'):
        assert rel_err(named[k].grad, sd[k].grad) < 1e-4, k
    for k in ('part_encoder.att_conv.weight', 'part_encoder.encoder.network.0.weight'):
        assert l2_rel_err(named[k].grad, sd[k].grad) < 2e-3, k


@pytest.mark.parametrize('case', ['enc', 'hard', 'sparse', 'color_temp'])
def test_emulated_whole_model_through_every_fused_path(emu, monkeypatch, case):
    """As test_emulated_whole_model_vs_reference_golden, but with the modules' own routing switched to the kernels
    everywhere (tensors report ``is_cuda`` for the duration of the test): fused capsule head of the part encoder,
    GEMM-form convolutions, set-attention blocks, LayerNorm, split-K linears with the colsum kernel, pose transform, loss
    head -- the train step's whole kernel inventory on the CPU, against the reference's golden loss and gradients."""
    from conftest import load_golden, sub
    from golden.cases import scae_case_params
    from torch_scae_b200 import factory
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True), raising=False)
    g = load_golden('scae_' + case)
    model = factory.make_scae(scae_case_params(case))
    model.load_state_dict(sub(g, 'param.'), strict=True)
    model.train()
    image, label = g['image'], g['label']
    noise = dict(part_presence=g['noise_part_presence'], caps=g['noise_caps'], vote=g['noise_vote'])
    before = emu.scae_launch_count()
    res = model(image, noise=noise)
    loss, log = model.loss(res, image, label)
    loss.backward()
    launches = emu.scae_launch_count() - before
    assert launches >= 25, launches                       # (the two likelihood paths alone are 9 launches)
    assert rel_err(loss, g['loss']) < 1e-5
    for k, ref in sub(g, 'log.').items():
        assert rel_err(log[k], ref) < 1e-5, k
    assert float(model.calculate_accuracy(res, label)) == float(g['accuracy'])
    grads = {name: p.grad for name, p in model.named_parameters()}
    layer = model.obj_decoder.capsule_layer
    for mod_name, mod in (('mlps', layer.mlps), ('caps_mlps', layer.caps_mlps)):
        for suffix, p in mod._named():
            for i in range(mod.n):
                grads[f'obj_decoder.capsule_layer.{mod_name}.{i}.{suffix}'] = p.grad[i]
    for k, ref in sub(g, 'g_param.').items():
        got = grads[k]
        if float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        e = l2_rel_err(got, ref) if k.startswith('part_encoder.') else rel_err(got, ref)
        assert e < (2e-3 if k.startswith('part_encoder.') else 1e-4), (k, e)


@pytest.mark.parametrize('B,C,H,W', [(3, 128, 9, 9), (2, 5, 3, 11), (1, 33, 1, 1), (4, 40, 5, 5)])
def test_emulated_nchw_rows_transposes(emu, B, C, H, W):
    """scae_transpose_batched behind ops._NchwToRows: (B,C,H,W) -> (B*H*W, C) and back in the backward; pure data
    movement, so exact."""
    from torch_scae_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
    rows = ops._NchwToRows.apply(x)
    assert torch.equal(rows, x.permute(0, 2, 3, 1).reshape(B * H * W, C))
    up = torch.randn(B * H * W, C, generator=g)
    (gx,) = torch.autograd.grad((rows * up).sum(), [x])
    assert torch.equal(gx, up.view(B, H, W, C).permute(0, 3, 1, 2))
