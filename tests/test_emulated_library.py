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
        assert t.is_contiguous() and not t.is_cuda
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


@pytest.mark.parametrize('case', ['enc', 'soft', 'hard'])
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
