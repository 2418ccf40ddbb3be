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
