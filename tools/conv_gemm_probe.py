"""Per-layer timing of the part encoder's 3x3 convolutions (strict fp32): cuDNN forward / dgrad / wgrad one by one,
against the cuBLAS SGEMMs of an im2col formulation on the same shapes (rows = B * output positions, K = 9 * C_in).
Decides whether a GEMM-form backward is worth building.   python tools/conv_gemm_probe.py > gpurun_out/conv_gemm.txt"""
import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = 'cuda'
B = 1024


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for name, cin, hin, stride in (('L1', 1, 40, 2), ('L2', 128, 19, 2), ('L3', 128, 9, 1), ('L4', 128, 7, 1)):
    cout = 128
    hout = (hin - 3) // stride + 1
    x = torch.randn(B, cin, hin, hin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev)
    g = torch.randn(B, cout, hout, hout, device=dev)
    gflop = 2.0 * B * hout * hout * cout * cin * 9 / 1e9
    t_f = timeit(lambda: torch.nn.functional.conv2d(x, w, None, stride))
    bw = lambda mask: torch.ops.aten.convolution_backward(g, x, w, None, (stride, stride), (0, 0), (1, 1), False, (0, 0),
                                                          1, mask)
    t_d = timeit(lambda: bw([True, False, False])) if cin > 1 else float('nan')
    t_w = timeit(lambda: bw([False, True, False]))
    rows, K = B * hout * hout, 9 * cin
    cols = torch.randn(rows, K, device=dev)
    w2 = torch.randn(cout, K, device=dev)
    g2 = torch.randn(rows, cout, device=dev)
    t_gf = timeit(lambda: cols @ w2.t())
    t_gd = timeit(lambda: g2 @ w2)
    s = max(d for d in range(1, 33) if rows % d == 0)
    t_gw = timeit(lambda: torch.bmm(g2.view(s, rows // s, cout).transpose(1, 2), cols.view(s, rows // s, K)).sum(0))
    t_gw1 = timeit(lambda: g2.t() @ cols)
    t_perm = timeit(lambda: g.permute(0, 2, 3, 1).reshape(rows, cout))
    dc = torch.randn(B, K, hout * hout, device=dev)
    t_fold = timeit(lambda: torch.nn.functional.fold(dc, (hin, hin), 3, stride=stride))
    t_unfold = timeit(lambda: torch.nn.functional.unfold(x, 3, stride=stride))
    tf = lambda t: gflop / t
    print(f'{name} {cin}->{cout} {hin}x{hin} s{stride} -> {hout}x{hout}  {gflop:.1f} GFLOP per pass, rows {rows}, K {K}')
    print(f'  cuDNN  fwd {t_f:.3f} ms ({tf(t_f):.1f} TF/s)  dgrad {t_d:.3f} ms ({tf(t_d):.1f})  wgrad {t_w:.3f} ms ({tf(t_w):.1f})')
    print(f'  cuBLAS fwd {t_gf:.3f} ms ({tf(t_gf):.1f} TF/s)  dgrad {t_gd:.3f} ms ({tf(t_gd):.1f})  '
          f'wgrad split-{s} {t_gw:.3f} ms ({tf(t_gw):.1f}) / plain {t_gw1:.3f} ms ({tf(t_gw1):.1f})')
    print(f'  layout passes: NCHW->rows permute of g {t_perm:.3f} ms, F.fold {t_fold:.3f} ms, F.unfold {t_unfold:.3f} ms, '
          f'cols {rows * K * 4 / 1e6:.0f} MB', flush=True)
