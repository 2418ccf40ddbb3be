"""Sanitizer passes over the kernels under the emulation (TEST INFRASTRUCTURE, run by hand).

Race check -- every emulated CUDA thread is a real thread, so ThreadSanitizer reports a shared-memory access that no
barrier orders (a missing __syncthreads / __syncwarp) as a data race:

    python tests/emu/build_lib.py /tmp/scae_emu_tsan tsan
    LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 history_size=4" \\
        python tests/emu/sanitizer_run.py /tmp/scae_emu_tsan/libscae_b200_emu.so 2> tsan.txt

Memory check -- with -DEMU_EXACT_SMEM every launch gets its dynamic shared memory as a heap block of exactly the size
it asked for, so AddressSanitizer catches overruns of a kernel's shared-memory layout as well as out-of-bounds global
accesses (the CPU counterpart of compute-sanitizer memcheck):

    python tests/emu/build_lib.py /tmp/scae_emu_asan asan
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 \\
        python tests/emu/sanitizer_run.py /tmp/scae_emu_asan/libscae_b200_emu.so 2> asan.txt

Round 2: the rewritten path-1 backward (run-walk kernel with the per-warp cell queue, tmpl_bwd.cu) is clean under both
passes at the MNIST, stress and colour shapes (whole-image and banded pixel records) and the C=3 temperature case.

Round 1 result (the shapes below and both hot paths again at the MNIST / stress shapes with their real launch
configurations): no memory errors; no races in the template, capsule (general and TMA-staged, with the deferred
asynchronous copies), loss-head, pooling, im2col / col2im, transpose and LayerNorm kernels.  The set-attention kernels
report races for ragged N only: threads beyond 4 N are clamped to the last token, recompute it redundantly and read its
rows across warps behind a __syncwarp; their results are discarded (no stores, no contribution to the parameter
gradients), so the race is benign.
"""
import sys, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from torch_scae_b200 import _lib, ops, set_transformer
import gpu_util
lib = ctypes.CDLL(sys.argv[1])
for name,(r,a) in _lib.SYMBOLS.items():
    fn=getattr(lib,name); fn.restype, fn.argtypes = r, a
hp = lambda t: None if t is None else t.data_ptr()
_lib._lib = lib; _lib.ptr = hp; ops.ptr = hp; ops._stream = lambda: None
ops._f32c = lambda t: None if t is None else t.float().contiguous()
gpu_util.DEV='cpu'
def f32(d):
    r = lambda t: t.float().double() if isinstance(t, torch.Tensor) else ([x.float().double() for x in t] if isinstance(t, list) else ({k: v.float().double() for k, v in t.items()} if isinstance(t, dict) else t))
    return {k: r(v) for k, v in d.items()}
print('template C=3 temperature', flush=True)
d = f32(gpu_util.make_template_inputs(2,4,3,7,9,12,10, alpha=False, presence=True, bg_image=True, learn_scale=True, seed=1))
gpu_util.template_cuda(d)
print('template MNIST-ish', flush=True)
d = f32(gpu_util.make_template_inputs(2,12,1,11,11,40,40, alpha=True, seed=2))
gpu_util.template_cuda(d)
print('template, several template groups of one image per CTA (records staged once, no barrier between the groups)', flush=True)
d = f32(gpu_util.make_template_inputs(2,40,1,7,7,20,20, alpha=True, seed=2))
gpu_util.template_cuda(d)
print('template stress shape (banded pixel records)', flush=True)
d = f32(gpu_util.make_template_inputs(1,9,1,21,21,64,64, alpha=True, seed=3))
gpu_util.template_cuda(d)
print('template colour shape (banded pixel records, 128-register variant)', flush=True)
d = f32(gpu_util.make_template_inputs(1,9,3,11,11,32,32, alpha=True, seed=4))
gpu_util.template_cuda(d)
flags = dict(similarity=False, learn_vote_scale=True, allow_deformations=True)
print('capsule general', flush=True)
d = f32(gpu_util.make_capsule_inputs(3,4,5, seed=1)); gpu_util.capsule_cuda(d, flags)
print('capsule fast', flush=True)
d = f32(gpu_util.make_capsule_inputs(4,10,12, seed=1)); gpu_util.capsule_cuda(d, flags, which=('ll_per_example','reg_per_example','posterior_mixing_prob','caps_presence'), part_grads=False)
print('loss head', flush=True)
B,O,V,K=70,10,8,10
cp=torch.rand(B,O,requires_grad=True); post=(torch.rand(B,O,V)/O).requires_grad_(True); label=torch.randint(0,K,(B,)); lin=torch.nn.Linear(O,K)
cfg=(1,0,1,2.0,.35,.7,.2,1.0,1.0,7.0)
t,_,_=ops._LossHead.apply(cp,post,label,lin.weight,lin.bias,cfg); t.backward()
print('attnpool_cl + conv cols + transposes', flush=True)
y=torch.randn(5,9,6*8,requires_grad=True); ops._AttentionPoolCL.apply(y,6,7).sum().backward()
conv=torch.nn.Conv2d(40,33,3,2); x=torch.randn(3,40,9,8,requires_grad=True); ops._Conv3x3Gemm.apply(x,conv.weight,conv.bias,2,True,True).sum().backward()
print('sab + layernorm', flush=True)
mab=set_transformer.MAB(d=16,n_heads=1,layer_norm=True); att=mab.mqkv
params=(att.q_projector.weight,att.q_projector.bias,att.k_projector.weight,att.k_projector.bias,att.v_projector.weight,att.v_projector.bias,att.o_projector.weight,att.o_projector.bias,mab.fc.weight,mab.fc.bias,mab.ln0.weight,mab.ln0.bias,mab.ln1.weight,mab.ln1.bias)
x=torch.randn(5,11,16,requires_grad=True); ops._SetAttentionBlock.apply(x,torch.rand(5,11),1e-5,1e-5,*params).sum().backward()
x=torch.randn(70,16,requires_grad=True); ops._LayerNorm.apply(x,torch.ones(16,requires_grad=True),torch.zeros(16,requires_grad=True),1e-5).sum().backward()
print('all done', flush=True)
