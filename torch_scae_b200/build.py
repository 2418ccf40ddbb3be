"""Builds libscae_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Used by ``__graft_entry__.build()`` and ``python -m torch_scae_b200.build``.  The library has no PyTorch or Python
dependency: plain CUDA runtime, ``extern "C"`` entry points declared in include/scae_b200.h.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libscae_b200.so')
SOURCES = ['api.cu', 'support.cu', 'loss_head.cu', 'attnpool_cl.cu', 'conv_cols.cu', 'sab.cu', 'caps_ll.cu', 'caps_ll2.cu', 'caps_ll3.cu', 'caps_ll3_bwd.cu', 'caps_explicit.cu', 'tmpl_fwd.cu', 'tmpl_bwd.cu']
HEADERS = ['common.cuh', 'caps_common.cuh', 'ptx_sm100.cuh', 'tmpl_common.cuh', os.path.join('..', '..', 'include', 'scae_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build libscae_b200.so')


def source_id():
    """sha256 over every source and header the library is built from (and the flags): compiled into the library
    (``scae_build_id()``) and compared by ``_lib.load()``, so a stale .so -- e.g. one that travelled to another machine
    with newer sources -- is noticed whatever the file times say."""
    import hashlib
    hsh = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            hsh.update(f.encode() + b'\0' + fh.read())
    return hsh.hexdigest()[:16]


def built_id():
    """build id recorded in the existing library, or None"""
    if not os.path.exists(LIB_PATH):
        return None
    import ctypes
    try:
        lib = ctypes.CDLL(LIB_PATH)
        fn = lib.scae_build_id
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except (OSError, AttributeError):
        return None


def is_stale():
    return built_id() != source_id()


def build(force=False, verbose=False):
    """Compiles every .cu into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    build_id = source_id()

    def compile_one(src):
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, f'-DSCAE_BUILD_ID="{build_id}"', '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    # the translation units are independent: compile them side by side (verbose: one after the other, readable output)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=1 if verbose else min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIB_PATH, *objs, '-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
