"""Builds torch_scae_b200/csrc/*.cu -- HOST CODE INCLUDED: validation, planning, launch sequences, the C ABI -- as a host
shared library that runs the kernels on the CPU under tests/emu/simt.h (TEST INFRASTRUCTURE).

Each .cu file is rewritten textually in two places only, then compiled with g++:
  * ``kernel<<<grid, block, smem, stream>>>(args);``  ->  ``emu_launch(dim3(grid), dim3(block), [&] { kernel(args); });``
  * ``extern __shared__ ... T name[];``              ->  ``T* name = reinterpret_cast<T*>(emu_dynamic_smem);``
Everything else comes from headers: tests/emu/simt.h (threads, barriers, warp intrinsics), tests/emu/ptx_emu.h (stand-ins
for csrc/ptx_sm100.cuh with deferred asynchronous copies), tests/emu/stubs/cuda_runtime.h (an imaginary device).  The
result exports the same ``scae_*`` symbols as libscae_b200.so and is called through ctypes with host pointers.
"""
import concurrent.futures
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU = os.path.join(ROOT, 'tests', 'emu')
CSRC = os.path.join(ROOT, 'torch_scae_b200', 'csrc')


def _matching(text, start, open_ch, close_ch):
    """index just past the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError('unbalanced')


def _split_top_level(s):
    parts, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':                # (angle brackets are not counted: `a->B`; no template argument lists with commas here)
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur)
            cur = ''
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(src):
    out, at = '', 0
    while True:
        i = src.find('<<<', at)
        if i < 0:
            return out + src[at:]
        # the kernel expression: an identifier, optionally with template arguments, right before <<<
        j = i
        if src[j - 1] == '>':                                  # template arguments
            depth, j = 0, j - 1
            while True:
                if src[j] == '>':
                    depth += 1
                elif src[j] == '<':
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in '_:'):
            j -= 1
        kernel = src[j:i]
        k = src.index('>>>', i)
        while src[k + 3] != '(':                               # '>>>' inside the launch configuration (none today)
            k = src.index('>>>', k + 1)
        cfg = _split_top_level(src[i + 3:k])
        end_args = _matching(src, k + 3, '(', ')')
        args = src[k + 4:end_args - 1]
        assert src[end_args] == ';', src[end_args - 40:end_args + 5]
        smem = cfg[2] if len(cfg) > 2 else '0'
        out += src[at:j] + f'emu_launch(dim3({cfg[0]}), dim3({cfg[1]}), [&] {{ {kernel}({args}); }}, (size_t)({smem}));'
        at = end_args + 1


def rewrite(src):
    src = rewrite_launches(src)
    return re.sub(r'extern __shared__ (?:__align__\(\d+\) )?(float|unsigned char) (\w+)\[\];',
                  r'\1* \2 = reinterpret_cast<\1*>(emu_dynamic_smem);', src)


def build(build_dir, sm_count=2, extra_flags=(), link_flags=()):
    """-> path of the emulated shared library.  ``extra_flags`` / ``link_flags``: e.g. ('-g', '-fsanitize=address',
    '-DEMU_EXACT_SMEM') and ('-fsanitize=address',) for the memcheck build, ('-g', '-fsanitize=thread') for the race
    check (tests/emu/README.md)."""
    from torch_scae_b200.build import SOURCES
    os.makedirs(build_dir, exist_ok=True)
    flags = ['-x', 'c++', '-std=c++20', '-O1', '-fPIC', '-pthread', '-ffp-contract=off', '-D_GNU_SOURCE',
             f'-DEMU_SM_COUNT={sm_count}', '-I', os.path.join(EMU, 'stubs'), '-I', EMU, '-I', CSRC,
             '-I', os.path.join(ROOT, 'include'), '-include', 'simt.h', '-include', 'ptx_emu.h', '-w', *extra_flags]

    def compile_one(name):
        text = rewrite(open(os.path.join(CSRC, name)).read())
        # quoted includes resolve relative to the including file: keep the rewritten copy next to nothing, add -I CSRC
        path = os.path.join(build_dir, name.replace('.cu', '.emu.cpp'))
        open(path, 'w').write(text)
        obj = path.replace('.cpp', '.o')
        r = subprocess.run(['g++', *flags, '-c', path, '-o', obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'{name}:\n{r.stderr[:4000]}')
        return obj
    with concurrent.futures.ThreadPoolExecutor(8) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    lib = os.path.join(build_dir, 'libscae_b200_emu.so')
    subprocess.run(['g++', '-shared', '-pthread', *link_flags, '-o', lib, *objs], check=True)
    return lib


if __name__ == '__main__':
    import sys
    sys.path.insert(0, ROOT)
    mode = sys.argv[2] if len(sys.argv) > 2 else ''
    flags = {'asan': (('-g', '-fsanitize=address', '-fno-omit-frame-pointer', '-DEMU_EXACT_SMEM'), ('-fsanitize=address',)),
             'tsan': (('-g', '-fsanitize=thread'), ('-fsanitize=thread',)), '': ((), ())}[mode]
    print(build(sys.argv[1] if len(sys.argv) > 1 else '/tmp/scae_emu_build', extra_flags=flags[0], link_flags=flags[1]))
