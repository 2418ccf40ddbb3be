"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/scae_b200.h declares, and the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from torch_scae_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'scae_b200.h')).read()
    declared = set(re.findall(r'\b(scae_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.scae_abi_version() == _lib.ABI_VERSION
    assert lib.scae_build_arch() == b'sm_100a'


def test_ctypes_structs_match_header_field_order():
    from torch_scae_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'scae_b200.h')).read()

    def fields(struct):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (struct, struct), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                names.append(re.findall(r'[A-Za-z_][A-Za-z0-9_]*', part)[-1])
        return names
    for struct, cls in (('scae_tmpl_args', _lib.TmplArgs), ('scae_caps_args', _lib.CapsArgs),
                        ('scae_caps_outputs', _lib.CapsOutputs), ('scae_caps_upstream', _lib.CapsUpstream),
                        ('scae_caps_saved', _lib.CapsSaved), ('scae_loss_head_args', _lib.LossHeadArgs)):
        assert fields(struct) == [f[0] for f in cls._fields_], struct


def test_invalid_arguments_are_reported_not_crashed():
    import ctypes
    from torch_scae_b200 import _lib
    lib = _lib.load()
    a = _lib.TmplArgs()          # all NULL / zero
    rc = lib.scae_tmpl_ll_fwd(ctypes.byref(a), None, None, None, None, None)
    assert rc == -1 and b'positive' in lib.scae_last_error()
    c = _lib.CapsArgs()
    assert lib.scae_caps_ll_fwd(ctypes.byref(c), ctypes.byref(_lib.CapsOutputs()), None) == -1
    assert lib.scae_caps_ll_bwd_workspace_bytes(ctypes.byref(c)) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-CUDA behaviour')
def test_no_cpu_fallback():
    from torch_scae_b200 import _lib, factory
    from golden.cases import scae_case_params
    model = factory.make_scae(scae_case_params('enc'))
    with pytest.raises(_lib.ScaeError):
        model(torch.rand(2, 1, 20, 20))
