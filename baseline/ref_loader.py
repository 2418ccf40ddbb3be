"""Imports the UNMODIFIED reference (bdsaglam/torch-scae) from ``baseline/_ref`` -- where ``baseline/install_ref.sh`` pip-installs
it from /root/reference (git-ignored, travels to the GPU box with the snapshot) -- for the reference arm of bench.py.
TEST / BENCH INFRASTRUCTURE: nothing under torch_scae_b200/ imports this.

Two run-time accommodations, the same as tests/golden/_reference_loader.py; the installed files stay
byte-identical to the reference:
  * ``monty`` is not installed and the reference only uses ``monty.collections.AttrDict``: a dict-with-attribute-access
    shim is put on ``sys.modules``;
  * ``theta *= 2. * math.pi`` (cv_ops.py:45) writes in place into a ``torch.split`` view, which autograd has rejected since
    torch 1.7: the function's source is loaded, that one statement is made out-of-place (bit-identical forward values) and
    the module attribute is rebound.
"""
import importlib
import inspect
import math
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


class _AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def available():
    return os.path.isfile(os.path.join(REF_DIR, 'torch_scae', 'stacked_capsule_auto_encoder.py'))


def load_reference():
    """-> the imported reference package ``torch_scae``; raises ImportError when baseline/_ref is absent."""
    if not available():
        raise ImportError(f'{REF_DIR} does not hold the reference (run baseline/install_ref.sh where /root/reference exists)')
    if 'monty' not in sys.modules:
        monty = types.ModuleType('monty')
        collections = types.ModuleType('monty.collections')
        collections.AttrDict = _AttrDict
        monty.collections = collections
        sys.modules['monty'] = monty
        sys.modules['monty.collections'] = collections
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    ref = importlib.import_module('torch_scae')
    cv_ops = importlib.import_module('torch_scae.cv_ops')
    if not getattr(cv_ops, '_b200_patched', False):
        src = inspect.getsource(cv_ops.geometric_transform)
        bad = 'theta *= 2. * math.pi'
        assert bad in src, 'reference cv_ops changed; re-check the patch'
        scope = {'torch': cv_ops.torch, 'math': math}
        exec(compile(src.replace(bad, 'theta = theta * (2. * math.pi)'), cv_ops.__file__, 'exec'), scope)
        cv_ops.geometric_transform = scope['geometric_transform']
        cv_ops._b200_patched = True
    for name in ('part_decoder', 'object_decoder', 'part_encoder', 'set_transformer', 'distributions',
                 'stacked_capsule_auto_encoder', 'factory'):
        importlib.import_module('torch_scae.' + name)
    return ref
