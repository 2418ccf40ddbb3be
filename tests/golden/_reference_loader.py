"""Import the UNMODIFIED reference (bdsaglam/torch-scae) from /root/reference to generate golden vectors.

This only works in the build container (``/root/reference`` does not exist on the GPU box), and only
``tests/golden/make_golden.py`` uses it.  Two accommodations, both described in SURVEY.md section 8(c):

* ``monty`` is not installed; the reference only uses ``monty.collections.AttrDict`` (part_decoder.py:22,
  object_decoder.py:19, part_encoder.py:19), so a dict-with-attribute-access shim is put on ``sys.modules``.
* ``theta *= 2. * math.pi`` (cv_ops.py:45) is an in-place write into a ``torch.split`` view, which modern autograd
  rejects.  The function's source is loaded, that single statement is rewritten out-of-place (forward values are
  bit-identical) and the module attribute is rebound; part_encoder.py:110 and object_decoder.py:239 look it up
  through the module so they pick it up.
"""
import importlib
import inspect
import math
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def load_reference():
    """Returns the imported ``torch_scae`` reference package (patched as described above)."""
    if "monty" not in sys.modules:
        monty = types.ModuleType("monty")
        collections = types.ModuleType("monty.collections")
        collections.AttrDict = _AttrDict
        monty.collections = collections
        sys.modules["monty"] = monty
        sys.modules["monty.collections"] = collections
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ref = importlib.import_module("torch_scae")
    cv_ops = importlib.import_module("torch_scae.cv_ops")
    if not getattr(cv_ops, "_b200_patched", False):
        src = inspect.getsource(cv_ops.geometric_transform)
        bad = "theta *= 2. * math.pi"
        assert bad in src, "reference cv_ops changed; re-check the patch"
        src = src.replace(bad, "theta = theta * (2. * math.pi)")
        scope = {"torch": cv_ops.torch, "math": math}
        exec(compile(src, cv_ops.__file__, "exec"), scope)
        cv_ops.geometric_transform = scope["geometric_transform"]
        cv_ops._b200_patched = True
    for name in ("part_decoder", "object_decoder", "part_encoder", "set_transformer", "distributions",
                 "stacked_capsule_auto_encoder", "factory", "math_ops", "nn_ext"):
        importlib.import_module("torch_scae." + name)
    return ref
