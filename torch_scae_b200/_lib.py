"""ctypes binding of libscae_b200.so (C ABI declared in include/scae_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.  The library
is built in-tree by ``torch_scae_b200.build`` (``__graft_entry__.build()``); on import we only load it.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_long, c_size_t, c_uint, c_ulonglong, c_void_p

from .build import LIB_PATH

ABI_VERSION = 5

TMPL_MODE_ALPHA = 0
TMPL_MODE_TEMPERATURE = 1
CAPS_SIMILARITY = 1
CAPS_LEARN_VOTE_SCALE = 2
CAPS_ALLOW_DEFORM = 4
CAPS_RELU_GRAD = 8
LOSS_TYPES = {'l2': 0, 'entropy': 1, 'kl': 2}


class ScaeError(RuntimeError):
    pass


class TmplArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ('templates', 'templates_alpha', 'pose', 'presence', 'bg_image', 'bg_value',
                                        'bg_mixing_logit', 'temperature_logit', 'scale', 'template_color')] + \
               [(n, c_int) for n in ('B', 'M', 'C', 'h', 'w', 'H', 'W', 'mode')]


class SabParams(Structure):
    _fields_ = [(n, c_void_p) for n in ('wq', 'bq', 'wk', 'bk', 'wv', 'bv', 'wo', 'bo', 'wf', 'bf', 'ln0_w', 'ln0_b',
                                        'ln1_w', 'ln1_b')] + [('eps0', c_float), ('eps1', c_float)]


class LossHeadArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ('caps_presence', 'posterior', 'label', 'cls_weight', 'cls_bias')] + \
               [(n, c_int) for n in ('B', 'O', 'V', 'K', 'sparsity', 'prior_type', 'posterior_type')] + \
               [(n, c_float) for n in ('prior_within_weight', 'prior_between_weight', 'posterior_within_weight',
                                       'posterior_between_weight', 'prior_within_constant',
                                       'posterior_within_constant', 'between_constant')]


class CapsArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ('all_param', 'cpr_static', 'bias_cvr', 'bias_caps', 'bias_vote', 'bias_scale',
                                        'noise_caps', 'noise_vote', 'x', 'presence', 'dummy_vote')] + \
               [('B', c_int), ('O', c_int), ('V', c_int), ('flags', c_uint)]


class CapsExplicitArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ('vote', 'scale', 'vote_presence', 'dummy_vote', 'x', 'presence', 'point_ll')] + \
               [('B', c_int), ('O', c_int), ('V', c_int)]


CAPS_OUTPUT_FIELDS = ('vote', 'scale', 'vote_presence', 'presence_logit_per_caps', 'presence_logit_per_vote',
                      'caps_presence', 'caps_presence_arg', 'log_prob_per_point', 'll_per_example', 'reg_per_example',
                      'vote_presence_binary', 'winner', 'winner_presence', 'winner_idx', 'is_from_capsule',
                      'soft_winner', 'soft_winner_presence', 'posterior_mixing_prob', 'mixing_log_prob',
                      'mixing_logit')
CAPS_UPSTREAM_FIELDS = ('g_ll_per_example', 'g_reg_per_example', 'g_posterior_mixing_prob', 'g_caps_presence',
                        'g_vote_presence', 'g_soft_winner', 'g_soft_winner_presence', 'g_winner',
                        'g_winner_presence', 'g_vote', 'g_scale', 'g_presence_logit_per_caps',
                        'g_presence_logit_per_vote', 'g_mixing_logit', 'g_mixing_log_prob')
CAPS_SAVED_FIELDS = ('posterior_mixing_prob', 'log_prob_per_point', 'caps_presence_arg', 'winner_idx')


class CapsOutputs(Structure):
    _fields_ = [(n, c_void_p) for n in CAPS_OUTPUT_FIELDS]


class CapsUpstream(Structure):
    _fields_ = [(n, c_void_p) for n in CAPS_UPSTREAM_FIELDS]


class CapsSaved(Structure):
    _fields_ = [(n, c_void_p) for n in CAPS_SAVED_FIELDS]


# every symbol include/scae_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'scae_abi_version': (c_int, []),
    'scae_build_id': (c_char_p, []),
    'scae_last_error': (c_char_p, []),
    'scae_build_arch': (c_char_p, []),
    'scae_launch_count': (c_ulonglong, []),
    'scae_caps_fast_path_count': (c_ulonglong, []),
    'scae_caps_persistent_path_count': (c_ulonglong, []),
    'scae_tmpl_ll_fwd': (c_int, [POINTER(TmplArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'scae_tmpl_ll_bwd_workspace_bytes': (c_size_t, [POINTER(TmplArgs)]),
    'scae_tmpl_ll_bwd': (c_int, [POINTER(TmplArgs)] + [c_void_p] * 11 + [c_size_t, c_void_p]),
    'scae_tmpl_render': (c_int, [POINTER(TmplArgs)] + [c_void_p] * 6),
    'scae_tmpl_mode_bwd': (c_int, [POINTER(TmplArgs)] + [c_void_p] * 8 + [c_size_t, c_void_p]),
    'scae_caps_ll_fwd': (c_int, [POINTER(CapsArgs), POINTER(CapsOutputs), c_void_p]),
    'scae_caps_ll_bwd_workspace_bytes': (c_size_t, [POINTER(CapsArgs)]),
    'scae_caps_ll_bwd': (c_int, [POINTER(CapsArgs), POINTER(CapsSaved), POINTER(CapsUpstream)] + [c_void_p] * 6 +
                         [c_size_t, c_void_p]),
    'scae_caps_explicit_fwd': (c_int, [POINTER(CapsExplicitArgs), POINTER(CapsOutputs), c_void_p]),
    'scae_caps_explicit_bwd_workspace_bytes': (c_size_t, [POINTER(CapsExplicitArgs)]),
    'scae_caps_explicit_bwd': (c_int, [POINTER(CapsExplicitArgs), POINTER(CapsSaved), POINTER(CapsUpstream)] +
                               [c_void_p] * 7 + [c_size_t, c_void_p]),
    'scae_colsum_workspace_bytes': (c_size_t, [c_long, c_int]),
    'scae_colsum': (c_int, [c_void_p, c_long, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'scae_layernorm_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_long, c_int, c_void_p, c_void_p, c_void_p]),
    'scae_layernorm_bwd_workspace_bytes': (c_size_t, [c_long, c_int]),
    'scae_layernorm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    'scae_bias_act_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'scae_bias_act_bwd_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'scae_bias_act_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_size_t, c_void_p]),
    'scae_sab_fwd': (c_int, [c_void_p, c_void_p, POINTER(SabParams), c_int, c_int, c_void_p, c_void_p]),
    'scae_sab_bwd_workspace_bytes': (c_size_t, [c_int, c_int]),
    'scae_sab_bwd': (c_int, [c_void_p, c_void_p, POINTER(SabParams), c_void_p, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    'scae_pose_transform': (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p]),
    'scae_loss_head_workspace_bytes': (c_size_t, [POINTER(LossHeadArgs)]),
    'scae_loss_head_fwd': (c_int, [POINTER(LossHeadArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'scae_loss_head_bwd': (c_int, [POINTER(LossHeadArgs)] + [c_void_p] * 6 + [c_size_t, c_void_p]),
    'scae_loss_head_fwd_rows': (c_int, [POINTER(LossHeadArgs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'scae_loss_head_fwd_finish': (c_int, [POINTER(LossHeadArgs), c_void_p, c_void_p, c_void_p, c_void_p]),
    'scae_rmsprop_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_float, c_float, c_float, c_float,
                                  c_void_p]),
    'scae_attnpool_fwd': (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_void_p]),
    'scae_attnpool_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p]),
    'scae_conv_cols_supported': (c_int, [c_int] * 5),
    'scae_im2col3x3': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'scae_col2im3x3': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'scae_transpose_batched': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'scae_attnpool_cl_supported': (c_int, [c_long, c_int, c_int, c_int]),
    'scae_attnpool_cl_fwd': (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p]),
    'scae_attnpool_cl_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """Loads the shared library (once) and declares every prototype.  Raises ScaeError when it is unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScaeError(f'{LIB_PATH} is missing: build it with `python -m torch_scae_b200.build` '
                        '(or __graft_entry__.build()); torch_scae_b200 has no CPU or PyTorch fallback')
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise ScaeError(f'cannot load {LIB_PATH}: {e}') from e
    for name, (restype, argtypes) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ScaeError(f'{LIB_PATH} does not export {name}; rebuild it') from e
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.scae_abi_version() != ABI_VERSION:
        raise ScaeError(f'ABI mismatch: library {lib.scae_abi_version()} != binding {ABI_VERSION}; rebuild')
    from . import build as _build
    have, want = lib.scae_build_id().decode(), _build.source_id()
    if have != want:
        raise ScaeError(f'{LIB_PATH} was built from other sources (build id {have}, sources {want}); rebuild it with '
                        '`python -m torch_scae_b200.build`')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().scae_last_error().decode('utf-8', 'replace')
        raise ScaeError(f'{what} failed with code {rc}: {msg}')


def ptr(t):
    """Device pointer of a tensor (or NULL for None).  The tensor must be CUDA, fp32/int, contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ScaeError('torch_scae_b200 kernels need CUDA tensors; there is no CPU fallback')
    if not t.is_contiguous():
        raise ScaeError('internal error: non-contiguous tensor passed to a kernel')
    return t.data_ptr()
