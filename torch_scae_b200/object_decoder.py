"""Capsule object decoder (hot path 2) and the capsule sparsity losses.

API and ``state_dict`` layout of the reference's object_decoder.py (CapsuleLayer :28-240, CapsuleLikelihood :243-372,
CapsuleObjectDecoder :375-428, sparsity losses :431-493).  What changes underneath:

* the 2 x O per-capsule MLPs are evaluated as batched GEMMs over stacked weights (``PerCapsuleMLP``) instead of
  2 x O Python-looped ``nn.Sequential`` calls; the state dict still exposes ``mlps.<i>.<0|2>.{weight,bias}`` /
  ``caps_mlps.<i>.<0|2>.weight``;
* everything after the MLPs -- pose non-linearities, 3x3 vote composition, presence/scale heads, Gaussian mixture
  likelihood, posterior, hard/soft winners -- is ONE fused CUDA kernel per direction (csrc/caps_ll.cu) reached
  through ``ops.CapsuleVoteLikelihood``.
"""
import math

import torch
import torch.nn as nn

from . import _lib, math_ops, ops
from .attrdict import AttrDict
from .general_utils import prod


class _ParamHead(torch.autograd.Function):
    """Last Linear + ReLU of the per-capsule parameter MLPs, producing ``all_param`` directly in the contiguous
    (B, O, A) layout the fused kernel stages with bulk copies.

    The batched GEMM writes through a strided (O, B, A) view of the output buffer (row stride O*A, batch stride A), so
    the (O,B,A)->(B,O,A) transpose copy of the plain bmm chain disappears; the backward feeds the kernel's (B,O,A)
    gradient to the two GEMMs through the same strided view.  With ``grad_is_masked`` the caller promises that the
    incoming gradient already carries the ReLU mask (the kernel applies it, SCAE_CAPS_RELU_GRAD), so no separate
    threshold pass runs.
    """

    @staticmethod
    def forward(ctx, h, w, grad_is_masked):
        n, B, _ = h.shape                              # h (O, B, H), w (O, A, H)
        out = h.new_empty(B, n, w.shape[1])
        torch.bmm(h, w.transpose(1, 2), out=out.transpose(0, 1))
        torch.relu_(out)
        ctx.save_for_backward(h, w, out)
        ctx.grad_is_masked = grad_is_masked
        return out

    @staticmethod
    def backward(ctx, g):
        h, w, out = ctx.saved_tensors
        if not ctx.grad_is_masked:
            g = g * (out > 0)
        gt = g.transpose(0, 1)                         # (O, B, A) view, unit stride in A: a legal GEMM operand
        gh = torch.bmm(gt, w) if ctx.needs_input_grad[0] else None
        gw = torch.bmm(gt.transpose(1, 2), h) if ctx.needs_input_grad[1] else None
        return gh, gw, None


class PerCapsuleMLP(nn.Module):
    """``n`` independent Linear/ReLU chains (final ReLU included, like nn_ext.MLP) as batched matmuls.

    Parameters are stored stacked -- ``w<j>`` (n, out, in), ``b<j>`` (n, out) -- but saved/loaded under the
    reference's per-capsule names ``<i>.<2j>.weight`` / ``<i>.<2j>.bias`` (ModuleList of nn.Sequential,
    object_decoder.py:86-107).
    """

    def __init__(self, n, sizes, bias=True):
        super().__init__()
        self.n, self.sizes, self.has_bias = n, list(sizes), bias
        self.n_layers = len(sizes) - 1
        for j, (fan_in, fan_out) in enumerate(zip(sizes[:-1], sizes[1:])):
            bound = 1.0 / math.sqrt(fan_in)          # nn.Linear's default init range for weight and bias
            self.register_parameter(f'w{j}', nn.Parameter(torch.empty(n, fan_out, fan_in).uniform_(-bound, bound)))
            if bias:
                self.register_parameter(f'b{j}', nn.Parameter(torch.empty(n, fan_out).uniform_(-bound, bound)))

    def _hidden(self, x, n_layers):
        h = x.transpose(0, 1)                         # (n, B, in)
        for j in range(n_layers):
            w = getattr(self, f'w{j}')
            if self.has_bias:
                h = torch.baddbmm(getattr(self, f'b{j}').unsqueeze(1), h, w.transpose(1, 2))
            else:
                h = torch.bmm(h, w.transpose(1, 2))
            h = torch.relu(h)
        return h

    def forward(self, x):
        """x (B, n, in) -> (B, n, out)."""
        return self._hidden(x, self.n_layers).transpose(0, 1)

    def forward_contiguous(self, x, grad_is_masked=False):
        """Same values as ``forward`` for a bias-free last layer, as a contiguous (B, n, out) tensor (see _ParamHead)."""
        assert not self.has_bias
        h = self._hidden(x, self.n_layers - 1).contiguous()
        return _ParamHead.apply(h, getattr(self, f'w{self.n_layers - 1}'), grad_is_masked)

    # ---- reference-compatible (de)serialisation ---------------------------------------------------------------------
    def _named(self):
        for j in range(self.n_layers):
            yield f'{2 * j}.weight', getattr(self, f'w{j}')
            if self.has_bias:
                yield f'{2 * j}.bias', getattr(self, f'b{j}')

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for suffix, p in self._named():
            src = p if keep_vars else p.detach()
            for i in range(self.n):
                destination[f'{prefix}{i}.{suffix}'] = src[i]

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        expected = set()
        for suffix, p in self._named():
            for i in range(self.n):
                key = f'{prefix}{i}.{suffix}'
                expected.add(key)
                if key not in state_dict:
                    missing_keys.append(key)
                    continue
                value = state_dict[key]
                if tuple(value.shape) != tuple(p.shape[1:]):
                    error_msgs.append(f'size mismatch for {key}: copying a param with shape {tuple(value.shape)} '
                                      f'from checkpoint, the shape in current model is {tuple(p.shape[1:])}.')
                    continue
                with torch.no_grad():
                    p[i].copy_(value)
        if strict:
            unexpected_keys.extend(k for k in state_dict if k.startswith(prefix) and k not in expected)


class CapsuleLayer(nn.Module):
    """Predicts, for each of ``n_caps`` object capsules, ``n_votes`` part-pose votes with presences and scales."""

    n_transform_params = 6

    def __init__(self, n_caps, dim_feature, n_votes, dim_caps, hidden_sizes=(128,), caps_dropout_rate=0.0,
                 learn_vote_scale=False, allow_deformations=True, noise_type=None, noise_scale=0.,
                 similarity_transform=True):
        super().__init__()
        self.n_caps = n_caps
        self.dim_feature = dim_feature
        self.hidden_sizes = list(hidden_sizes)
        self.dim_caps = dim_caps
        self.caps_dropout_rate = caps_dropout_rate
        self.n_votes = n_votes
        self.learn_vote_scale = learn_vote_scale
        self.allow_deformations = allow_deformations
        self.noise_type = noise_type
        self.noise_scale = noise_scale
        self.similarity_transform = similarity_transform

        P = self.n_transform_params
        self.output_shapes = ([n_votes, P], [1, P], [1], [n_votes], [n_votes])
        self.splits = [prod(s) for s in self.output_shapes]
        self.n_outputs = sum(self.splits)             # A = 8V + 7
        self.mlps = PerCapsuleMLP(n_caps, [dim_feature] + self.hidden_sizes + [dim_caps])
        # no output bias: the static part of the object-part relationship lives in cpr_static / caps_bias_list
        self.caps_mlps = PerCapsuleMLP(n_caps, [dim_caps + 1] + self.hidden_sizes + [self.n_outputs], bias=False)
        self.caps_bias_list = nn.ParameterList(
            [nn.Parameter(torch.zeros(1, n_caps, *shape)) for shape in self.output_shapes[1:]])
        self.cpr_static = nn.Parameter(torch.zeros(1, n_caps, n_votes, P))

    def kernel_flags(self):
        """Flag word for the fused kernel.  RELU_GRAD: the backward kernel masks g_all_param with (all_param > 0), the
        backward of the final ReLU of ``caps_mlps`` (nn_ext.py:19-31), which ``predict_all_param`` then skips."""
        return (_lib.CAPS_SIMILARITY if self.similarity_transform else 0) \
            | (_lib.CAPS_LEARN_VOTE_SCALE if self.learn_vote_scale else 0) \
            | (_lib.CAPS_ALLOW_DEFORM if self.allow_deformations else 0) | _lib.CAPS_RELU_GRAD

    def predict_all_param(self, feature, grad_is_masked=False):
        """(B,O,F) -> all_param (B,O,A), contiguous: the per-capsule MLP half of the reference's forward (:137-158).

        ``grad_is_masked=True`` is for callers that hand all_param to the fused kernel with ``kernel_flags()`` (and to
        nothing else): the kernel's gradient then already contains the final ReLU's mask."""
        if self.caps_dropout_rate != 0.0:
            # the reference deletes `caps_exist` before using it (object_decoder.py:152,:196): unsupported there too
            raise NameError("name 'caps_exist' is not defined")
        raw = self.mlps(feature)
        ones = raw.new_ones(*raw.shape[:2], 1)
        return self.caps_mlps.forward_contiguous(torch.cat([raw, ones], -1), grad_is_masked)

    def draw_noise(self, all_param):
        """The two presence-logit noises, drawn like the reference (always on, also in eval mode; :198-212)."""
        B, O, _ = all_param.shape
        if self.noise_type == 'uniform':
            caps = (torch.rand(B, O, 1, device=all_param.device, dtype=all_param.dtype) - 0.5) * self.noise_scale
            vote = (torch.rand(B, O, self.n_votes, device=all_param.device, dtype=all_param.dtype) - 0.5) \
                * self.noise_scale
            return caps, vote
        if self.noise_type == 'logistic':
            raise NotImplementedError("noise_type='logistic' (a LogisticNormal sample in the reference, "
                                      "object_decoder.py:202-204) is not supported by the fused kernel")
        if not self.noise_type:
            return None, None
        raise ValueError(f'Invalid noise type: {self.noise_type}')

    def forward(self, feature, parent_transform=None, parent_presence=None):
        """Stand-alone use (no likelihood).  Returns the reference's keys, with ``vote`` as (B,O,V,3,3)."""
        if parent_transform is not None or parent_presence is not None:
            return self._forward_hierarchical(feature, parent_transform, parent_presence)
        all_param = self.predict_all_param(feature, grad_is_masked=True)
        B, O, _ = all_param.shape
        V = self.n_votes
        noise_caps, noise_vote = self.draw_noise(all_param)
        # the fused kernel needs part poses; with no likelihood requested feed zeros and ignore those outputs
        x = all_param.new_zeros(B, V, 6)
        dummy = all_param.new_zeros(1, 1, V, 6)
        out = dict(zip(ops.CAPS_RETURNS, ops.CapsuleVoteLikelihood.apply(
            all_param, self.cpr_static, *self.caps_bias_list, dummy, x, None, noise_caps, noise_vote,
            self.kernel_flags())))
        last_row = all_param.new_tensor([0., 0., 1.]).expand(B, O, V, 1, 3)
        vote = torch.cat([out['vote'].view(B, O, V, 2, 3), last_row], -2)
        return AttrDict(vote=vote, scale=out['scale'], vote_presence=out['vote_presence'],
                        presence_logit_per_caps=out['presence_logit_per_caps'],
                        presence_logit_per_vote=out['presence_logit_per_vote'],
                        cpr_dynamic_reg_loss=out['reg_per_example'].sum() / B)


    def _forward_hierarchical(self, feature, parent_transform, parent_presence):
        """``parent_transform`` (B,O,1,3,3) replaces the capsule -> viewer transform and / or ``parent_presence`` (B,O,1)
        the capsule presence (object_decoder.py:183-188, :214-217; stacked capsule layers).  Not on the hot path -- SCAE
        never passes them -- so this is the reference's sequence with PyTorch ops, differentiable w.r.t. both."""
        from . import cv_ops
        all_param = self.predict_all_param(feature)
        B, O, V = all_param.shape[0], self.n_caps, self.n_votes
        parts = torch.split(all_param, self.splits, -1)
        cpr_dynamic, cvr, logit_caps, logit_vote, scale = [t.view(B, O, *shape) for t, shape in zip(parts, self.output_shapes)]
        if not self.allow_deformations:
            cpr_dynamic = torch.zeros_like(cpr_dynamic)
        reg = math_ops.l2_loss(cpr_dynamic) / B
        as_matrix = lambda t: cv_ops.geometric_transform(t, self.similarity_transform, nonlinear=True, as_matrix=True)
        cpr = as_matrix(cpr_dynamic + self.cpr_static)
        cvr, logit_caps, logit_vote, scale = [t + bias for t, bias in zip((cvr, logit_caps, logit_vote, scale),
                                                                        self.caps_bias_list)]
        cvr = as_matrix(cvr) if parent_transform is None else parent_transform
        vote = torch.matmul(cvr.expand(B, O, V, 3, 3), cpr)
        noise_caps, noise_vote = self.draw_noise(all_param)
        if noise_caps is not None:
            logit_caps, logit_vote = logit_caps + noise_caps, logit_vote + noise_vote
        presence_caps = torch.sigmoid(logit_caps) if parent_presence is None else parent_presence
        vote_presence = presence_caps * torch.sigmoid(logit_vote)
        scale = torch.nn.functional.softplus(scale + .5) + 1e-2 if self.learn_vote_scale else torch.ones_like(scale)
        return AttrDict(vote=vote, scale=scale, vote_presence=vote_presence, presence_logit_per_caps=logit_caps,
                        presence_logit_per_vote=logit_vote, cpr_dynamic_reg_loss=reg)


class CapsuleLikelihood:
    """Capsule voting mechanism on explicit vote tensors (object_decoder.py:243-372).

    For code that builds the likelihood from its own votes (the reference's own test does).  CUDA tensors go through
    one kernel per direction (``ops.CapsuleExplicitLikelihood``, csrc/caps_explicit.cu: same mixture arithmetic as the
    fused decoder path minus the vote composition); host tensors run the reference's op sequence in PyTorch (this class
    is not on a hot path -- ``CapsuleObjectDecoder`` uses the fused kernels and never comes here).
    """

    def __init__(self, vote, scale, vote_presence, dummy_vote):
        self.n_caps = vote.shape[1]
        self.vote, self.scale, self.vote_presence, self.dummy_vote = vote, scale, vote_presence, dummy_vote

    def __call__(self, x, presence=None):
        if x.is_cuda and x.shape[-1] == 6 and self.vote.shape[-1] == 6:
            r = dict(zip(ops.EXPLICIT_RETURNS, ops.CapsuleExplicitLikelihood.apply(
                self.vote, self.scale, self.vote_presence, self.dummy_vote, x, presence)))
            ll = r.pop('ll_per_example')
            return AttrDict(log_prob=ll.mean(), **r)
        import torch.nn.functional as F
        B, V, P = x.shape
        O = self.n_caps
        dev = x.device
        s = self.scale.unsqueeze(-1)
        lp = (-((x.unsqueeze(1) - self.vote) ** 2) / (2 * s ** 2) - s.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)
        dummy = torch.full((B, 1, V), math.log(0.01), device=dev)
        lp = torch.cat([lp, dummy], 1)
        mixing_logit = torch.cat([math_ops.log_safe(self.vote_presence), dummy], 1)
        mixing_log_prob = mixing_logit - mixing_logit.logsumexp(1, keepdim=True)
        binary = (mixing_logit[:, :-1] > mixing_logit[:, -1:]).float()
        post_logit = mixing_logit + lp
        per_point = post_logit.logsumexp(1)
        if presence is not None:
            per_point = per_point * presence.float()
        log_prob = per_point.sum(1).mean()
        win = torch.argmax(post_logit[:, :-1], 1)
        winner = torch.gather(self.vote, 1, win.view(B, 1, V, 1).expand(B, 1, V, P)).squeeze(1)
        winner_presence = torch.gather(self.vote_presence, 1, win.unsqueeze(1)).squeeze(1)
        assert winner.shape == (B, V, P)
        assert winner_presence.shape == (B, V)
        post = F.softmax(post_logit, 1)
        votes = torch.cat([self.vote, self.dummy_vote.expand(B, 1, V, P)], 1)
        pres = torch.cat([self.vote_presence, torch.zeros(B, 1, V, device=dev)], 1)
        soft_winner = torch.sum(post.unsqueeze(-1) * votes, 1)
        soft_winner_presence = torch.sum(post * pres, 1)
        assert soft_winner.shape == (B, V, P)
        assert soft_winner_presence.shape == (B, V)
        return AttrDict(log_prob=log_prob, vote_presence_binary=binary, winner=winner, winner_presence=winner_presence,
                        soft_winner=soft_winner, soft_winner_presence=soft_winner_presence,
                        posterior_mixing_prob=post[:, :-1], mixing_log_prob=mixing_log_prob, mixing_logit=mixing_logit,
                        is_from_capsule=win // V)


class CapsuleObjectDecoder(nn.Module):
    def __init__(self, capsule_layer: CapsuleLayer):
        super().__init__()
        self.capsule_layer = capsule_layer
        self.dummy_vote = nn.Parameter(torch.zeros(1, 1, capsule_layer.n_votes, capsule_layer.n_transform_params))

    @property
    def n_obj_capsules(self):
        return self.capsule_layer.n_caps

    def forward(self, obj_encoding: torch.Tensor, part_pose: torch.Tensor, part_presence: torch.Tensor = None,
                noise=None):
        """obj_encoding (B,O,D), part_pose (B,M,6), part_presence (B,M)|None -> AttrDict with the reference's 17 keys.

        ``noise`` = (noise_caps (B,O,1), noise_vote (B,O,V)), already scaled, overrides the internally drawn presence
        noise (used by parity tests to inject the reference's draws).
        """
        layer = self.capsule_layer
        all_param = layer.predict_all_param(obj_encoding, grad_is_masked=True)
        noise_caps, noise_vote = noise if noise is not None else layer.draw_noise(all_param)
        out = dict(zip(ops.CAPS_RETURNS, ops.CapsuleVoteLikelihood.apply(
            all_param, layer.cpr_static, *layer.caps_bias_list, self.dummy_vote, part_pose, part_presence,
            noise_caps, noise_vote, layer.kernel_flags())))
        B = all_param.shape[0]
        res = AttrDict(vote=out['vote'], scale=out['scale'], vote_presence=out['vote_presence'],
                       presence_logit_per_caps=out['presence_logit_per_caps'],
                       presence_logit_per_vote=out['presence_logit_per_vote'],
                       cpr_dynamic_reg_loss=out['reg_per_example'].sum() / B,
                       caps_presence=out['caps_presence'],
                       log_prob=out['ll_per_example'].mean())
        for key in ('vote_presence_binary', 'winner', 'winner_presence', 'soft_winner', 'soft_winner_presence',
                    'posterior_mixing_prob', 'mixing_log_prob', 'mixing_logit', 'is_from_capsule'):
            res[key] = out[key]
        return res


# ---- capsule sparsity losses: (B,O)-sized, stay in PyTorch (object_decoder.py:431-493) ------------------------------

def _batch_stats(caps_presence, sync_batch_stats):
    """(sum over the batch (O,), batch size): of the local shard like the reference under Lightning DDP, or -- with
    ``sync_batch_stats`` -- of the global batch, summed over the data-parallel ranks."""
    from . import ddp
    total, batch_size = caps_presence.sum(0), caps_presence.shape[0]
    if sync_batch_stats:
        total, batch_size = ddp.global_sum(total), batch_size * ddp.world_size()
    return total, batch_size


def _between(term, sync_batch_stats):
    from . import ddp
    return ddp.global_stat_loss(term) if sync_batch_stats else term


def capsule_l2_loss(caps_presence, n_classes: int, within_example_constant=None, sync_batch_stats=False,
                    **unused_kwargs):
    del unused_kwargs
    num_caps = caps_presence.shape[1]
    if within_example_constant is None:
        within_example_constant = float(num_caps) / n_classes
    within_example = torch.mean((caps_presence.sum(1) - within_example_constant) ** 2)
    total, batch_size = _batch_stats(caps_presence, sync_batch_stats)
    between_example = torch.mean((total - float(batch_size) / n_classes) ** 2)
    return within_example, _between(between_example, sync_batch_stats)


def capsule_entropy_loss(caps_presence, k=1, sync_batch_stats=False, **unused_kwargs):
    del unused_kwargs
    within_prob = math_ops.normalize(caps_presence, 1)
    within_example = math_ops.cross_entropy_safe(within_prob, within_prob * k)
    total, _ = _batch_stats(caps_presence, sync_batch_stats)
    between_prob = math_ops.normalize(total, 0)
    between_example = math_ops.cross_entropy_safe(between_prob, between_prob * k)
    return within_example, _between(-between_example, sync_batch_stats)   # negated: this entropy is to be increased


def neg_capsule_kl(caps_presence, sync_batch_stats=False, **unused_kwargs):
    del unused_kwargs
    return capsule_entropy_loss(caps_presence, k=int(caps_presence.shape[-1]), sync_batch_stats=sync_batch_stats)


def sparsity_loss(loss_type, *args, **kwargs):
    if loss_type == 'l2':
        return capsule_l2_loss(*args, **kwargs)
    if loss_type == 'entropy':
        return capsule_entropy_loss(*args, **kwargs)
    if loss_type == 'kl':
        return neg_capsule_kl(*args, **kwargs)
    raise ValueError(f"Invalid sparsity loss: {loss_type}")
