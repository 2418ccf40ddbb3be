"""Mixture distributions.

``GaussianMixture`` keeps the reference's eager class (distributions.py:20-89) for user code that builds a mixture
from explicit tensors.  ``TemplateMixture`` is what the fused decoder returns: it has the same interface
(``log_prob``, ``mode``, ``mean``, ``mixing_log_prob``, ``n_components``, ``dist``, ``mixing_logits``) but holds only
the decoder *inputs*; ``log_prob`` launches the fused sm_100a kernel (no B x K x C x H x W tensor is formed) and the
materialised views are rendered on demand.
"""
import torch
import torch.nn.functional as F
from torch.distributions import Normal

from . import ops


class GaussianMixture:
    def __init__(self, normal_dist: Normal, mixing_logits):
        self.dist = normal_dist
        self.mixing_logits = mixing_logits

    @property
    def n_components(self):
        return self.mixing_logits.shape[1]

    def mixing_log_prob(self):
        return F.log_softmax(self.mixing_logits, 1)

    def mean(self):
        return torch.sum(F.softmax(self.mixing_logits, 1) * self.dist.mean, 1)

    def log_prob(self, x):
        return torch.logsumexp(self.dist.log_prob(x.unsqueeze(1)) + self.mixing_log_prob(), 1)

    def mode(self, straight_through_gradient=False, maximum=False):
        loc = self.dist.loc
        mlp = self.mixing_log_prob()
        if maximum:
            mlp = mlp + self.dist.log_prob(loc)
        mask = F.one_hot(mlp.argmax(1), mlp.shape[1]).movedim(-1, 1)
        if straight_through_gradient:
            soft = F.softmax(mlp, 1)
            mask = (mask - soft).detach() + soft
        return torch.sum(mask * loc, 1)

    @classmethod
    def make_from_stats(cls, loc, scale, mixing_logits):
        return cls(Normal(loc, scale), mixing_logits)


class TemplateMixture:
    """Lazy per-pixel mixture over M warped templates + background (part_decoder.py:233-243's ``rec_pdf``)."""

    def __init__(self, decoder, templates, pose, presence, bg_image, template_color=None):
        """``template_color`` (B,M,C): fused colourisation -- ``templates`` are then the batch-shared raw templates
        (1,M,C,h,w) and the per-image templates ``raw * colour`` are formed inside the kernels only."""
        self._decoder = decoder
        self._inputs = (templates, pose, presence, bg_image)
        self._color = template_color
        self._materialized = None

    def _full_templates(self):
        templates = self._inputs[0]
        if self._color is None:
            return templates
        return templates.reshape(1, *templates.shape[-4:]) * self._color[:, :, :, None, None]

    # ---- the fused hot path ---------------------------------------------------------------------------------------
    def _fused(self, x):
        templates, pose, presence, bg_image = self._inputs
        d = self._decoder
        return ops.TemplateMixtureLogProb.apply(
            templates, pose, presence, bg_image, x,
            d.templates_alpha if d.use_alpha_channel else None,
            d.bg_value if d.background_value else None,
            d.bg_mixing_logit if d.use_alpha_channel else None,
            None if d.use_alpha_channel else d.temperature_logit,
            d.scale if d.learn_output_scale else None, tuple(d.output_size), self._color)

    def log_prob(self, x):
        """(B,C,H,W) per-pixel log-likelihood; GaussianMixture.log_prob (distributions.py:41-44)."""
        return self._fused(x)[0]

    def log_likelihood(self, x):
        """(B,) = log_prob(x) summed over (C,H,W) inside the kernel (what SCAE.loss reduces to, :220-221)."""
        return self._fused(x)[1]

    # ---- materialised views (off the hot path) --------------------------------------------------------------------
    @property
    def n_components(self):
        return self._inputs[0].shape[-4] + 1

    def _render(self, *want):
        templates, pose, presence, bg_image = self._inputs
        d = self._decoder
        return ops.template_render(
            templates, pose, presence, bg_image, d.templates_alpha if d.use_alpha_channel else None,
            d.bg_value if d.background_value else None, d.bg_mixing_logit if d.use_alpha_channel else None,
            None if d.use_alpha_channel else d.temperature_logit, d.scale if d.learn_output_scale else None,
            tuple(d.output_size), want, self._color)

    def materialize(self):
        """(transformed_templates, mixing_logits) as the reference's decoder returns them.

        Under ``torch.no_grad()`` (or when nothing requires grad) this is one render-kernel launch.  When gradients
        are needed through these tensors (e.g. ``recon_mse_weight > 0`` uses ``mode()``), they are built with
        differentiable PyTorch CUDA ops instead -- off the hot path, see DESIGN.md section 7.
        """
        if self._materialized is None:
            needs_grad = torch.is_grad_enabled() and any(
                t is not None and t.requires_grad
                for t in self._inputs + (self._color,) + tuple(self._decoder.parameters()))
            if needs_grad:
                self._materialized = self._decoder.differentiable_materialize(self._full_templates(),
                                                                              *self._inputs[1:])
            else:
                r = self._render('transformed_templates', 'mixing_logits')
                self._materialized = (r['transformed_templates'], r['mixing_logits'])
        return self._materialized

    @property
    def transformed_templates(self):
        return self.materialize()[0]

    @property
    def mixing_logits(self):
        return self.materialize()[1]

    @property
    def dist(self):
        return Normal(self.transformed_templates, self._decoder.output_scale())

    def _eager(self):
        loc, logits = self.materialize()
        return GaussianMixture(Normal(loc, self._decoder.output_scale()), logits)

    def mixing_log_prob(self):
        return F.log_softmax(self.mixing_logits, 1)

    def _fast_point_estimate(self, which):
        needs_grad = torch.is_grad_enabled() and any(
            t is not None and t.requires_grad
            for t in self._inputs + (self._color,) + tuple(self._decoder.parameters()))
        if needs_grad or self._materialized is not None:
            return None
        return self._render(which)[which]

    def mean(self):
        fast = self._fast_point_estimate('mean')
        return fast if fast is not None else self._eager().mean()

    def mode(self, straight_through_gradient=False, maximum=False):
        # `maximum` adds the same constant (-log sigma - log sqrt(2 pi)) to every component, so the argmax is unchanged
        if not straight_through_gradient:
            fast = self._fast_point_estimate('mode')
            if fast is not None:
                return fast
            if self._materialized is None:
                # gradients wanted (recon_mse_weight > 0, stacked_capsule_auto_encoder.py:226-230): render + mode-backward
                # kernels; the B x (M+1) x C x H x W tensors are not formed in either direction
                templates, pose, presence, bg_image = self._inputs
                d = self._decoder
                return ops.TemplateMixtureMode.apply(
                    templates, pose, presence, bg_image, d.templates_alpha if d.use_alpha_channel else None,
                    d.bg_value if d.background_value else None, d.bg_mixing_logit if d.use_alpha_channel else None,
                    None if d.use_alpha_channel else d.temperature_logit, d.scale if d.learn_output_scale else None,
                    tuple(d.output_size), self._color)
        return self._eager().mode(straight_through_gradient, maximum)
