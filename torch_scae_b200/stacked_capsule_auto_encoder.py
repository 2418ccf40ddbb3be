"""Stacked Capsule Auto-Encoder: wiring of the five sub-modules, composite loss, accuracy.

Constructor, ``forward(image) -> AttrDict``, ``loss(res, target, label) -> (loss, log)`` and
``calculate_accuracy`` follow the reference (stacked_capsule_auto_encoder.py:22-297), quirks included (SURVEY.md
section 9).  Differences are only in *when* things are computed: the reconstruction pdf is lazy (its likelihood is a
fused kernel launched from ``loss``), ``res.transformed_templates`` is rendered on first read, and the alternative
reconstructions cost nothing unless somebody looks at them.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .attrdict import LazyAttrDict
from .object_decoder import sparsity_loss

_VOTE_SOURCES = ('enc', 'soft', 'hard')


class SCAE(nn.Module):
    """Stacked Capsule Auto-Encoder"""

    def __init__(self, part_encoder, template_generator, part_decoder, obj_encoder, obj_decoder, n_classes=None,
                 vote_type='soft', presence_type='enc', stop_grad_caps_input=True, stop_grad_caps_target=True,
                 recon_mse_weight=0, part_caps_sparsity_weight=0., cpr_dynamic_reg_weight=0., caps_ll_weight=0.,
                 prior_sparsity_loss_type='l2', prior_within_example_sparsity_weight=0.,
                 prior_between_example_sparsity_weight=0., prior_within_example_constant=None,
                 posterior_sparsity_loss_type='entropy', posterior_within_example_sparsity_weight=0.,
                 posterior_between_example_sparsity_weight=0., reconstruct_alternatives=True,
                 sync_batch_stats=False):
        super().__init__()
        self.part_encoder = part_encoder
        self.template_generator = template_generator
        self.part_decoder = part_decoder
        self.obj_encoder = obj_encoder
        self.obj_decoder = obj_decoder
        self.n_classes = n_classes
        self.vote_type = vote_type
        self.presence_type = presence_type
        self.stop_grad_caps_input = stop_grad_caps_input
        self.stop_grad_caps_target = stop_grad_caps_target
        if n_classes:
            n_caps = obj_decoder.n_obj_capsules
            self.prior_classifier = nn.Sequential(nn.Linear(n_caps, n_classes), nn.Softmax(-1))
            self.posterior_classifier = nn.Sequential(nn.Linear(n_caps, n_classes), nn.Softmax(-1))
        else:
            self.prior_classifier = None
            self.posterior_classifier = None
        self.cpr_dynamic_reg_weight = cpr_dynamic_reg_weight
        self.caps_ll_weight = caps_ll_weight
        self.recon_mse_weight = recon_mse_weight
        self.prior_sparsity_loss_type = prior_sparsity_loss_type
        self.prior_within_example_sparsity_weight = prior_within_example_sparsity_weight
        self.prior_between_example_sparsity_weight = prior_between_example_sparsity_weight
        self.prior_within_example_constant = prior_within_example_constant
        self.posterior_sparsity_loss_type = posterior_sparsity_loss_type
        self.posterior_within_example_sparsity_weight = posterior_within_example_sparsity_weight
        self.posterior_between_example_sparsity_weight = posterior_between_example_sparsity_weight
        self.part_caps_sparsity_weight = part_caps_sparsity_weight
        self.reconstruct_alternatives = reconstruct_alternatives
        # Extension (not in the reference's signature): under data parallelism the between-example sparsity statistics
        # are batch-GLOBAL sums; False = per-shard statistics, what the reference computes under Lightning DDP; True =
        # statistics of the global batch (two O-float all-reduces per step), for N-GPU == 1-GPU equivalence.
        self.sync_batch_stats = sync_batch_stats

    def forward(self, image, noise=None):
        """``noise``: optional dict(part_presence, caps, vote) of pre-scaled noises replacing the internal draws."""
        noise = noise or {}
        B = image.shape[0]
        enc = self.part_encoder(image, presence_noise=noise.get('part_presence'))
        # Fused colourisation: when the templates are coloured per image and the object
        # encoder does not need a gradient through them, the decoder gets (raw templates, colours) and the kernels
        # colour on the fly -- the differentiable (B,M,C,h,w) product and its backward are never formed.  The object
        # encoder's detached template features are the only materialised copy.
        gen = self.template_generator
        fused = self.stop_grad_caps_input and hasattr(gen, 'color') and hasattr(gen, 'raw_templates')
        color = gen.color(enc.feature) if fused else None
        if color is not None:
            raw = gen.raw_templates()
            with torch.no_grad():
                templates = raw * color[:, :, :, None, None]            # values only (features / logging)

            def decode(pose, presence, rep=1, detach=False):
                r, c = (raw.detach(), color.detach()) if detach else (raw, color)
                return self.part_decoder(templates=r, pose=pose, presence=presence,
                                         template_color=c if rep == 1 else c.repeat_interleave(rep, dim=0))
        else:
            templates = gen(feature=enc.feature, batch_size=B).templates

            def decode(pose, presence, rep=1, detach=False):
                t = templates.detach() if detach else templates
                return self.part_decoder(templates=t if rep == 1 else t.repeat_interleave(rep, dim=0), pose=pose,
                                         presence=presence)

        # object encoder input: [pose, 1 - presence, features, flattened templates] per part
        part_param = torch.cat([enc.pose, 1. - enc.presence.unsqueeze(-1)], -1)
        part_presence = enc.presence
        flat_templates = templates
        if self.stop_grad_caps_input:
            part_param, part_presence = part_param.detach(), part_presence.detach()
            flat_templates = templates.detach()
        if enc.feature is not None:
            part_param = torch.cat([part_param, enc.feature], -1)
        flat_templates = flat_templates.reshape(*flat_templates.shape[:2], -1)
        obj_encoding = self.obj_encoder(torch.cat([part_param, flat_templates], -1), part_presence)

        target_pose, target_presence = enc.pose, enc.presence
        if self.stop_grad_caps_target:
            target_pose, target_presence = target_pose.detach(), target_presence.detach()
        caps_noise = (noise['caps'], noise['vote']) if 'caps' in noise else None
        dec = self.obj_decoder(obj_encoding, target_pose, target_presence, noise=caps_noise)
        res = LazyAttrDict(dec)
        res.part_presence = enc.presence

        if self.vote_type not in _VOTE_SOURCES:
            raise ValueError(f'Invalid vote_type: {self.vote_type}')
        if self.presence_type not in _VOTE_SOURCES:
            raise ValueError(f'Invalid presence_type: {self.presence_type}')
        dec_pose = {'enc': enc.pose, 'soft': res.soft_winner, 'hard': res.winner}[self.vote_type]
        dec_presence = {'enc': enc.presence, 'soft': res.soft_winner_presence,
                        'hard': res.winner_presence}[self.presence_type]
        res.rec = decode(dec_pose, dec_presence)

        if self.reconstruct_alternatives:
            # all lazy: nothing is launched until a validation/logging step reads the pdf or the rendered tensors
            with torch.no_grad():
                pres = enc.presence.detach()
                res.bottom_up_rec = decode(enc.pose.detach(), pres, detach=True)
                res.top_down_rec = decode(res.winner.detach(), pres, detach=True)
                O = res.vote.shape[1]
                td_presence = pres.repeat_interleave(O, dim=0) * res.vote_presence_binary.view(-1, pres.shape[1])
                res.top_down_per_caps_rec = decode(res.vote.detach().view(-1, *res.vote.shape[2:]), td_presence, O,
                                                   detach=True)
        if color is not None:
            # the differentiable product exists only if somebody reads it (a user loss / regulariser on res.templates gets
            # its gradient; the training step itself never forms the (B,M,C,h,w) tensor with a graph)
            res.set_lazy('templates', lambda: raw * color[:, :, :, None, None])
        else:
            res.templates = templates
        res.template_presence = enc.presence
        rec = res.rec
        res.set_lazy('transformed_templates', lambda: rec.transformed_templates)
        if self.n_classes is not None:
            assert self.prior_classifier is not None
            assert self.posterior_classifier is not None
            # lazy: on CUDA the loss head (csrc/loss_head.cu) evaluates both heads inside SCAE.loss and stores its
            # probabilities here; read before that (or on the PyTorch path) they are computed by the modules
            caps_presence, posterior = res.caps_presence, res.posterior_mixing_prob
            res.set_lazy('prior_cls_prob', lambda: self.prior_classifier(caps_presence.detach()))
            # sic: the reference feeds the posterior mass through the *prior* head (:211)
            res.set_lazy('posterior_cls_prob', lambda: self.prior_classifier(posterior.sum(-1).detach()))
        return res

    def _fused_loss_head(self, res, label):
        """Sparsity losses + classifier cross-entropies through ops.loss_head, or None (-> the PyTorch ops below).
        SCAE_B200_LOSS_HEAD=0 in the environment forces the PyTorch ops (A/B timing, bisecting)."""
        if os.environ.get('SCAE_B200_LOSS_HEAD', '1') == '0':
            return None
        sparsity = self.prior_within_example_sparsity_weight > 0 or self.prior_between_example_sparsity_weight > 0
        cp, post = res.get('caps_presence'), res.get('posterior_mixing_prob')
        if not (torch.is_tensor(cp) and cp.is_cuda and torch.is_tensor(post)):
            return None
        if label is not None:
            head = self.prior_classifier
            if not (self.n_classes is not None and isinstance(head, nn.Sequential) and len(head) == 2
                    and isinstance(head[0], nn.Linear) and isinstance(head[1], nn.Softmax)
                    and isinstance(res, LazyAttrDict) and 'prior_cls_prob' in res
                    and not res.is_materialized('prior_cls_prob') and not res.is_materialized('posterior_cls_prob')):
                return None
        from . import ops
        return ops.loss_head(cp, post, label, self.prior_classifier[0] if label is not None else None, self.n_classes,
                             self.prior_sparsity_loss_type, self.posterior_sparsity_loss_type,
                             (self.prior_within_example_sparsity_weight, self.prior_between_example_sparsity_weight,
                              self.posterior_within_example_sparsity_weight,
                              self.posterior_between_example_sparsity_weight),
                             self.prior_within_example_constant, sparsity, self.sync_batch_stats)

    def loss(self, res, reconstruction_target, label=None):
        log = dict()
        pdf = res.rec.pdf
        if hasattr(pdf, 'log_likelihood'):       # fused: per-image sum comes straight out of the kernel
            rec_ll = pdf.log_likelihood(reconstruction_target).mean()
        else:
            per_pixel = pdf.log_prob(reconstruction_target)
            rec_ll = per_pixel.view(per_pixel.shape[0], -1).sum(-1).mean()
        loss = -rec_ll
        log.update(rec_ll_loss=-rec_ll)
        if self.recon_mse_weight > 0:
            mse_per_pixel = (reconstruction_target - pdf.mode()) ** 2
            mse = mse_per_pixel.view(mse_per_pixel.shape[0], -1).sum(-1).mean()
            loss = loss + self.recon_mse_weight * mse
            log.update(mse=mse)
        if self.part_caps_sparsity_weight > 0:
            part_caps_l1 = res.part_presence.sum(-1).mean()
            loss = loss + self.part_caps_sparsity_weight * part_caps_l1
            log.update(part_caps_loss=part_caps_l1)
        loss = loss - self.caps_ll_weight * res.log_prob
        log.update(log_prob_loss=-res.log_prob)
        head = self._fused_loss_head(res, label)
        if head is not None:
            # the (B,O)-sized tail -- sparsity terms (:243-271) and classifier cross-entropies (:279-285) -- as one
            # kernel pair per direction; `total` is their weighted sum, `terms` the individual values for the log
            from .ops import LOSS_HEAD_TERMS
            total, terms, probs = head
            loss = loss + total + self.cpr_dynamic_reg_weight * res.cpr_dynamic_reg_loss
            if self.prior_within_example_sparsity_weight > 0 or self.prior_between_example_sparsity_weight > 0:
                log.update({k: terms[i] for i, k in enumerate(LOSS_HEAD_TERMS[:4])})
            log.update(cpr_dynamic_reg_loss=res.cpr_dynamic_reg_loss)
            if label is not None:
                assert self.n_classes is not None
                log.update({k: terms[4 + i] for i, k in enumerate(LOSS_HEAD_TERMS[4:])})
                res.prior_cls_prob, res.posterior_cls_prob = probs[0], probs[1]
            return loss, log
        # both sparsity terms are gated by the *prior* weights in the reference (:243-244, :258-259)
        if self.prior_within_example_sparsity_weight > 0 or self.prior_between_example_sparsity_weight > 0:
            within, between = sparsity_loss(self.prior_sparsity_loss_type, res.caps_presence,
                                            n_classes=self.n_classes,
                                            within_example_constant=self.prior_within_example_constant,
                                            sync_batch_stats=self.sync_batch_stats)
            loss = loss + self.prior_within_example_sparsity_weight * within \
                + self.prior_between_example_sparsity_weight * between
            log.update(prior_within_sparsity_loss=within, prior_between_sparsity_loss=between)
            n_points = res.posterior_mixing_prob.shape[-1]
            mass = res.posterior_mixing_prob.sum(-1)
            within, between = sparsity_loss(self.posterior_sparsity_loss_type, mass / n_points,
                                            n_classes=self.n_classes, sync_batch_stats=self.sync_batch_stats)
            loss = loss + self.posterior_within_example_sparsity_weight * within \
                + self.posterior_between_example_sparsity_weight * between
            log.update(posterior_within_sparsity_loss=within, posterior_between_sparsity_loss=between)
        loss = loss + self.cpr_dynamic_reg_weight * res.cpr_dynamic_reg_loss
        log.update(cpr_dynamic_reg_loss=res.cpr_dynamic_reg_loss)
        if label is not None:
            assert self.n_classes is not None
            prior_cls_xe = F.cross_entropy(res.prior_cls_prob, target=label)          # on softmax outputs (sic)
            posterior_cls_xe = F.cross_entropy(res.posterior_cls_prob, target=label)
            loss = loss + prior_cls_xe + posterior_cls_xe
            log.update(prior_cls_xe=prior_cls_xe, posterior_cls_xe=posterior_cls_xe)
        return loss, log
    def calculate_accuracy(self, res, label: torch.Tensor):
        prior_acc = (res.prior_cls_prob.argmax(-1) == label).float().mean()
        posterior_acc = (res.posterior_cls_prob.argmax(-1) == label).float().mean()
        return torch.max(prior_acc, posterior_acc)
