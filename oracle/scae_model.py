"""Oracle: functional restatement of the whole SCAE forward + loss on a reference-compatible state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used for whole-model parity (same weights loaded into the CUDA
build and into this), and as the "port" CPU baseline in bench.py.  The two hot paths call the op-for-op
restatements in ``template_likelihood`` / ``capsule_likelihood``; everything else follows the reference's op
sequence too (including the per-capsule MLP loops) so that the CPU timing is representative of the reference.

``sd``  : dict name -> tensor with the reference's ``state_dict`` key names ("Weights").
``cfg`` : the dict returned by ``factory.prepare_model_params`` (factory.py:10-149).
``noise``: optional dict(part_presence [B,M], caps [B,O,1], vote [B,O,V]) of already-scaled additive noises; if a
key is missing the noise is drawn with ``torch.rand_like`` in the reference's call order.
"""
import math

import torch
import torch.nn.functional as F

from . import capsule_likelihood as cl
from . import template_likelihood as tl
from .pose import pose_to_affine
from .template_likelihood import log_safe


def _linear(sd, prefix, x, bias=True):
    return F.linear(x, sd[prefix + '.weight'], sd.get(prefix + '.bias') if bias else None)


def _mlp(sd, prefix, x):
    """nn_ext.py:19-31: Linear+ReLU pairs, final ReLU included (activate_final=True everywhere in the reference)."""
    i = 0
    while f'{prefix}.{i}.weight' in sd:
        x = F.relu(F.linear(x, sd[f'{prefix}.{i}.weight'], sd.get(f'{prefix}.{i}.bias')))
        i += 2
    return x


def _activation(name):
    """nn_utils.py:55-66."""
    if name == 'sigmoid':
        return torch.sigmoid
    if name == 'relu1':
        return lambda t: F.relu6(t * 6.) / 6.                                  # nn_ext.py:139-140
    return getattr(F, name)


# ---------------------------------------------------------------------------------------------- part encoder

def part_encoder(sd, cfg, image, noise=None, training=True):
    """part_encoder.py:86-113 (+ nn_ext.py:34-59 conv stack, :76-101 attention pooling)."""
    enc, cnn = cfg['pcae_encoder'], cfg['pcae_cnn_encoder']
    M, P, S = enc['n_caps'], enc['n_poses'], enc['n_special_features']
    h = image
    n_conv = len(cnn['out_channels'])
    for i, stride in enumerate(cnn['strides']):
        h = F.conv2d(h, sd[f'part_encoder.encoder.network.{2 * i}.weight'],
                     sd[f'part_encoder.encoder.network.{2 * i}.bias'], stride=stride)
        if i < n_conv - 1 or cnn.get('activate_final', True):
            h = F.relu(h)
    h = h + sd['part_encoder.img_embedding_bias'].unsqueeze(0)
    h = F.conv2d(h, sd['part_encoder.att_conv.weight'], sd['part_encoder.att_conv.bias'])
    B = h.shape[0]
    D = P + 1 + S
    h = h.view(B, M, D + 1, -1)
    h = (h[:, :, :-1] * F.softmax(h[:, :, -1:], -1)).sum(-1)                   # [B,M,D]
    pose, logit, feature = torch.split(h, [P, 1, S], -1)
    logit = logit.squeeze(-1)
    ns = enc.get('noise_scale', 4.)
    if training and ns > 0.:                                                   # part_encoder.py:105-107
        n = noise.get('part_presence') if noise else None
        if n is None:
            n = (torch.rand_like(logit) - .5) * ns
        logit = logit + n
    return dict(pose=pose_to_affine(pose, enc['similarity_transform']), presence=torch.sigmoid(logit),
                feature=feature if S > 0 else None)


# ---------------------------------------------------------------------------------------------- templates

def template_generator(sd, cfg, feature, batch_size):
    """part_decoder.py:75-110."""
    tg = cfg['pcae_template_generator']
    raw = _activation(tg['template_nonlin'])(sd['template_generator.template_logits'])
    if tg['colorize_templates'] and feature is not None:
        B, M, _ = feature.shape
        color = _mlp(sd, 'template_generator.templates_color_mlp', feature.reshape(B * M, -1))
        if tg['color_nonlin'] == 'relu1':
            color = color + .99
        color = _activation(tg['color_nonlin'])(color).view(B, M, -1)
        return raw * color[:, :, :, None, None]
    return raw.repeat(batch_size, 1, 1, 1, 1)


# ---------------------------------------------------------------------------------------------- set transformer

def _attention(sd, prefix, n_heads, q, k, v, presence=None):
    """set_transformer.py:24-104."""
    B, N, _ = q.shape
    Mk = k.shape[1]
    H = n_heads
    qp, kp, vp = (_linear(sd, f'{prefix}.{n}_projector', t) for n, t in (('q', q), ('k', k), ('v', v)))

    def heads(t, L):
        return t.view(B, L, H, -1).permute(2, 0, 1, 3).reshape(H * B, L, -1)

    qh, kh, vh = heads(qp, N), heads(kp, Mk), heads(vp, Mk)
    routing = torch.matmul(qh, kh.transpose(1, 2))
    if presence is not None:
        routing = routing - (1. - presence.repeat(H, 1).unsqueeze(-2)) * 1e32
    routing = F.softmax(routing / math.sqrt(qh.shape[-1]), -1)
    o = torch.matmul(routing, vh)
    o = o.view(H, B, N, -1).permute(1, 2, 0, 3).reshape(B, N, -1)
    return _linear(sd, f'{prefix}.o_projector', o)


def _mab(sd, prefix, n_heads, layer_norm, q, k, presence=None):
    """set_transformer.py:107-133."""
    h = _attention(sd, prefix + '.mqkv', n_heads, q, k, k, presence) + q
    if presence is not None:
        h = h * presence.unsqueeze(-1)
    d = h.shape[-1]
    if layer_norm:
        h = F.layer_norm(h, (d,), sd[prefix + '.ln0.weight'], sd[prefix + '.ln0.bias'])
    h = h + F.relu(_linear(sd, prefix + '.fc', h))
    if layer_norm:
        h = F.layer_norm(h, (d,), sd[prefix + '.ln1.weight'], sd[prefix + '.ln1.bias'])
    return h


def set_transformer(sd, cfg, x, presence=None):
    """set_transformer.py:174-223 (SAB and ISAB layers)."""
    st = cfg['ocae_encoder_set_transformer']
    H, ln = st['n_heads'], st['layer_norm']
    B = x.shape[0]
    h = _linear(sd, 'obj_encoder.fc1', x)
    for i in range(st['n_layers']):
        p = f'obj_encoder.sabs.{i}'
        if st.get('n_inducing_points') is None:
            h = _mab(sd, p + '.mab', H, ln, h, h, presence)
        else:
            ind = _mab(sd, p + '.mab0', H, ln, sd[p + '.I'].repeat(B, 1, 1), h, presence)
            h = _mab(sd, p + '.mab1', H, ln, h, ind)
    z = _linear(sd, 'obj_encoder.fc2', h)
    seeds = sd['obj_encoder.seeds'].repeat(B, 1, 1)
    return _attention(sd, 'obj_encoder.multi_head_attention', H, seeds, z, z, presence)


# ---------------------------------------------------------------------------------------------- object decoder

def capsule_mlps(sd, cfg, obj_encoding):
    """object_decoder.py:137-158: O separate MLP pairs, looped exactly like the reference. -> all_param [B,O,A]."""
    O = cfg['ocae_decoder_capsule']['n_caps']
    B = obj_encoding.shape[0]
    p = 'obj_decoder.capsule_layer'
    raw = torch.stack([_mlp(sd, f'{p}.mlps.{i}', obj_encoding[:, i]) for i in range(O)], 1)
    ext = torch.cat([raw, torch.ones(B, O, 1, dtype=raw.dtype, device=raw.device)], -1)
    return torch.stack([_mlp(sd, f'{p}.caps_mlps.{i}', ext[:, i]) for i in range(O)], 1)


def object_decoder(sd, cfg, obj_encoding, part_pose, part_presence, noise=None):
    """object_decoder.py:393-428."""
    cc = cfg['ocae_decoder_capsule']
    if cc.get('caps_dropout_rate', 0.) != 0.:
        raise NotImplementedError('caps_dropout_rate>0 hits a NameError in the reference (object_decoder.py:152,:196)')
    all_param = capsule_mlps(sd, cfg, obj_encoding)
    p = 'obj_decoder.capsule_layer'
    B, O, _ = all_param.shape
    V = cc['n_votes']
    nt, ns = cc['noise_type'], cc['noise_scale']
    n_caps = n_vote = None
    if nt == 'uniform':                                                         # object_decoder.py:198-212
        n_caps = noise.get('caps') if noise else None
        n_vote = noise.get('vote') if noise else None
        if n_caps is None:
            n_caps = (torch.rand(B, O, 1, dtype=all_param.dtype, device=all_param.device) - .5) * ns
        if n_vote is None:
            n_vote = (torch.rand(B, O, V, dtype=all_param.dtype, device=all_param.device) - .5) * ns
    elif nt:
        raise ValueError(f'Invalid noise type: {nt}')
    return cl.object_decoder_post_mlp(
        all_param, sd[p + '.cpr_static'], [sd[f'{p}.caps_bias_list.{i}'] for i in range(4)],
        sd['obj_decoder.dummy_vote'], part_pose, part_presence, n_caps, n_vote,
        similarity=cc['similarity_transform'], learn_vote_scale=cc['learn_vote_scale'],
        allow_deformations=cc['allow_deformations'])


# ---------------------------------------------------------------------------------------------- SCAE

def _decoder_params(sd, cfg):
    d = cfg['pcae_decoder']
    out = dict(bg_mixing_logit=sd['part_decoder.bg_mixing_logit'], bg_value=sd.get('part_decoder.bg_value'))
    if d['use_alpha_channel']:
        out['templates_alpha'] = sd['part_decoder.templates_alpha']
    else:
        out['temperature_logit'] = sd['part_decoder.temperature_logit']
    if d['learn_output_scale']:
        out['scale'] = sd['part_decoder.scale']
    return out


def scae_forward(sd, cfg, image, noise=None, training=True):
    """stacked_capsule_auto_encoder.py:92-215 with reconstruct_alternatives=False."""
    sc = cfg['scae']
    B = image.shape[0]
    pe = part_encoder(sd, cfg, image, noise, training)
    templates = template_generator(sd, cfg, pe['feature'], B)
    inp = torch.cat([pe['pose'], 1. - pe['presence'].unsqueeze(-1)], -1)
    inp_presence = pe['presence']
    if sc['stop_grad_caps_input']:
        inp, inp_presence = inp.detach(), inp_presence.detach()
    if pe['feature'] is not None:
        inp = torch.cat([inp, pe['feature']], -1)
    tflat = (templates.detach() if sc['stop_grad_caps_input'] else templates).reshape(B, templates.shape[1], -1)
    obj_encoding = set_transformer(sd, cfg, torch.cat([inp, tflat], -1), inp_presence)
    tpose, tpres = pe['pose'], pe['presence']
    if sc['stop_grad_caps_target']:
        tpose, tpres = tpose.detach(), tpres.detach()
    res = object_decoder(sd, cfg, obj_encoding, tpose, tpres, noise)
    res['part_presence'] = pe['presence']
    vt, pt = sc['vote_type'], sc['presence_type']
    if vt not in ('enc', 'soft', 'hard'):
        raise ValueError(f'Invalid vote_type: {vt}')
    if pt not in ('enc', 'soft', 'hard'):
        raise ValueError(f'Invalid presence_type: {pt}')
    dec_pose = {'enc': pe['pose'], 'soft': res['soft_winner'], 'hard': res['winner']}[vt]
    dec_pres = {'enc': pe['presence'], 'soft': res['soft_winner_presence'], 'hard': res['winner_presence']}[pt]
    loc, sigma, logits = tl.decode(templates, dec_pose, cfg['pcae_decoder']['output_size'], dec_pres,
                                   **_decoder_params(sd, cfg))
    res.update(rec=dict(loc=loc, sigma=sigma, logits=logits), templates=templates,
               template_presence=pe['presence'], transformed_templates=loc, part_pose=pe['pose'])
    if sc['n_classes'] is not None:                                             # :203-213 (sic: prior head twice)
        def head(t):
            return F.softmax(F.linear(t, sd['prior_classifier.0.weight'], sd['prior_classifier.0.bias']), -1)
        res['prior_cls_prob'] = head(res['caps_presence'].detach())
        res['posterior_cls_prob'] = head(res['posterior_mixing_prob'].sum(-1).detach())
    return res


def _sparsity(kind, p, n_classes=None, within_example_constant=None):
    """object_decoder.py:431-493."""
    if kind == 'l2':
        B, O = p.shape
        wc = float(O) / n_classes if within_example_constant is None else within_example_constant
        return torch.mean((p.sum(1) - wc) ** 2), torch.mean((p.sum(0) - float(B) / n_classes) ** 2)
    if kind in ('entropy', 'kl'):
        k = 1 if kind == 'entropy' else int(p.shape[-1])

        def xe(q):
            return torch.mean(-torch.sum(q * log_safe(q * k), dim=-1))
        within = p / (p.sum(1, keepdim=True) + 1e-8)
        tot = p.sum(0)
        between = tot / (tot.sum(0, keepdim=True) + 1e-8)
        return xe(within), -xe(between)
    raise ValueError(f'Invalid sparsity loss: {kind}')


def scae_loss(res, cfg, target, label=None):
    """stacked_capsule_auto_encoder.py:217-287.  Returns (loss, log dict)."""
    sc = cfg['scae']
    g = lambda k, d=0.: sc.get(k, d)
    rec = res['rec']
    ll = tl.mixture_log_prob(rec['loc'], rec['sigma'], rec['logits'], target)
    rec_ll = ll.reshape(ll.shape[0], -1).sum(-1).mean()
    loss = -rec_ll
    log = dict(rec_ll_loss=-rec_ll)
    if g('recon_mse_weight') > 0:
        mse = ((target - tl.mixture_mode(rec['loc'], rec['sigma'], rec['logits'])) ** 2)
        mse = mse.reshape(mse.shape[0], -1).sum(-1).mean()
        loss = loss + g('recon_mse_weight') * mse
        log['mse'] = mse
    if g('part_caps_sparsity_weight') > 0:
        l1 = res['part_presence'].sum(-1).mean()
        loss = loss + g('part_caps_sparsity_weight') * l1
        log['part_caps_loss'] = l1
    loss = loss - g('caps_ll_weight') * res['log_prob']
    log['log_prob_loss'] = -res['log_prob']
    pw, pb = g('prior_within_example_sparsity_weight'), g('prior_between_example_sparsity_weight')
    if pw > 0 or pb > 0:
        w, b = _sparsity(g('prior_sparsity_loss_type', 'l2'), res['caps_presence'], sc['n_classes'],
                         g('prior_within_example_constant', None))
        loss = loss + pw * w + pb * b
        log.update(prior_within_sparsity_loss=w, prior_between_sparsity_loss=b)
        # gated by the *prior* weights in the reference (:258-259)
        V = res['posterior_mixing_prob'].shape[-1]
        w, b = _sparsity(g('posterior_sparsity_loss_type', 'entropy'), res['posterior_mixing_prob'].sum(-1) / V,
                         sc['n_classes'])
        loss = loss + g('posterior_within_example_sparsity_weight') * w \
            + g('posterior_between_example_sparsity_weight') * b
        log.update(posterior_within_sparsity_loss=w, posterior_between_sparsity_loss=b)
    loss = loss + g('cpr_dynamic_reg_weight') * res['cpr_dynamic_reg_loss']
    log['cpr_dynamic_reg_loss'] = res['cpr_dynamic_reg_loss']
    if label is not None:
        a = F.cross_entropy(res['prior_cls_prob'], label)                       # on softmax outputs (sic) :281-282
        b = F.cross_entropy(res['posterior_cls_prob'], label)
        loss = loss + a + b
        log.update(prior_cls_xe=a, posterior_cls_xe=b)
    return loss, log


def accuracy(res, label):
    """stacked_capsule_auto_encoder.py:289-297."""
    a = (res['prior_cls_prob'].argmax(-1) == label).float().mean()
    b = (res['posterior_cls_prob'].argmax(-1) == label).float().mean()
    return torch.max(a, b)
