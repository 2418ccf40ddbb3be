"""Hyper-parameter defaults and model construction, API-compatible with the reference's ``factory`` module.

``prepare_model_params`` / ``make_scae`` keep the reference signatures (factory.py:10-23, :152) and return the
same resolved dictionaries (checked against reference output in tests/golden/factory.json), so the Hydra YAML of
the reference's experiment (configs/model/mnist.yaml) can be passed through unchanged.
"""
from .object_decoder import CapsuleLayer, CapsuleObjectDecoder
from .part_decoder import TemplateBasedImageDecoder, TemplateGenerator
from .part_encoder import CapsuleImageEncoder, CNNEncoder
from .set_transformer import SetTransformer
from .stacked_capsule_auto_encoder import SCAE


def _merged(defaults, overrides, reserved=()):
    overrides = overrides or {}
    for key in reserved:                     # keys the factory derives itself (factory.py:32,:42,:52-54, ...)
        assert key not in overrides
    out = dict(defaults)
    out.update(overrides)
    return out


def prepare_model_params(image_shape, n_classes, n_part_caps, n_obj_caps, pcae_cnn_encoder_params=None,
                         pcae_encoder_params=None, pcae_template_generator_params=None, pcae_decoder_params=None,
                         ocae_encoder_set_transformer_params=None, ocae_decoder_capsule_params=None,
                         scae_params=None):
    """Resolves user overrides against the reference defaults (factory.py:24-149)."""
    cnn = _merged(dict(input_shape=image_shape, out_channels=[128] * 4, kernel_sizes=[3, 3, 3, 3],
                       strides=[2, 2, 1, 1], activate_final=True),
                  pcae_cnn_encoder_params, ('input_shape',))
    enc = _merged(dict(input_shape=image_shape, n_caps=n_part_caps, n_poses=6, n_special_features=16,
                       similarity_transform=False),
                  pcae_encoder_params, ('input_shape',))
    gen = _merged(dict(n_templates=enc['n_caps'], n_channels=image_shape[0], template_size=(11, 11),
                       template_nonlin='sigmoid', dim_feature=enc['n_special_features'], colorize_templates=True,
                       color_nonlin='sigmoid'),
                  pcae_template_generator_params, ('n_templates', 'n_channels', 'dim_feature'))
    dec = _merged(dict(n_templates=gen['n_templates'], template_size=gen['template_size'],
                       output_size=image_shape[1:], learn_output_scale=False, use_alpha_channel=True,
                       background_value=True),
                  pcae_decoder_params, ('n_templates', 'template_size', 'output_size'))
    # pose + features + (1 - presence) + flattened template; template_size[0] is used twice (factory.py:83-85)
    st_dim_in = enc['n_poses'] + gen['dim_feature'] + 1 + gen['n_channels'] * gen['template_size'][0] ** 2
    st = _merged(dict(n_layers=3, n_heads=1, dim_in=st_dim_in, dim_hidden=16, dim_out=256, n_outputs=n_obj_caps,
                      layer_norm=True),
                 ocae_encoder_set_transformer_params, ('_ocae_st_dim_in', 'n_obj_caps'))
    caps = _merged(dict(n_caps=st['n_outputs'], dim_feature=st['dim_out'], n_votes=dec['n_templates'], dim_caps=32,
                        hidden_sizes=(128,), caps_dropout_rate=0.0, learn_vote_scale=True, allow_deformations=True,
                        noise_type='uniform', noise_scale=4., similarity_transform=False),
                   ocae_decoder_capsule_params, ('n_caps', 'dim_feature', 'n_votes'))
    scae = _merged(dict(n_classes=n_classes, vote_type='enc', presence_type='enc', stop_grad_caps_input=True,
                        stop_grad_caps_target=True, caps_ll_weight=1., cpr_dynamic_reg_weight=10,
                        prior_sparsity_loss_type='l2', prior_within_example_sparsity_weight=2.0,
                        prior_between_example_sparsity_weight=0.35, posterior_sparsity_loss_type='entropy',
                        posterior_within_example_sparsity_weight=0.7,
                        posterior_between_example_sparsity_weight=0.2),
                   scae_params, ('n_classes',))
    return dict(image_shape=image_shape, n_classes=n_classes, n_part_caps=n_part_caps, n_obj_caps=n_obj_caps,
                pcae_cnn_encoder=cnn, pcae_encoder=enc, pcae_template_generator=gen, pcae_decoder=dec,
                ocae_encoder_set_transformer=st, ocae_decoder_capsule=caps, scae=scae)


def make_scae(model_params: dict):
    """Builds the five sub-modules and the SCAE wrapper (factory.py:152-178)."""
    cfg = prepare_model_params(**model_params)
    # sub-modules are constructed in the reference's order (cnn, part encoder, template generator, part decoder, object
    # encoder, object decoder) so that the same torch seed consumes the RNG stream in the same order; the stacked
    # per-capsule MLP weights (PerCapsuleMLP) are the one place whose initial draws are not stream-identical
    part_encoder = CapsuleImageEncoder(encoder=CNNEncoder(**cfg['pcae_cnn_encoder']), **cfg['pcae_encoder'])
    template_generator = TemplateGenerator(**cfg['pcae_template_generator'])
    part_decoder = TemplateBasedImageDecoder(**cfg['pcae_decoder'])
    obj_encoder = SetTransformer(**cfg['ocae_encoder_set_transformer'])
    obj_decoder = CapsuleObjectDecoder(CapsuleLayer(**cfg['ocae_decoder_capsule']))
    return SCAE(part_encoder=part_encoder, template_generator=template_generator, part_decoder=part_decoder,
                obj_encoder=obj_encoder, obj_decoder=obj_decoder, **cfg['scae'])
