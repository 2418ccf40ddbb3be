"""Case tables shared by make_golden.py (which needs the reference) and the tests (which do not)."""

DECODER_CASES = dict(
    alpha_c1=dict(B=3, M=4, C=1, tsize=(5, 5), osize=(12, 10), alpha=True, learn_scale=False, bg_value=True,
                  presence=True, bg_image=False),
    alpha_c3_nopres=dict(B=2, M=3, C=3, tsize=(6, 4), osize=(9, 11), alpha=True, learn_scale=True, bg_value=True,
                         presence=False, bg_image=False),
    temp_c3_bgimage=dict(B=2, M=3, C=3, tsize=(7, 9), osize=(12, 10), alpha=False, learn_scale=True,
                         bg_value=False, presence=True, bg_image=True),
    temp_c1=dict(B=2, M=5, C=1, tsize=(5, 5), osize=(8, 8), alpha=False, learn_scale=False, bg_value=True,
                 presence=False, bg_image=False),
)


CAPSULE_CASES = dict(
    default=dict(B=3, O=4, V=5, F=12, D=8, hidden=(16,), learn_vote_scale=True, allow_deformations=True,
                 noise_type='uniform', noise_scale=4., similarity=False, presence=True),
    similarity_plain=dict(B=2, O=3, V=4, F=10, D=6, hidden=(8,), learn_vote_scale=False, allow_deformations=False,
                          noise_type=None, noise_scale=0., similarity=True, presence=False),
    wide=dict(B=2, O=35, V=6, F=8, D=6, hidden=(8,), learn_vote_scale=True, allow_deformations=True,
              noise_type='uniform', noise_scale=4., similarity=False, presence=True),
)


def tiny_model_params(**scae):
    return dict(image_shape=(1, 20, 20), n_classes=10, n_part_caps=5, n_obj_caps=4,
                pcae_cnn_encoder_params=dict(out_channels=[8] * 4, strides=[2, 1, 1, 1]),
                pcae_template_generator_params=dict(template_size=(5, 5)),
                ocae_encoder_set_transformer_params=dict(dim_out=16),
                ocae_decoder_capsule_params=dict(dim_caps=8, hidden_sizes=(16,)),
                scae_params=dict(reconstruct_alternatives=False, **scae))


SCAE_CASES = dict(
    enc=dict(),
    soft=dict(vote_type='soft', presence_type='soft', stop_grad_caps_target=False),
    hard=dict(vote_type='hard', presence_type='hard', recon_mse_weight=0.5, part_caps_sparsity_weight=0.1,
              posterior_sparsity_loss_type='kl'),
    # the other sparsity-loss types of object_decoder.py:431-493: entropy on the prior, l2 on the posterior
    sparse=dict(prior_sparsity_loss_type='entropy', posterior_sparsity_loss_type='l2',
                prior_within_example_sparsity_weight=1.3, prior_between_example_sparsity_weight=0.6),
)



# whole-model cases that also change the model outside the SCAE wrapper: colour images, temperature-mode decoder with a
# learnt output scale (part_decoder.py:215-223)
SCAE_MODEL_OVERRIDES = dict(
    color_temp=dict(image_shape=(3, 20, 20), pcae_decoder_params=dict(use_alpha_channel=False, learn_output_scale=True)),
)
SCAE_CASES['color_temp'] = dict(vote_type='soft', presence_type='enc')


def scae_case_params(case):
    params = tiny_model_params(**SCAE_CASES[case])
    params.update(SCAE_MODEL_OVERRIDES.get(case, {}))
    return params
