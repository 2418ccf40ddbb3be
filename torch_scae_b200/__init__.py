"""torch_scae_b200: B200-native (sm_100a) implementation of SCAE's two likelihood hot paths behind the module API of
bdsaglam/torch-scae.

    from torch_scae_b200 import factory
    model = factory.make_scae(dict(image_shape=(1, 40, 40), n_classes=10, n_part_caps=40, n_obj_caps=32)).cuda()
    res = model(image); loss, log = model.loss(res, image, label); loss.backward()

The fused kernels live in csrc/ and are reached through the C ABI in include/scae_b200.h; they require a CUDA
device and the built library (``python -m torch_scae_b200.build``).  There is no CPU fallback.
"""
from . import factory
from .attrdict import AttrDict, LazyAttrDict
from .distributions import GaussianMixture, TemplateMixture
from .object_decoder import CapsuleLayer, CapsuleLikelihood, CapsuleObjectDecoder
from .part_decoder import TemplateBasedImageDecoder, TemplateGenerator
from .part_encoder import CapsuleImageEncoder, CNNEncoder
from .set_transformer import SetTransformer
from .stacked_capsule_auto_encoder import SCAE

StackedCapsuleAutoEncoder = SCAE          # the name BASELINE.json uses

__all__ = ['factory', 'AttrDict', 'LazyAttrDict', 'GaussianMixture', 'TemplateMixture', 'CapsuleLayer',
           'CapsuleLikelihood', 'CapsuleObjectDecoder', 'TemplateBasedImageDecoder', 'TemplateGenerator',
           'CapsuleImageEncoder', 'CNNEncoder', 'SetTransformer', 'SCAE', 'StackedCapsuleAutoEncoder']
