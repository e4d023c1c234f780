"""vit_pytorch_robust — B200-native drop-in for the hot path of RandallBalestriero/noise-robust-vit.

Same import path and class names as the reference package (reference __init__.py:1 exports
SimpleViT; its line 7 imports a module that does not exist, which this package does not repeat).
"""
from .simple_vit import SimpleViT  # noqa: F401
from . import vit  # noqa: F401
from .vit import VisionTransformer, vit_b_16, vit_b_32, vit_l_16, vit_l_32, vit_h_14, ViT  # noqa: F401
from .optim import FusedAdamW, clip_grad_norm_  # noqa: F401
from .functional import softmax_cross_entropy, add_gaussian_noise  # noqa: F401
from .parallel import DataParallel  # noqa: F401
