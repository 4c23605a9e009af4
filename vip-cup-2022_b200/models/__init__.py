"""Backbones of the shipped registry (ckpts/ckpts.json of the reference) on the B200 kernels."""
from .resnet_rs import ResNetRS, ResNetRS50, ResNetRS101, ResNetRS152, ResNetRS200  # noqa: F401
from .gcvit import GCViT, GCViTBase, GCViTSmall, GCViTTiny, GCViTXTiny, GCViTXXTiny  # noqa: F401,E402
from .convnext import ConvNeXt  # noqa: F401,E402
from .efficientnet import EfficientNet  # noqa: F401,E402
from .nfnet import ECANFNetL0  # noqa: F401,E402
from .resnest import ResNeSt50  # noqa: F401,E402
