from .clip import CLIP
from .siglip import SigLIP
from .vit import VisionTransformer

__all__ = [
    "VisionTransformer",
    "CLIP",
    "SigLIP",
]
