"""Drop-in for the reference's Dino/modules/vision_transformer.py."""
from ccd_b200.encoder import VisionTransformer, vit_tiny, vit_small, vit_base  # noqa: F401
from ccd_b200.head import DINOHead  # noqa: F401
