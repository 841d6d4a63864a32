"""Drop-in for the reference's Dino/modules/segmentor.py."""
from ccd_b200.segmentor import SegHead, MLAHead, Conv_MLA  # noqa: F401
