"""Drop-in for the reference's Dino/convertor/attn.py."""
from ccd_b200.finetune import AttnConvertor  # noqa: F401
