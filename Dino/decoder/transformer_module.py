"""Drop-in for the reference's Dino/decoder/transformer_module.py."""
from ccd_b200.finetune import MultiHeadAttention, PositionalEncoding, PositionwiseFeedForward  # noqa: F401
