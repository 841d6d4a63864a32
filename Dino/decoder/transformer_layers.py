"""Drop-in for the reference's Dino/decoder/transformer_layers.py (decoder layer; the encoder-side layers there are unused by CCD)."""
from ccd_b200.finetune import TFDecoderLayer  # noqa: F401
