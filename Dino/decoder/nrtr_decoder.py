"""Drop-in for the reference's Dino/decoder/nrtr_decoder.py."""
from ccd_b200.finetune import NRTRDecoder  # noqa: F401
