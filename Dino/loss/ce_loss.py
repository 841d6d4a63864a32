"""Drop-in for the reference's Dino/loss/ce_loss.py (the loss DINO_Finetune uses)."""
from ccd_b200.finetune import TFLoss  # noqa: F401
