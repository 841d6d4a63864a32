"""Drop-in for the reference's Dino/loss/Dino_loss.py."""
from ccd_b200.loss import DINOLoss, SegLoss  # noqa: F401
