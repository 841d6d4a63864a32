"""Drop-in for the reference's Dino/model/dino_vision.py (pretraining part, :21-115)."""
from ccd_b200.model import ABIDINOModel, ClusterMaps  # noqa: F401
