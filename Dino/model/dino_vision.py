"""Drop-in for the reference's Dino/model/dino_vision.py: ABIDINOModel (pretraining, :21-115) and DINO_Finetune (:135-290)."""
from ccd_b200.finetune import DINO_Finetune, Mlp  # noqa: F401
from ccd_b200.model import ABIDINOModel, ClusterMaps  # noqa: F401
