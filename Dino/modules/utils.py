"""The subset of the reference's Dino/modules/utils.py the pretraining step calls (train.py:96,131,249-250)."""
from ccd_b200.train_utils import (clip_gradients, cancel_gradients_last_layer, get_params_groups, has_batchnorms,  # noqa: F401
                                  cosine_iter_scheduler)
