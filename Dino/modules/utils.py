"""Drop-in for the part of the reference's Dino/modules/utils.py that train.py calls (train.py:47-48,96,117,131-137,
144-171,186-216,249-250): the step glue on the sm_100a kernels and the host-side helpers around it."""
from ccd_b200.encoder import trunc_normal_  # noqa: F401
from ccd_b200.host_utils import (LARS, MetricLogger, SmoothedValue, bool_flag, fix_random_seeds, get_rank, get_world_size,  # noqa: F401
                                 init_distributed_mode, is_dist_avail_and_initialized, is_main_process,
                                 restart_from_checkpoint, save_on_master, setup_for_distributed)
from ccd_b200.train_utils import (cancel_gradients_last_layer, clip_gradients, cosine_iter_scheduler, get_params_groups,  # noqa: F401
                                  has_batchnorms)
