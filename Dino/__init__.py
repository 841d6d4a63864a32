"""Drop-in package: the module surface the reference's train.py:27-35 imports, backed by ccd_b200 (sm_100a kernels)."""
