"""Drop-in package: the module surface the reference's train.py:27-35 imports, backed by ccd_b200 (sm_100a kernels).

Only the hot path is shipped here (model / modules / loss / decoder / convertor).  The reference's data pipeline, config and
logging code (Dino/dataset/*, Dino/utils/*, Dino/configs/*) stays in the reference checkout: point CCD_REFERENCE_ROOT at it
and those sub-packages resolve there, while everything this package defines keeps winning."""
import os as _os

_ref = _os.environ.get("CCD_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "Dino")):
    __path__.append(_os.path.join(_ref, "Dino"))
