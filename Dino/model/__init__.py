import os as _os

# modules of the reference that this package does not replace resolve in the reference checkout (see Dino/__init__.py)
_ref = _os.environ.get("CCD_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "Dino", "model")):
    __path__.append(_os.path.join(_ref, "Dino", "model"))
