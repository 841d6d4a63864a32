"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (TongkunGuan/CCD at /root/reference)
in this container so the oracle restatement (oracle/ccd_oracle.py) can be pinned against it and golden
vectors can be generated (tests/golden/make_golden.py).

/root/reference does not exist on the GPU box; there the reference is the hash-verified, byte-compiled tree
oracle/_ref built by oracle/build_ref.py (build outputs only, git-ignored, travels with the snapshot).  Users: tests/, bench.py's reference
arms (--impl reference / stock-cuda) and nothing else; the product path never imports this module.  Only third-party *imports* that are missing from this image are stubbed; no reference
arithmetic is replaced except skimage.measure.label, which is restated with scipy.ndimage.label
(8-connectivity, raster-order labels -- the same labelling skimage's default connectivity produces;
reference call site Dino/utils/DBSCAN.py:80).
"""
import importlib.machinery
import json
import math
import os
import random
import sys
import types
import warnings
from pathlib import Path

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """CCD_REFERENCE_ROOT if set; else the reference checkout of this container; else the sourceless byte-compiled tree that
    oracle/build_ref.py built from it (oracle/_ref, hash-checked against oracle/ref_manifest.json) -- the only form in which
    the reference reaches the GPU box."""
    env = os.environ.get("CCD_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/Dino"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Dino"))


def reference_sources_available() -> bool:
    """True only for a SOURCE checkout (tests that read the reference's text, e.g. an ast scan of train.py)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import numpy as np
    import PIL
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from scipy import ndimage
    from torch.utils.data import Dataset

    # fastai.vision is used purely as a star-import namespace (Dino/model/dino_vision.py:6)
    if "fastai" not in sys.modules:
        fa = _mod("fastai")
        fav = _mod(
            "fastai.vision", nn=nn, F=F, np=np, math=math, json=json, Path=Path, PIL=PIL, warnings=warnings,
            random=random, Dataset=Dataset, PathOrStr=str, tensor=torch.tensor,
            ifnone=lambda a, b: b if a is None else a, torch=torch,
        )
        fav.__all__ = [k for k in fav.__dict__ if not k.startswith("__")]
        fa.vision = fav

    # skimage.measure.label -> scipy 8-connected labelling (Dino/utils/DBSCAN.py:6,80)
    if "skimage" not in sys.modules:
        def label(mask, *a, **k):
            return ndimage.label(np.asarray(mask) != 0, structure=np.ones((3, 3), dtype=int))[0]

        sk = _mod("skimage")
        sk.measure = _mod("skimage.measure", label=label)

    # never executed on the pretraining path, only imported
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "pylab", "mkl", "faiss", "lmdb",
                 "imgaug", "editdistance", "natsort", "tensorboardX"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _mod(name)
    mp = sys.modules["matplotlib"]
    if not hasattr(mp, "pyplot"):
        mp.pyplot = sys.modules["matplotlib.pyplot"]
        mp.colors = sys.modules["matplotlib.colors"]


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's hot-path classes (unmodified code)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    try:
        # first import of torchvision registers torch.library fakes and walks the caller's frames with inspect; do it before
        # the reference's __file__-less namespace package `Dino` is on the stack (Dino/model/dino_vision.py:9 imports it)
        import torchvision  # noqa: F401
    except Exception:
        pass
    # our own repo ships a `Dino` drop-in package: make sure the reference's wins for this process
    for k in [k for k in sys.modules if k == "Dino" or k.startswith("Dino.")]:
        del sys.modules[k]
    # the reference's Dino/ has no __init__.py (namespace package): a regular `Dino` package anywhere on sys.path
    # (this repository's drop-in) would win over it, so hide those entries while the reference is imported
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [e for e in saved_path
                                      if not os.path.isfile(os.path.join(e or os.getcwd(), "Dino", "__init__.py"))]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vits = importlib.import_module("Dino.modules.vision_transformer")
            seg = importlib.import_module("Dino.modules.segmentor")
            loss = importlib.import_module("Dino.loss.Dino_loss")
            dv = importlib.import_module("Dino.model.dino_vision")
            mutils = importlib.import_module("Dino.modules.utils")
            dbscan = importlib.import_module("Dino.utils.DBSCAN")
    finally:
        sys.path[:] = saved_path
        ref_mods = {k: v for k, v in sys.modules.items() if k == "Dino" or k.startswith("Dino.")}
        for k in ref_mods:
            del sys.modules[k]
    _LOADED = types.SimpleNamespace(vits=vits, seg=seg, loss=loss, dv=dv, utils=mutils, dbscan=dbscan,
                                    modules=ref_mods)
    return _LOADED


def ensure_gloo_group():
    """DINOLoss.update_center calls dist.all_reduce unconditionally (Dino/loss/Dino_loss.py:139)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
