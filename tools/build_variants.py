"""Experiment builds of the library with extra -D switches (never shipped as the product path):
    python tools/build_variants.py dbg1:CCD_DBG_EPI=1 dbg2:CCD_DBG_EPI=2 ...
-> ccd_b200/libccd_b200_<name>.so, selected at run time with CCD_LIB=<path>."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccd_b200 import lib

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(lib.HERE, f"libccd_b200_{name}.so")
    lib.build(force=True, defines=[d for d in defs.split(",") if d], out=out, tag="_" + name)
    print("built", out)
