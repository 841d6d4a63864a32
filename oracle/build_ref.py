"""TEST INFRASTRUCTURE ONLY -- copy recipe that lets the UNMODIFIED reference travel to the GPU box.

The reference (TongkunGuan/CCD) is pure Python: its "build" is a verbatim copy of its own source files from where they
lie under /root/reference into oracle/_ref/ (git-ignored, so no reference source ever enters this repository's history;
NOT gpurun-ignored, so the copy ships with the snapshot exactly like the in-tree libccd_b200.so).  Nothing is edited:
`oracle/ref_manifest.json` (tracked) records the SHA-256 of every file at the pinned reference commit and
`verify()` re-checks the copy against it, so a test or bench arm that runs from oracle/_ref can state that it executed
the reference's own code.

Users of oracle/_ref (the same three places that may use anything under oracle/):
  tests/                       reference-vs-product parity on the GPU (train.py driving the drop-in, B1 stock-CUDA modules)
  bench.py --impl reference    the reference's own CPU path (cpu_baseline.kind = "reference")
  bench.py --impl stock-cuda   BASELINE.md B1: the reference's modules on CUDA wrapped like train.py:93-110
The product path (ccd_b200/, Dino/) never imports it.

    python oracle/build_ref.py            # copy + verify;  --manifest rewrites oracle/ref_manifest.json from the source tree
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CCD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")
# what travels: the python sources of the path and of its callers + the YAML configs / charsets they read at run time
KEEP_EXT = (".py", ".yaml", ".txt")
SKIP_DIRS = {".git", "graph", "__pycache__"}


def _files(root):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith(KEEP_EXT):
                out.append(os.path.relpath(os.path.join(d, f), root))
    return out


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def write_manifest(src=SRC):
    man = {rel: _sha(os.path.join(src, rel)) for rel in _files(src)}
    with open(MANIFEST, "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)
    return man


def load_manifest():
    with open(MANIFEST) as f:
        return json.load(f)


def verify(dst=DST):
    """Every manifest file present in the copy with the recorded hash (i.e. the reference, unmodified)."""
    man = load_manifest()
    bad = [rel for rel, h in man.items() if not os.path.isfile(os.path.join(dst, rel)) or _sha(os.path.join(dst, rel)) != h]
    if bad:
        raise RuntimeError(f"oracle/_ref differs from the pinned reference in {len(bad)} file(s), e.g. {bad[:3]}")
    return len(man)


def available(dst=DST):
    try:
        verify(dst)
        return True
    except Exception:
        return False


def build(src=SRC, dst=DST):
    """Copy the manifest's files verbatim; returns the number of files, or 0 when the source tree is absent (GPU box:
    the prebuilt copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(src, "Dino")):
        return 0
    man = load_manifest()
    for rel in man:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
    return verify(dst)


if __name__ == "__main__":
    if "--manifest" in sys.argv:
        print("manifest:", len(write_manifest()), "files")
    print("oracle/_ref:", build(), "files copied and verified")
