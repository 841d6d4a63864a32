"""TEST INFRASTRUCTURE ONLY -- build recipe that lets the UNMODIFIED reference travel to the GPU box.

The reference (TongkunGuan/CCD) is pure Python.  Its "build" here is a byte-compilation: every `.py` of the pinned commit is
compiled FROM WHERE IT LIES under /root/reference into a sourceless `.pyc` tree under oracle/_ref/ (Python imports
`pkg/module.pyc` directly), and the YAML configs / charsets the code reads at run time are packed into ONE data blob
(oracle/_ref/data_files.json).  No reference source file is copied into this repository -- oracle/_ref/ holds build outputs only;
it is git-ignored (never in history) and NOT gpurun-ignored (it ships with the snapshot exactly like the in-tree
libccd_b200.so).  Provenance chain:
    oracle/ref_manifest.json (tracked)  SHA-256 of every reference source at the pinned commit
      -> build(): each source is re-hashed against it BEFORE it is compiled (a modified reference refuses to build)
      -> oracle/_ref/BUILD_MANIFEST.json  SHA-256 of every produced file + the source hashes they came from
      -> verify(): re-hashes the build outputs (integrity on the GPU box) and checks the recorded source hashes against the
         tracked manifest.

Users of oracle/_ref (the same three places that may use anything under oracle/):
  tests/                       reference-vs-product parity on the GPU (train.py driving the drop-in, full-size comparator)
  bench.py --impl reference    the reference's own CPU path (cpu_baseline.kind = "reference")
  bench.py --impl stock-cuda   BASELINE.md B1: the reference's train() + modules on CUDA
The product path (ccd_b200/, Dino/) never imports it.

    python oracle/build_ref.py            # build + verify;  --manifest rewrites oracle/ref_manifest.json from the source tree
"""
import hashlib
import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CCD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")
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
    """The build outputs are intact and were produced from the pinned, unmodified sources."""
    with open(os.path.join(dst, "BUILD_MANIFEST.json")) as f:
        built = json.load(f)
    if built["sources"] != load_manifest():
        raise RuntimeError("oracle/_ref was not built from the pinned reference (source hashes differ from oracle/ref_manifest.json)")
    if built["python"] != list(sys.version_info[:2]):
        raise RuntimeError(f"oracle/_ref was compiled by Python {built['python']}, this is {list(sys.version_info[:2])}")
    bad = [rel for rel, h in built["outputs"].items() if not os.path.isfile(os.path.join(dst, rel)) or _sha(os.path.join(dst, rel)) != h]
    if bad:
        raise RuntimeError(f"oracle/_ref: {len(bad)} build output(s) missing or altered, e.g. {bad[:3]}")
    return len(built["outputs"])


def available(dst=DST):
    try:
        verify(dst)
        return True
    except Exception:
        return False


def data_files(dst=DST):
    """{relative path: text} of the reference's run-time data files (YAML configs, charsets)."""
    with open(os.path.join(dst, "data_files.json")) as f:
        return json.load(f)


def materialize_data(workdir, dst=DST):
    """Writes the packed data files under `workdir` with their reference-relative paths (Config opens
    'Dino/configs/template.yaml' relative to the working directory, Dino/utils/utils.py:208)."""
    for rel, text in data_files(dst).items():
        out = os.path.join(workdir, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            f.write(text)


def build(src=SRC, dst=DST):
    """Byte-compile the pinned sources into oracle/_ref; returns the number of build outputs, or 0 when the source tree is
    absent (GPU box: the prebuilt tree that travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(src, "Dino")):
        return 0
    man = load_manifest()
    for rel, h in man.items():
        if _sha(os.path.join(src, rel)) != h:
            raise RuntimeError(f"{rel} under {src} differs from the pinned reference (oracle/ref_manifest.json)")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    outputs, data = {}, {}
    for rel in man:
        if rel.endswith(".py"):
            out_rel = rel[:-3] + ".pyc"
            out = os.path.join(dst, out_rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            py_compile.compile(os.path.join(src, rel), cfile=out, dfile=rel, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            outputs[out_rel] = _sha(out)
        else:
            with open(os.path.join(src, rel)) as f:
                data[rel] = f.read()
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(dst, "data_files.json"), "w") as f:
        json.dump(data, f)
    outputs["data_files.json"] = _sha(os.path.join(dst, "data_files.json"))
    with open(os.path.join(dst, "BUILD_MANIFEST.json"), "w") as f:
        json.dump({"sources": man, "outputs": outputs, "python": list(sys.version_info[:2])}, f, indent=0, sort_keys=True)
    return verify(dst)


if __name__ == "__main__":
    if "--manifest" in sys.argv:
        print("manifest:", len(write_manifest()), "files")
    print("oracle/_ref:", build(), "build outputs (sourceless .pyc + data blob), verified")
