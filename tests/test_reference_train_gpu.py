"""The reference's OWN training loop drives the drop-in.  oracle/run_ref_train.py imports the unmodified `train.py` (from the
hash-checked, byte-compiled tree oracle/_ref), replaces only its LMDB loader by a synthetic one, and calls `train(config)`:
   --impl reference : reference modules (the stock PyTorch-CUDA path)
   --impl dropin    : this repository's `Dino` package first on sys.path -- every `Dino.model / Dino.modules / Dino.loss` name
                      train.py uses resolves here, `Dino.utils / Dino.dataset` resolve in the reference tree.
Both start from the same checkpoint (loaded by train()'s own restart_from_checkpoint) and see the same batches; the loss traces
and the parameters after three iterations of (forward, backward, per-parameter clip, torch.optim.AdamW, `.data` EMA with a
LOW teacher momentum so a stale teacher copy would show) must agree."""
import json
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
RUNNER = os.path.join(ROOT, "oracle", "run_ref_train.py")


def _run(impl, dump, extra=(), port=29641):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK") and not k.startswith("TORCHELASTIC")}
    env.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    cmd = [sys.executable, RUNNER, "--impl", impl, "--arch", "vit_tiny", "--out_dim", "4096", "--batch", "8", "--iters", "3", "--warmup", "0",
           "--drop_path", "0.0", "--init-seed", "0", "--lr", "0.02", "--momentum-teacher", "0.5", "--dump-state", dump, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=tempfile.gettempdir())
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{"impl"')][-1]
    return json.loads(line), r.stdout


@pytest.mark.gpu
def test_reference_train_loop_drives_the_dropin():
    import ref_import
    if not ref_import.reference_available():
        pytest.skip("no reference tree (oracle/_ref)")
    with tempfile.TemporaryDirectory() as d:
        ref, _ = _run("reference", os.path.join(d, "ref.pt"), port=29641)
        ours, log = _run("dropin", os.path.join(d, "ours.pt"), port=29642)
        a, b = torch.load(os.path.join(d, "ref.pt")), torch.load(os.path.join(d, "ours.pt"))
    assert ours["dino_package"].startswith(ROOT) and ref["dino_package"].startswith("namespace:")
    assert "All keys matched successfully" in log                      # train()'s own checkpoint loader accepted the drop-in modules
    assert len(ours["losses"]) == len(ref["losses"]) == 3
    print("loss trace reference:", ref["losses"], "drop-in:", ours["losses"])
    for i, (x, y) in enumerate(zip(ours["losses"], ref["losses"])):
        assert abs(x - y) / abs(y) <= (1e-3 if i == 0 else 4e-3), (i, x, y)
    for x, y in zip(ours["mask_loss"], ref["mask_loss"]):
        assert abs(x - y) <= 3e-3
    assert ours["lr"] == ref["lr"]
    # state after 3 iterations: same names; teacher = EMA(0.5) of a student that moved with lr 0.02 -> a stale bf16 teacher
    # copy or a missed EMA would be far outside these bounds
    assert list(a["student"]) == list(b["student"]) and list(a["teacher"]) == list(b["teacher"])
    for part in ("student", "teacher"):
        num = sum(((a[part][k].double() - b[part][k].double()) ** 2).sum() for k in a[part] if a[part][k].dtype.is_floating_point)
        den = sum((a[part][k].double() ** 2).sum() for k in a[part] if a[part][k].dtype.is_floating_point)
        assert (num / den).sqrt().item() <= 2e-2, part
    moved = sum(((a["teacher"][k].double() - a["student"][k].double()) ** 2).sum() for k in a["teacher"] if k in a["student"])
    assert moved > 0
    assert (a["center"] - b["center"]).abs().max() <= 1e-3


@pytest.mark.needs_reference
def test_train_py_names_resolve_in_the_dropin():
    """CPU: every attribute train.py reads from `utils`, `vits` and every `from Dino.* import` resolves with the drop-in first on
    the path and CCD_REFERENCE_ROOT set (AST scan of the unmodified train.py; no CUDA needed)."""
    import ast
    import ref_import
    if not ref_import.reference_sources_available():
        pytest.skip("needs the reference's source text (a checkout, not the compiled oracle/_ref)")
    src = open(os.path.join(ref_import.REFERENCE_ROOT, "train.py")).read()
    tree = ast.parse(src)
    used = sorted({n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "utils"})
    imports = [(n.module, [a.name for a in n.names]) for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and n.module and
               n.module.startswith("Dino")]
    code = ("import sys, os\n"
            f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'oracle')!r}]\n"
            "import ref_import; ref_import._install_stubs()\n"
            "for n in ('imgaug.random', 'imgaug.augmenters', 'imgaug.augmenters.geometric', 'imgaug.augmentables', 'imgaug.augmentables.segmaps'):\n"
            "    ref_import._mod(n)\n"
            "sys.modules['imgaug.augmenters.geometric']._warp_affine_arr = None\n"
            "sys.modules['imgaug.augmentables.segmaps'].SegmentationMapsOnImage = object\n"
            "import importlib, Dino\n"
            "assert Dino.__file__.startswith(%r)\n" % ROOT +
            "from Dino.modules import utils\n"
            f"missing = [a for a in {used!r} if not hasattr(utils, a)]\n"
            "assert not missing, missing\n"
            f"for mod, names in {imports!r}:\n"
            "    m = importlib.import_module(mod)\n"
            "    for n in names:\n"
            "        assert hasattr(m, n), (mod, n)\n"
            "import Dino.model.dino_vision as dv, Dino.loss.Dino_loss as dl, Dino.modules.vision_transformer as vt\n"
            f"assert all(x.__file__.startswith({ROOT!r}) for x in (dv, dl, vt, utils))\n"
            "print('ok', len(missing))\n")
    env = dict(os.environ, CCD_REFERENCE_ROOT=ref_import.REFERENCE_ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300, cwd=tempfile.gettempdir())
    assert r.returncode == 0 and "ok 0" in r.stdout, r.stderr[-2000:]
    assert {"init_distributed_mode", "fix_random_seeds", "MetricLogger", "restart_from_checkpoint", "save_on_master", "get_world_size",
            "is_main_process", "LARS", "bool_flag", "clip_gradients", "cancel_gradients_last_layer", "get_params_groups",
            "has_batchnorms", "cosine_iter_scheduler"} <= set(used)


@pytest.mark.needs_reference
def test_reference_train_harness_runs_on_cpu():
    """CPU self-test of the harness (reference modules, gloo, .cuda() as a no-op): train() completes and logs finite losses."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29643")
    r = subprocess.run([sys.executable, RUNNER, "--impl", "reference", "--arch", "vit_tiny", "--out_dim", "512", "--batch", "2", "--iters", "2",
                        "--warmup", "0", "--init-seed", "0", "--cpu-debug"], capture_output=True, text=True, env=env, timeout=600,
                       cwd=tempfile.gettempdir())
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith('{"impl"')][-1])
    assert len(d["losses"]) == 2 and all(v == v and v > 0 for v in d["losses"]) and d["dino_package"].startswith("namespace:")
