"""TEST / BASELINE INFRASTRUCTURE -- runs the reference's OWN `train()` (train.py:45-302, unmodified, from /root/reference or
its byte-compiled copy oracle/_ref, see oracle/build_ref.py) for a few iterations on synthetic batches, with either

  --impl reference   the reference's own modules (stock PyTorch-CUDA path = BASELINE.md B1: SyncBN convert, teacher DDP,
                     student DDP(find_unused_parameters=True), per-parameter clip with .item(), torch.optim.AdamW, the
                     `.data.mul_` EMA loop, CPU connected components -- whatever train.py does), or
  --impl dropin      this repository's `Dino` package first on sys.path (CCD_REFERENCE_ROOT pointing at the reference for the
                     parts the drop-in does not replace: Dino/utils, Dino/dataset, configs).

What the harness supplies, and nothing else:
  * stubs for third-party imports MISSING from this image (fastai, lmdb, imgaug, matplotlib, skimage.measure.label ->
    scipy.ndimage.label): oracle/ref_import.py's list plus the imgaug sub-modules the dataset code imports;
  * a YAML config derived from the reference's Dino/configs/CCD_pretrain_ViT_small.yaml with the size knobs overridden;
  * `train._get_databaunch` replaced by a synthetic loader (the LMDB datasets do not exist here; BASELINE.md section 3 inputs);
  * optionally an initial `checkpoint.pth` (deterministic weights) that train()'s own restart_from_checkpoint loads, so two
    runs start from identical parameters;
  * `config.writer`: a recorder instead of a TensorBoard SummaryWriter (train.py:281-290 logs mask_loss / Dino_loss to it);
  * --autocast bf16: torch.cuda.amp.autocast is pointed at torch.autocast(bfloat16) for BASELINE.md B1 (ii).
One process per GPU: launch with torch.distributed.run for N > 1 (RANK / LOCAL_RANK / WORLD_SIZE), or plainly for one GPU.
Prints ONE JSON line on rank 0:  {"impl", "losses": [...], "mask_loss": [...], "dino_loss": [...], "iter_ms": [...], ...}
"""
import argparse
import importlib.util
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "dropin"], required=True)
    ap.add_argument("--arch", default="vit_small")
    ap.add_argument("--out_dim", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3, help="iterations excluded from the reported mean time")
    ap.add_argument("--drop_path", type=float, default=0.1)
    ap.add_argument("--autocast", choices=["none", "bf16"], default="none")
    ap.add_argument("--init-seed", type=int, default=-1, help=">= 0: write a deterministic checkpoint.pth for train() to resume from")
    ap.add_argument("--clip-grad", type=float, default=3.0)
    ap.add_argument("--optimizer", default="adamw")
    ap.add_argument("--lr", type=float, default=0.0005)
    ap.add_argument("--momentum-teacher", type=float, default=0.9995)
    ap.add_argument("--fresh-batches", type=int, default=1, help="1: a new synthetic batch per iteration; 0: the same batch")
    ap.add_argument("--dump-state", default="", help="path: save student/teacher/centre after the run (rank 0)")
    ap.add_argument("--cpu-debug", action="store_true", help="harness self-test on a box without a GPU (reference impl only): "
                    "gloo instead of NCCL, .cuda() / set_device / synchronize become no-ops")
    args = ap.parse_args()

    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])

    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
    ref_import = _load_by_path("ref_import", os.path.join(HERE, "ref_import.py"))
    ref_root = ref_import.REFERENCE_ROOT
    if not ref_import.reference_available():
        print(json.dumps({"impl": args.impl, "unavailable": f"no reference at {ref_root} (run oracle/build_ref.py)"}))
        return 0
    ref_import._install_stubs()
    for name in ("imgaug.random", "imgaug.augmenters", "imgaug.augmenters.geometric", "imgaug.augmentables",
                 "imgaug.augmentables.segmaps"):
        if name not in sys.modules:
            ref_import._mod(name)
    sys.modules["imgaug.augmenters.geometric"]._warp_affine_arr = None
    sys.modules["imgaug.augmentables.segmaps"].SegmentationMapsOnImage = object
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    import torchvision  # noqa: F401  (before the reference's namespace package is on the stack, see ref_import)

    if args.impl == "dropin":
        os.environ["CCD_REFERENCE_ROOT"] = ref_root
        sys.path.insert(0, REPO)                 # the regular package Dino/ of this repository wins; it extends its __path__
    sys.path.insert(1 if args.impl == "dropin" else 0, ref_root)
    S = _load_by_path("ccd_synthetic", os.path.join(REPO, "ccd_b200", "synthetic.py"))

    import numpy as np
    import torch
    import yaml

    work = tempfile.mkdtemp(prefix=f"ccd_ref_train_{args.impl}_r{rank}_")
    # Config reads Dino/configs/template.yaml relative to the working directory (Dino/utils/utils.py:208)
    if os.path.isfile(os.path.join(ref_root, "data_files.json")):               # the compiled copy oracle/_ref: unpack its data blob
        build_ref = _load_by_path("build_ref", os.path.join(HERE, "build_ref.py"))
        build_ref.verify(ref_root)
        build_ref.materialize_data(work, ref_root)
    else:                                                                        # a source checkout
        os.symlink(os.path.join(ref_root, "Dino"), os.path.join(work, "Dino"))
    os.chdir(work)
    with open(os.path.join(work, "Dino", "configs", "CCD_pretrain_ViT_small.yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    E = {"vit_tiny": 192, "vit_small": 384, "vit_base": 512}[args.arch]
    cfg.update(arch=args.arch, out_dim=args.out_dim, batch_size_per_gpu=args.batch, drop_path_rate=args.drop_path, warmup_epoch=0,
               clip_grad=args.clip_grad, optimizer=args.optimizer, lr=args.lr, momentum_teacher=args.momentum_teacher, output_dir=os.path.join(work, "saved_models"), seed=0)
    cfg["model"]["seg_channel"] = E
    cfg["training"].update(epochs=1, show_iters=1)
    cfg["global"]["name"] = "run"
    cfg_path = os.path.join(work, "cfg.yaml")
    with open(cfg_path, "w") as f:
        yaml.dump(cfg, f)

    if args.autocast == "bf16":
        torch.cuda.amp.autocast = lambda *a, **k: torch.autocast("cuda", dtype=torch.bfloat16)

    if args.cpu_debug:
        import torch.distributed as _d
        import torch.nn as _nn
        _init = _d.init_process_group
        _d.init_process_group = lambda backend=None, **k: _init(backend="gloo", **k)
        torch.Tensor.cuda = lambda self, *a, **k: self
        _nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.set_device = lambda *a, **k: None
        torch.cuda.synchronize = lambda *a, **k: None
        _ddp = _nn.parallel.DistributedDataParallel.__init__
        _nn.parallel.DistributedDataParallel.__init__ = lambda self, m, device_ids=None, **k: _ddp(self, m, **k)
        _nn.SyncBatchNorm.convert_sync_batchnorm = classmethod(lambda cls, m, process_group=None: m)   # SyncBN is GPU-only
        torch.cuda.amp.autocast = lambda *a, **k: torch.autocast("cpu", enabled=False)

    import train as T                                   # the reference's train.py, imported as a module (no __main__ block)
    config = T.Config(cfg_path)
    os.makedirs(os.path.join(config.output_dir, config.global_name), exist_ok=True)

    class Sampler:
        def set_epoch(self, e):
            pass

    class Loader:
        """Synthetic stand-in for the LMDB DataLoader: BASELINE.md section 3 batches, host tensors (train.py moves them)."""

        def __init__(self):
            self.sampler = Sampler()
            self.stamps = []
            self.batches = [S.make_batch(args.batch, seed=1234 + rank + 1000 * i)
                            for i in range(args.iters if args.fresh_batches else 1)]

        def __len__(self):
            return args.iters

        def __iter__(self):
            # train.py breaks only when iteration > epochs * len(loader): feed exactly `iters` batches
            for i in range(args.iters):
                torch.cuda.synchronize()
                self.stamps.append(time.perf_counter())
                x, m, th = self.batches[i % len(self.batches)]
                yield (x, m, th) if args.cpu_debug else (x.pin_memory(), m.pin_memory(), th.pin_memory())
            torch.cuda.synchronize()
            self.stamps.append(time.perf_counter())

    loader = Loader()
    T._get_databaunch = lambda cfg_: loader

    class Recorder:
        def __init__(self):
            self.scalars = {}

        def add_scalar(self, tag, scalar_value, global_step):
            self.scalars.setdefault(tag, []).append(float(np.asarray(scalar_value)))

    config.writer = Recorder()

    if args.init_seed >= 0 and rank == 0:
        from Dino.model.dino_vision import ABIDINOModel
        from Dino.modules import vision_transformer as vits
        from Dino.modules.segmentor import SegHead
        st = ABIDINOModel(vits.__dict__[args.arch](patch_size=4), SegHead(in_channels=E, mla_channels=128, mlahead_channels=64,
                                                                         num_classes=2),
                          vits.DINOHead(E, args.out_dim, use_bn=False, norm_last_layer=False))
        te = ABIDINOModel(vits.__dict__[args.arch](patch_size=4), None, vits.DINOHead(E, args.out_dim, False))
        ssd = S.fill_state_dict({k: v.shape for k, v in st.state_dict().items()}, args.init_seed + 1, 0.05)
        tsd = S.fill_state_dict({k: v.shape for k, v in te.state_dict().items()}, args.init_seed + 2, 0.05)
        torch.save({"student": {"module." + k: v for k, v in ssd.items()}, "teacher": {"module." + k: v for k, v in tsd.items()}},
                   os.path.join(config.output_dir, config.global_name, "checkpoint.pth"))
        del st, te

    captured = {}
    if args.dump_state:
        # observe (never alter) the objects train() builds: DINOLoss.__init__ and DDP.__init__ are wrapped to remember them
        import torch.nn as nn
        from Dino.loss import Dino_loss as DL
        ddp_init, loss_init = nn.parallel.DistributedDataParallel.__init__, DL.DINOLoss.__init__

        def ddp_spy(self, module, *a, **k):
            ddp_init(self, module, *a, **k)
            captured.setdefault("ddp", []).append(self)

        def loss_spy(self, *a, **k):
            loss_init(self, *a, **k)
            captured["loss"] = self

        nn.parallel.DistributedDataParallel.__init__ = ddp_spy
        DL.DINOLoss.__init__ = loss_spy

    t0 = time.perf_counter()
    T.train(config)
    wall = time.perf_counter() - t0
    import builtins
    plain = getattr(builtins.print, "_ccd_plain", None)     # the drop-in's rank filter keeps the original; the reference's does not
    stamps = loader.stamps
    iter_ms = [1e3 * (b - a) for a, b in zip(stamps[:-1], stamps[1:])]
    timed = iter_ms[args.warmup:] if len(iter_ms) > args.warmup else iter_ms
    import torch.distributed as dist
    mean_ms = sum(timed) / max(1, len(timed))
    if dist.is_initialized() and world > 1:
        t = torch.tensor([mean_ms], device="cpu" if args.cpu_debug else "cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mean_ms = float(t.item())
    sc = config.writer.scalars
    out = {"impl": args.impl, "arch": args.arch, "out_dim": args.out_dim, "batch_per_gpu": args.batch, "world": world,
           "iters": args.iters, "warmup": args.warmup, "autocast": args.autocast, "drop_path": args.drop_path,
           "mask_loss": sc.get("metric/mask_loss", []), "dino_loss": sc.get("metric/Dino_loss", []),
           "losses": [a + b for a, b in zip(sc.get("metric/mask_loss", []), sc.get("metric/Dino_loss", []))],
           "lr": sc.get("metric/lr", []), "iter_ms": [round(v, 3) for v in iter_ms], "mean_iter_ms": mean_ms,
           "images_per_s": 1e3 * args.batch * world / mean_ms if mean_ms > 0 else None, "train_wall_s": round(wall, 2),
           "reference_root": ref_root, "dino_package": sys.modules["Dino"].__dict__.get("__file__") or "namespace:" + ref_root}
    if args.dump_state and rank == 0 and captured.get("ddp"):
        student = captured["ddp"][-1]
        teacher = captured["ddp"][0]
        torch.save({"student": {k: v.detach().cpu() for k, v in student.state_dict().items()},
                    "teacher": {k: v.detach().cpu() for k, v in teacher.state_dict().items()},
                    "center": captured["loss"].center.detach().cpu()}, args.dump_state)
    if rank == 0:
        (plain or builtins.print)(json.dumps(out), **({} if plain else {"force": True}))
    if dist.is_initialized():
        dist.destroy_process_group()
    os.chdir(HERE)
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
