"""Host-side helpers that the reference's train.py calls on `Dino.modules.utils` (train.py:47-48,96,131,137,144-171,186-216,
249-250,276-278): process-group setup, seeding, metric logging, checkpoint resume / save, LARS.

These are caller-side glue, not hot-path work: they exist so that the reference's `train()` runs
UNMODIFIED on top of the drop-in package (tests/test_reference_train_gpu.py executes it).  Behaviour follows the reference
functions named in each docstring (Dino/modules/utils.py); the code is written for this repository.
"""
import argparse
import builtins
import collections
import datetime
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist


# ---- process group ---------------------------------------------------------------------------------------------------------
def is_dist_avail_and_initialized():                     # utils.py:434-439
    return dist.is_available() and dist.is_initialized()


def get_world_size():                                    # utils.py:442-445
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():                                          # utils.py:448-451
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():                                   # utils.py:454-455
    return get_rank() == 0


def save_on_master(*args, **kwargs):                     # utils.py:458-460
    if is_main_process():
        torch.save(*args, **kwargs)


def setup_for_distributed(is_master):
    """utils.py:463-475: print() becomes a no-op on non-master ranks unless called with force=True."""
    plain = builtins.print
    if getattr(plain, "_ccd_rank_filter", False):        # idempotent: never stack filters
        plain = plain._ccd_plain

    def rank_print(*args, **kwargs):
        force = kwargs.pop("force", False)
        if is_master or force:
            plain(*args, **kwargs)

    rank_print._ccd_rank_filter, rank_print._ccd_plain = True, plain
    builtins.print = rank_print


def init_distributed_mode(args):
    """utils.py:478-510.  Fills args.rank / args.world_size / args.gpu from the launcher's environment (torchrun), SLURM or
    the single-GPU default, creates the NCCL process group with args.dist_url, selects the device, barriers, and silences
    print() on non-master ranks."""
    env = os.environ
    if "RANK" in env and "WORLD_SIZE" in env:
        args.rank, args.world_size, args.gpu = int(env["RANK"]), int(env["WORLD_SIZE"]), int(env["LOCAL_RANK"])
    elif "SLURM_PROCID" in env:
        args.rank = int(env["SLURM_PROCID"])
        args.gpu = args.rank % torch.cuda.device_count()
        # the reference forgets world_size on this branch; take SLURM's
        args.world_size = int(env.get("SLURM_NTASKS", getattr(args, "world_size", None) or 1))
    elif torch.cuda.is_available():
        print("Will run the code on one GPU.")
        args.rank, args.gpu, args.world_size = 0, 0, 1
        env["MASTER_ADDR"], env["MASTER_PORT"] = "127.0.0.1", "29501"
    else:
        print("Does not support training without GPU.")
        sys.exit(1)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", init_method=args.dist_url, world_size=args.world_size, rank=args.rank)
    torch.cuda.set_device(args.gpu)
    print(f"| distributed init (rank {args.rank}): {args.dist_url}", flush=True)
    dist.barrier()
    setup_for_distributed(args.rank == 0)


def fix_random_seeds(seed=31):                           # utils.py:226-232
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)


def bool_flag(s):                                        # utils.py:212-223
    v = s.lower()
    if v in ("off", "false", "0"):
        return False
    if v in ("on", "true", "1"):
        return True
    raise argparse.ArgumentTypeError("invalid value for a boolean flag")


# ---- checkpoints -----------------------------------------------------------------------------------------------------------
def restart_from_checkpoint(ckp_path, run_variables=None, **kwargs):
    """utils.py:152-184: for every keyword (student=, teacher=, optimizer=, fp16_scaler=, dino_loss=) present in the file,
    `obj.load_state_dict(ckpt[key], strict=False)` (falling back to the one-argument form for optimizers / scalers), then
    copy the listed run variables (epoch, iteration) out of the file.  A missing file is not an error."""
    if not os.path.isfile(ckp_path):
        return
    print(f"Found checkpoint at {ckp_path}")
    ckpt = torch.load(ckp_path, map_location="cpu", weights_only=False)
    for key, obj in kwargs.items():
        if obj is None or key not in ckpt:
            print(f"=> key '{key}' not found in checkpoint: '{ckp_path}'")
            continue
        try:
            msg = obj.load_state_dict(ckpt[key], strict=False)
            print(f"=> loaded '{key}' from checkpoint '{ckp_path}' with msg {msg}")
        except TypeError:                                # optimizers / GradScaler take no `strict`
            try:
                obj.load_state_dict(ckpt[key])
                print(f"=> loaded '{key}' from checkpoint: '{ckp_path}'")
            except ValueError:
                print(f"=> failed to load '{key}' from checkpoint: '{ckp_path}'")
    for name in (run_variables or {}):
        if name in ckpt:
            run_variables[name] = ckpt[name]


# ---- metric logging --------------------------------------------------------------------------------------------------------
class SmoothedValue:
    """utils.py:235-294: a sliding window (median / avg / max / last value) plus the global average of a scalar series."""

    def __init__(self, window_size=20, fmt=None):
        self.deque = collections.deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.6f} ({global_avg:.6f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        """Sums count / total over the ranks (not the window), as the reference does."""
        if not is_dist_avail_and_initialized():
            return
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device="cuda")
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), float(t[1].item())

    median = property(lambda self: torch.tensor(list(self.deque)).median().item())
    avg = property(lambda self: torch.tensor(list(self.deque), dtype=torch.float32).mean().item())
    global_avg = property(lambda self: self.total / self.count)
    max = property(lambda self: max(self.deque))
    value = property(lambda self: self.deque[-1])

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class MetricLogger:
    """utils.py:324-411: named SmoothedValue meters + `log_every`, the generator train.py:187 wraps the data loader in."""

    def __init__(self, delimiter="\t"):
        self.meters = collections.defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getattr__(self, attr):
        meters = self.__dict__.get("meters", {})
        if attr in meters:
            return meters[attr]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{attr}'")

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        header = header or ""
        total = len(iterable)
        width = len(str(total))
        iter_time, data_time = SmoothedValue(fmt="{avg:.6f}"), SmoothedValue(fmt="{avg:.6f}")
        start = end = time.time()
        cuda = torch.cuda.is_available()
        for i, obj in enumerate(iterable):
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if i % print_freq == 0 or i == total - 1:
                eta = datetime.timedelta(seconds=int(iter_time.global_avg * (total - i)))
                fields = [header, f"[{i:{width}d}/{total}]", f"eta: {eta}", str(self), f"time: {iter_time}", f"data: {data_time}"]
                if cuda:
                    fields.append(f"max mem: {torch.cuda.max_memory_allocated() / 2 ** 20:.0f}")
                print(self.delimiter.join(fields))
            end = time.time()
        spent = time.time() - start
        print(f"{header} Total time: {datetime.timedelta(seconds=int(spent))} ({spent / max(1, total):.6f} s / it)")


# ---- LARS (train.py:136-137, `optimizer: lars`) ------------------------------------------------------------------------------
class LARS(torch.optim.Optimizer):
    """utils.py:564-602: momentum SGD whose update of every >1-D parameter is weight-decayed and rescaled by
    eta * ||p|| / ||update|| (trust ratio; 1 when either norm is zero); 1-D parameters get neither."""

    def __init__(self, params, lr=0, weight_decay=0, momentum=0.9, eta=0.001, weight_decay_filter=None,
                 lars_adaptation_filter=None):
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, momentum=momentum, eta=eta,
                                      weight_decay_filter=weight_decay_filter, lars_adaptation_filter=lars_adaptation_filter))

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                upd = p.grad
                if p.ndim != 1:
                    upd = upd.add(p, alpha=group["weight_decay"])
                    pn, un = torch.norm(p), torch.norm(upd)
                    trust = torch.where((pn > 0.) & (un > 0.), group["eta"] * pn / un, torch.ones_like(pn))
                    upd = upd.mul(trust)
                state = self.state[p]
                if "mu" not in state:
                    state["mu"] = torch.zeros_like(p)
                mu = state["mu"].mul_(group["momentum"]).add_(upd)
                p.add_(mu, alpha=-group["lr"])
