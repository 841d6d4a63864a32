"""CPU, world_size 2 over gloo: the multi-rank semantics the path relies on -- the centre update's SUM all-reduce with the
local_rows*world divisor (Dino/loss/Dino_loss.py:133-143) and per-rank synthetic sharding of bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    rows = 6 + 4 * rank                                   # ragged: different R per rank
    zt = torch.randn(rows, 512, generator=g)
    c0 = torch.zeros(1, 512)
    s = zt.sum(0, keepdim=True)
    dist.all_reduce(s)
    c = O.updated_center(c0, zt, world_sum=s, world_size=world)
    # every rank divides by ITS OWN row count: centres differ across ranks (the reference quirk we replicate)
    gathered = [torch.zeros_like(c) for _ in range(world)]
    dist.all_gather(gathered, c)
    x, masks, _ = S.make_batch(4, seed=1234 + rank)
    xs = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(xs, x)
    if rank == 0:
        q.put((float((gathered[0] * 6 - gathered[1] * 10).abs().max()), bool(torch.equal(xs[0], xs[1])),
               float((gathered[0] - s / (6 * world) * 0.1).abs().max())))
    dist.destroy_process_group()


def test_center_allreduce_divisor_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    scaled_diff, same_inputs, err0 = q.get(timeout=10)
    assert scaled_diff < 1e-5          # c_r * rows_r is rank-independent  <=>  divisor = local_rows * world
    assert not same_inputs             # ranks draw different synthetic shards (seed 1234 + rank)
    assert err0 < 1e-6


@pytest.mark.needs_reference
def test_reference_update_center_single_rank_matches_oracle():
    import warnings
    import ref_import
    import ccd_oracle as O
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference()
    ref_import.ensure_gloo_group()
    crit = ref.loss.DINOLoss(256, 2, 0.04, 0.04, 0, 10)
    zt = torch.randn(14, 256)
    want = O.updated_center(crit.center.clone(), zt)
    crit.update_center(zt)
    assert (crit.center - want).abs().max() < 1e-7


def _host_worker(rank, world, port, tmp, q):
    """The host-side helpers train.py calls under N > 1 (Dino.modules.utils -> ccd_b200/host_utils.py) in a 2-rank gloo group."""
    import contextlib
    import io
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root]
    from Dino.modules import utils
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert utils.get_world_size() == world and utils.get_rank() == rank and utils.is_main_process() == (rank == 0)
    path = os.path.join(tmp, f"ckpt_{rank}.pth")
    utils.save_on_master({"epoch": 5, "iteration": 9}, path)          # only rank 0 writes (train.py:204-207)
    utils.setup_for_distributed(rank == 0)                            # print() is silent on the other ranks unless force=True
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        print("log line")
        print("fatal", force=True)
    dist.barrier()
    q.put((rank, os.path.exists(path), buf.getvalue()))
    dist.destroy_process_group()


def test_host_helpers_two_ranks(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = {r: (wrote, out) for r, wrote, out in (q.get(timeout=10), q.get(timeout=10))}
    assert got[0] == (True, "log line\nfatal\n")
    assert got[1] == (False, "fatal\n")
