"""CPU: the host-side helpers `Dino.modules.utils` exports for the reference's train.py (ccd_b200/host_utils.py, train_utils.py)
against the reference's own functions (Dino/modules/utils.py), executed live where the reference tree is present."""
import io
import contextlib
import os

import numpy as np
import pytest
import torch

from ccd_b200 import host_utils as H
from ccd_b200 import train_utils as T


@pytest.fixture(scope="module")
def ref_utils():
    import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    return ref_import.load_reference().utils


def test_bool_flag_and_seeds():
    assert H.bool_flag("On") is True and H.bool_flag("0") is False
    with pytest.raises(Exception):
        H.bool_flag("maybe")
    H.fix_random_seeds(7)
    a = torch.rand(3), np.random.rand()
    H.fix_random_seeds(7)
    b = torch.rand(3), np.random.rand()
    assert torch.equal(a[0], b[0]) and a[1] == b[1]


def test_smoothed_value_and_metric_logger_match_reference(ref_utils):
    vals = [0.5, 2.0, 1.25, 7.0, 3.5, 3.5, 0.1]
    for cls_a, cls_b in ((H.SmoothedValue, ref_utils.SmoothedValue),):
        a, b = cls_a(window_size=4), cls_b(window_size=4)
        for v in vals:
            a.update(v); b.update(v)
        assert (a.median, a.avg, a.global_avg, a.max, a.value) == (b.median, b.avg, b.global_avg, b.max, b.value)
        assert str(a) == str(b)
    la, lb = H.MetricLogger(delimiter="  "), ref_utils.MetricLogger(delimiter="  ")
    for lg in (la, lb):
        lg.update(loss=torch.tensor(1.5), lr=0.1)
        lg.update(loss=2.5, lr=0.2)
    assert str(la) == str(lb) and la.loss.global_avg == lb.loss.global_avg
    with pytest.raises(AttributeError):
        la.nothing
    # log_every yields every item and prints the same fields (timings differ)
    out_a, out_b = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out_a):
        got_a = list(la.log_every(list(range(5)), 2, "Epoch: [0/1]"))
    with contextlib.redirect_stdout(out_b):
        got_b = list(lb.log_every(list(range(5)), 2, "Epoch: [0/1]"))
    assert got_a == got_b == list(range(5))
    assert out_a.getvalue().count("\n") == out_b.getvalue().count("\n")
    assert [l.split("eta")[0] for l in out_a.getvalue().splitlines()[:-1]] == [l.split("eta")[0] for l in out_b.getvalue().splitlines()[:-1]]


def test_lars_matches_reference(ref_utils):
    torch.manual_seed(0)
    w0 = [torch.randn(6, 5), torch.randn(7), torch.zeros(3, 3)]
    grads = [[torch.randn_like(w) for w in w0] for _ in range(4)]
    grads[1][0].zero_()                                         # a zero update norm -> trust ratio 1
    ps_a = [torch.nn.Parameter(w.clone()) for w in w0]
    ps_b = [torch.nn.Parameter(w.clone()) for w in w0]
    oa = H.LARS(ps_a, lr=0.1, weight_decay=1e-2, momentum=0.9)
    ob = ref_utils.LARS(ps_b, lr=0.1, weight_decay=1e-2, momentum=0.9)
    for g in grads:
        for p, q, gg in zip(ps_a, ps_b, g):
            p.grad, q.grad = gg.clone(), gg.clone()
        oa.step(); ob.step()
    for p, q in zip(ps_a, ps_b):
        assert torch.allclose(p, q, atol=1e-7, rtol=1e-6)


def test_schedules_param_groups_and_cancel_match_reference(ref_utils):
    for args in ((1e-3, 1e-6, 100, 10, 0), (0.04, 0.4, 57, 0, 0), (0.9995, 1, 33, 0, 0)):
        assert np.allclose(T.cosine_iter_scheduler(*args), ref_utils.cosine_iter_scheduler(*args), rtol=0, atol=1e-15)
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3), torch.nn.Linear(3, 2, bias=False))
    m[2].weight.requires_grad = False
    ga, gb = T.get_params_groups(m), ref_utils.get_params_groups(m)
    assert [[id(p) for p in g["params"]] for g in ga] == [[id(p) for p in g["params"]] for g in gb]
    assert ga[1]["weight_decay"] == gb[1]["weight_decay"] == 0.
    assert T.has_batchnorms(torch.nn.Sequential(torch.nn.BatchNorm2d(3))) and not T.has_batchnorms(m)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(2, 2)
            self.last_layer = torch.nn.Linear(2, 2)
    for fn in (T.cancel_gradients_last_layer, ref_utils.cancel_gradients_last_layer):
        n = Net()
        for p in n.parameters():
            p.grad = torch.ones_like(p)
        fn(0, n, 1)
        assert n.last_layer.weight.grad is None and n.body.weight.grad is not None
        for p in n.parameters():
            p.grad = torch.ones_like(p)
        fn(1, n, 1)
        assert n.last_layer.weight.grad is not None


def test_restart_from_checkpoint_and_save_on_master(tmp_path):
    net = torch.nn.Linear(3, 2)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    net(torch.ones(1, 3)).sum().backward(); opt.step()
    path = os.path.join(tmp_path, "checkpoint.pth")
    H.save_on_master({"student": net.state_dict(), "optimizer": opt.state_dict(), "epoch": 3, "iteration": 77}, path)
    net2 = torch.nn.Linear(3, 2)
    opt2 = torch.optim.SGD(net2.parameters(), lr=0.1, momentum=0.9)
    run = {"epoch": 0, "iteration": 0}
    with contextlib.redirect_stdout(io.StringIO()) as out:
        H.restart_from_checkpoint(path, run_variables=run, student=net2, teacher=None, optimizer=opt2, dino_loss=torch.nn.Linear(1, 1))
        H.restart_from_checkpoint(os.path.join(tmp_path, "missing.pth"), run_variables=run, student=net2)
    assert run == {"epoch": 3, "iteration": 77}
    assert torch.equal(net2.weight, net.weight) and opt2.state_dict()["state"].keys() == opt.state_dict()["state"].keys()
    assert "key 'dino_loss' not found" in out.getvalue() and "key 'teacher' not found" in out.getvalue()
    assert H.get_world_size() == 1 and H.get_rank() == 0 and H.is_main_process()


def test_setup_for_distributed_is_idempotent_and_restorable():
    import builtins
    orig = builtins.print
    try:
        H.setup_for_distributed(False)
        H.setup_for_distributed(False)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            print("hidden")
            print("shown", force=True)
        assert buf.getvalue() == "shown\n"
        assert builtins.print._ccd_plain is orig
    finally:
        builtins.print = orig
