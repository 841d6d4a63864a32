"""checkpoints written by the reference's train.py (:197-211: {'student', 'teacher', 'optimizer', 'epoch',
'iteration', 'dino_loss'}, DDP-prefixed keys) load into the drop-in modules and vice versa, the fine-tuning model picks the
pretrained teacher backbone out of such a file exactly as train_finetune.py:193-200 does, and the optimizer state dict
round-trips with torch.optim.AdamW.  Host logic only (modules are constructed on the CPU; no kernel runs)."""
import io
import warnings

import pytest
import torch


def _pair(arch="vit_tiny", E=192, K=512):
    from Dino.loss.Dino_loss import DINOLoss
    from Dino.model.dino_vision import ABIDINOModel
    from Dino.modules import vision_transformer as vits
    from Dino.modules.segmentor import SegHead
    student = ABIDINOModel(vits.__dict__[arch](patch_size=4, drop_path_rate=0.1), SegHead(in_channels=E), vits.DINOHead(E, K, norm_last_layer=False))
    teacher = ABIDINOModel(vits.__dict__[arch](patch_size=4), None, vits.DINOHead(E, K))
    return student, teacher, DINOLoss(K, 2, 0.04, 0.04, 0, 10)


def _ddp_keys(sd):
    return {"module." + k: v for k, v in sd.items()}           # what DistributedDataParallel.state_dict() produces


def _roundtrip(obj):
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return torch.load(buf, weights_only=False)


def test_train_py_checkpoint_format_roundtrip():
    from ccd_b200 import synthetic as S
    from ccd_b200.optim import AdamW
    from ccd_b200.train_utils import get_params_groups
    student, teacher, loss = _pair()
    student.load_state_dict(S.fill_state_dict({k: v.shape for k, v in student.state_dict().items()}, 1, 0.05))
    loss.center.copy_(torch.randn_like(loss.center))
    opt = AdamW(get_params_groups(student), lr=1e-3)
    ck = _roundtrip({"student": _ddp_keys(student.state_dict()), "teacher": _ddp_keys(teacher.state_dict()),
                     "optimizer": opt.state_dict(), "epoch": 3, "iteration": 1234, "dino_loss": loss.state_dict()})
    s2, t2, l2 = _pair()
    s2.load_state_dict({k[len("module."):]: v for k, v in ck["student"].items()})               # strict
    t2.load_state_dict({k[len("module."):]: v for k, v in ck["teacher"].items()})
    l2.load_state_dict(ck["dino_loss"])
    assert all(torch.equal(a, b) for a, b in zip(student.state_dict().values(), s2.state_dict().values()))
    assert torch.equal(l2.center, loss.center) and list(ck["dino_loss"].keys()) == ["center"]


def test_finetune_reads_pretrained_teacher_backbone():
    """train_finetune.py:193-200: dd[name] = ckpt['teacher'][name] for every key of DataParallel(DINO_Finetune).state_dict()
    that exists in the pretraining checkpoint (the 'module.backbone.*' entries); everything else keeps its initialisation."""
    from Dino.model.dino_vision import DINO_Finetune
    from ccd_b200 import synthetic as S
    _, teacher, _ = _pair("vit_small", 384)
    teacher.load_state_dict(S.fill_state_dict({k: v.shape for k, v in teacher.state_dict().items()}, 7, 0.05))
    ckpt = _roundtrip({"teacher": _ddp_keys(teacher.state_dict())})
    model = torch.nn.DataParallel(DINO_Finetune(S.finetune_config("vit_small")))
    dd = model.state_dict()
    taken = 0
    for name in dd.keys():
        try:
            dd[name] = ckpt["teacher"][name]
            taken += 1
        except KeyError:
            pass
    model.load_state_dict(dd)
    bb = {k: v for k, v in teacher.state_dict().items() if k.startswith("backbone.")}
    assert taken == len(bb) > 150
    got = model.module.state_dict()
    assert all(torch.equal(got[k], v) for k, v in bb.items())


def test_optimizer_state_dict_interchange_with_torch_adamw():
    from ccd_b200.optim import AdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11))]
    groups = lambda: [{"params": [ps[0]]}, {"params": [ps[1]], "weight_decay": 0.0}]
    ref = torch.optim.AdamW(groups(), lr=1e-3, weight_decay=0.04)
    for p in ps:
        p.grad = torch.randn_like(p)
    ref.step()
    ref.step()
    mine = AdamW(groups(), lr=1e-3, weight_decay=0.04)
    mine.load_state_dict(_roundtrip(ref.state_dict()))
    assert all(isinstance(st["step"], int) and st["step"] == 2 for st in mine.state.values())
    back = torch.optim.AdamW(groups(), lr=1e-3, weight_decay=0.04)
    back.load_state_dict(_roundtrip(mine.state_dict()))
    a, b = ref.state_dict(), back.state_dict()
    assert a["param_groups"] == b["param_groups"]
    for k in a["state"]:
        assert float(a["state"][k]["step"]) == float(b["state"][k]["step"])
        assert torch.equal(a["state"][k]["exp_avg"], b["state"][k]["exp_avg"]) and torch.equal(a["state"][k]["exp_avg_sq"], b["state"][k]["exp_avg_sq"])


@pytest.mark.needs_reference
def test_checkpoints_interchange_with_the_live_reference():
    import ref_import
    from ccd_b200 import synthetic as S
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference()
    student, teacher, loss = _pair()
    rs = ref.dv.ABIDINOModel(ref.vits.vit_tiny(patch_size=4, drop_path_rate=0.1), ref.seg.SegHead(in_channels=192),
                             ref.vits.DINOHead(192, 512, norm_last_layer=False))
    rl = ref.loss.DINOLoss(512, 2, 0.04, 0.04, 0, 10)
    sd = S.fill_state_dict({k: v.shape for k, v in rs.state_dict().items()}, 3, 0.05)
    rs.load_state_dict(sd)
    student.load_state_dict(_roundtrip(rs.state_dict()))        # reference checkpoint -> drop-in (strict)
    rs.load_state_dict(_roundtrip(student.state_dict()))        # drop-in checkpoint -> reference (strict)
    loss.load_state_dict(_roundtrip(rl.state_dict()))
    rl.load_state_dict(_roundtrip(loss.state_dict()))
    # fine-tuning model: reference-written file into the drop-in and back
    rf = ref.dv.DINO_Finetune(S.finetune_config("vit_tiny"))
    from Dino.model.dino_vision import DINO_Finetune
    mf = DINO_Finetune(S.finetune_config("vit_tiny"))
    mf.load_state_dict(_roundtrip(rf.state_dict()))
    rf.load_state_dict(_roundtrip(mf.state_dict()))
