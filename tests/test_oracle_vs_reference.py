"""Build container only (needs /root/reference): the oracle against the UNMODIFIED reference executed live, on fresh
random inputs (not the committed fixtures) -- this is what pins the oracle ("parity pinned by running the reference")."""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref():
    import ref_import
    warnings.simplefilter("ignore")
    r = ref_import.load_reference()
    ref_import.ensure_gloo_group()
    return r


def test_label_cluster_live(ref):
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    lab = ref.dbscan.label_cluster()
    for m in S.random_masks(12, seed=2024).numpy():
        z = lab(m)
        assert np.array_equal(O.label_cluster(m)[0], z)


def test_full_step_live(ref):
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    E, K, B = 192, 2048, 5
    student = ref.dv.ABIDINOModel(ref.vits.vit_tiny(patch_size=4, drop_path_rate=0.0), ref.seg.SegHead(in_channels=E),
                                  ref.vits.DINOHead(E, K, norm_last_layer=False))
    teacher = ref.dv.ABIDINOModel(ref.vits.vit_tiny(patch_size=4), None, ref.vits.DINOHead(E, K))
    ssd = S.fill_state_dict({k: v.shape for k, v in student.state_dict().items()}, 21, 0.05)
    tsd = S.fill_state_dict({k: v.shape for k, v in teacher.state_dict().items()}, 22, 0.05)
    student.load_state_dict(ssd); teacher.load_state_dict(tsd)
    x, masks, metrics = S.make_batch(B, seed=99)
    crit = ref.loss.DINOLoss(K, 2, 0.04, 0.04, 0, 10)
    so = student(x, metrics, masks, 0, clusters=None)
    to = teacher(x, metrics, None, None, clusters=so["zero"], index=so["index"])
    ag = F.affine_grid(metrics[:, :2, :], size=(B, 1, 32, 128))
    so["gt"] = [masks, (F.grid_sample(masks.unsqueeze(1), ag) > 0.1).float().squeeze()]
    loss = crit(so, to, 0)
    L, parts = O.pretrain_loss(ssd, tsd, "vit_tiny", x, metrics, masks, torch.zeros(1, K), 0, 0.04)
    assert abs(loss.item() - L.item()) < 1e-5 * abs(loss.item())
    assert torch.equal(parts["student"]["zero"], so["zero"])
    assert (parts["student"]["instances_view"] - so["instances_view"]).abs().max() < 1e-5
    assert (parts["center"] - crit.center).abs().max() < 1e-7


def test_dropin_state_dict_equals_reference(ref):
    from Dino.model.dino_vision import ABIDINOModel
    from Dino.modules import vision_transformer as vits
    from Dino.modules.segmentor import SegHead
    for arch, E in (("vit_tiny", 192), ("vit_small", 384), ("vit_base", 512)):
        r = ref.dv.ABIDINOModel(getattr(ref.vits, arch)(patch_size=4, drop_path_rate=0.1), ref.seg.SegHead(in_channels=E),
                                ref.vits.DINOHead(E, 1024, norm_last_layer=False))
        m = ABIDINOModel(vits.__dict__[arch](patch_size=4, drop_path_rate=0.1), SegHead(in_channels=E),
                         vits.DINOHead(E, 1024, norm_last_layer=False))
        assert [(k, tuple(v.shape)) for k, v in r.state_dict().items()] == [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
        assert [p.requires_grad for p in r.parameters()] == [p.requires_grad for p in m.parameters()]
