"""CPU: the oracle restatement against the fixtures the UNMODIFIED reference produced (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import ccd_oracle as O
from ccd_b200 import synthetic as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
import sys
sys.path.insert(0, GOLD)
from cases import CASES, EPOCH, SEG_KEEP, col_stride  # noqa: E402


def shapes(arch, E, K, student=True):
    """Parameter / buffer names+shapes of the reference modules, from the drop-in (asserted equal to the reference's
    in test_dropin_surface.py)."""
    from ccd_b200.encoder import VisionTransformer
    from ccd_b200.head import DINOHead
    from ccd_b200.model import ABIDINOModel
    from ccd_b200.segmentor import SegHead
    heads = {192: 3, 384: 6, 512: 8}[E]
    bb = VisionTransformer(patch_size=4, embed_dim=E, depth=12, num_heads=heads, mlp_ratio=4, qkv_bias=True)
    m = ABIDINOModel(bb, SegHead(in_channels=E) if student else None, DINOHead(E, K))
    return {k: v.shape for k, v in m.state_dict().items()}


def test_ccl_matches_reference_label_cluster():
    z = np.load(os.path.join(GOLD, "ccl_cases.npz"))
    for i, m in enumerate(z["masks"]):
        onehot, compact = O.label_cluster(m.astype(np.float32))
        assert np.array_equal(compact, z["compact"][i]), i
        assert onehot.sum(0).max() <= 1


def test_ccl_raster_labels_match_scipy():
    from scipy import ndimage
    for m in S.random_masks(10, seed=77).numpy():
        want = ndimage.label(m != 0, structure=np.ones((3, 3)))[0]
        assert np.array_equal(O.ccl_labels_raster(m), want)


def test_pos_embed_is_linear_operator():
    p = torch.randn(1, 256, 24, dtype=torch.float64)
    W = O.pos_resample_matrix()
    assert (O.pos_embed_resampled(p)[0] - W @ p[0]).abs().max() < 1e-10
    # parity trap: resampling with size=(8,32) instead of the scale factor is a DIFFERENT operator
    q = torch.nn.functional.interpolate(p.reshape(1, 16, 16, 24).permute(0, 3, 1, 2), size=(8, 32), mode="bicubic")
    assert (q.permute(0, 2, 3, 1).reshape(256, 24) - W @ p[0]).abs().max() > 1e-3


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_step_matches_reference_golden(name):
    arch, E, B, K, sseed, tseed, std, norm_last = CASES[name]
    epoch, cs = EPOCH.get(name, 0), col_stride(name)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ssd = S.fill_state_dict(shapes(arch, E, K, True), sseed, std)
    tsd = S.fill_state_dict(shapes(arch, E, K, False), tseed, std)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in ssd.items()}
    x, masks, metrics = S.make_batch(B, seed=1234)
    center0 = 0.01 * torch.randn(1, K, generator=torch.Generator().manual_seed(5))
    # epoch >= 30: the mask is a threshold of the segmentation logits -- compare the logits, then label the reference's
    # (a last-bit difference on a pixel at the threshold would otherwise change the component structure)
    own = None
    if epoch >= 30:
        own = torch.tensor(g["seg_logits"])
        with torch.no_grad():
            seg = O.student_forward(ssd, arch, x, metrics, masks, 0)["mask"]
        assert (seg - own).abs().max() < 1e-4
        assert ((seg[:, 1] > seg[:, 0]) != (own[:, 1] > own[:, 0])).float().mean() < 1e-4
    L, parts = O.pretrain_loss(sd, tsd, arch, x, metrics, masks, center0, epoch, 0.04, self_mask_logits=own)
    L.backward()
    assert abs(L.item() - g["loss"]) < 2e-5 * abs(g["loss"])
    assert abs(parts["mask_loss"].item() - g["mask_loss"]) < 1e-5
    assert abs(parts["Dino_loss"].item() - g["dino_loss"]) < 2e-5 * g["dino_loss"]
    assert np.abs(parts["student"]["instances_view"].detach().numpy()[:, ::cs] - g["student_logits"]).max() < 1e-5
    assert np.abs(parts["teacher"]["instances_view"].detach().numpy()[:, ::cs] - g["teacher_logits"]).max() < 1e-5
    assert np.array_equal(parts["student"]["index"].numpy(), g["new_index"])
    assert np.array_equal(parts["gt"][B:].numpy().astype(np.uint8), g["gt_warped"])
    assert np.abs(parts["center"].numpy()[:, ::cs] - g["center_after"]).max() < 1e-7
    dense = parts["student"]["zero"].numpy()
    assert np.array_equal((dense[:B] * np.arange(1, 27)[None, :, None, None]).sum(1).astype(np.uint8), g["clusters_compact"][:B])
    for n, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        gr = sd[str(n)].grad
        assert gr is not None, n
        if norm > 1e-6:
            assert abs(gr.norm().item() - norm) < 1e-2 * norm + 1e-7, (n, gr.norm().item(), norm)


def test_explicit_forms_of_layernorm_and_gelu():
    x = torch.randn(7, 192, dtype=torch.float64)
    w, b = torch.randn(192, dtype=torch.float64), torch.randn(192, dtype=torch.float64)
    assert (O.layer_norm(x, w, b) - O.layer_norm_explicit(x, w, b)).abs().max() < 1e-12
    assert (O.gelu(x) - O.gelu_explicit(x)).abs().max() < 1e-12


def test_clip_ema_schedule_restatements():
    g = {"a": torch.full((10,), 2.0), "b": torch.full((4,), 0.1)}
    c = O.clip_per_parameter(g, 3.0)
    assert abs(c["a"].norm().item() - 3.0) < 1e-4 and torch.equal(c["b"], g["b"])
    sch = O.cosine_iter_schedule(1.0, 0.1, 100, warmup_iters=10)
    assert len(sch) == 100 and sch[0] == 0 and abs(sch[10] - 1.0) < 1e-12 and sch[-1] > 0.1
    t = {"backbone.w": torch.ones(3), "head.w": torch.ones(3), "other": torch.ones(3)}
    s = {"backbone.w": torch.zeros(3), "head.w": torch.zeros(3), "segmentation.w": torch.zeros(3)}
    out = O.ema_update(dict(t), s, 0.9)
    assert torch.allclose(out["backbone.w"], torch.full((3,), 0.9)) and torch.equal(out["other"], t["other"])
