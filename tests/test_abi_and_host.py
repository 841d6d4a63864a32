"""CPU: the C-ABI library loads and exports every symbol include/ccd_b200.h declares (no compute calls without a GPU),
argument validation returns error codes, the drop-in surface refuses to run on the CPU, host-side helpers."""
import ctypes

import numpy as np
import pytest
import torch


def test_library_exports_every_declared_symbol():
    from ccd_b200 import lib
    L = lib.load()
    syms = lib.declared_symbols()
    assert len(syms) >= 47
    for s in syms:
        assert hasattr(L, s), s
    assert L.ccd_abi_version() == 2


def test_argument_validation_without_gpu():
    from ccd_b200 import ops
    ops._bind()
    f = ops._FN
    assert f["ccd_set_option"](99, 0) == -1 and f["ccd_set_option"](0, 1) == 0
    assert f["ccd_gemm_bf16"](None, None, 128, 128, 64, 0, 0, 0, None, None, None, None, None, 0, 1, None) == -1
    assert f["ccd_gemm_bf16"](1, 1, 128, 100, 64, 0, 0, 0, None, 1, None, None, None, 0, 1, None) == -1     # N % 8
    assert f["ccd_mhsa_fwd"](None, None, None, 1, 3, 0, None) == -1
    assert f["ccd_layernorm_fwd"](1, 1, 1, 1, None, 10, 1024, 1e-6, None) == -1                               # E > 512
    assert f["ccd_ccl_label"](1, 7, 1, None, None, 4, None) == -1                                              # bad mode
    assert f["ccd_dino_ce_fwd"](1, 1, 1, 0.1, 0.04, 1, 1, 1, 4, 1001, None) == -1                             # K % 4
    # recognition decoder entry points
    assert f["ccd_dec_attn_fwd"](None, 512, None, 512, None, 512, None, 512, None, None, 92, 4, 8, 25, 25, 0.0, 0, None) == -1
    assert f["ccd_dec_attn_fwd"](16, 512, 16, 512, 16, 512, 16, 512, None, None, 92, 4, 8, 33, 256, 0.0, 0, None) == -1     # tq > 32
    assert f["ccd_dec_attn_fwd"](16, 512, 16, 512, 16, 512, 16, 512, None, None, 92, 4, 8, 25, 257, 0.0, 0, None) == -1     # tk > 256
    assert f["ccd_dec_attn_fwd"](16, 512, 16, 512, 16, 512, 16, 512, None, 16, 92, 4, 8, 25, 256, 0.0, 0, None) == -1      # mask needs tk == tq
    assert f["ccd_dec_attn_fwd"](16, 516, 16, 512, 16, 512, 16, 512, None, None, 92, 4, 8, 25, 25, 0.0, 0, None) == -1      # ld % 8
    assert f["ccd_dec_attn_fwd"](16, 512, 16, 512, 16, 512, 16, 512, None, None, 92, 4, 8, 25, 25, 1.0, 0, None) == -1      # p_drop < 1
    assert f["ccd_tf_ce"](16, 90, 92, 16, 4, 25, 92, 16, 16, None) == -1                                       # ld < classes
    assert f["ccd_tf_ce"](16, 96, 92, 16, 4, 1, 92, 16, 16, None) == -1                                        # T must exceed 1
    assert f["ccd_dropout"](16, 0, None, 16, 0, 0, 0.1, 0, None) == -1 and f["ccd_dropout"](16, 0, None, 16, 0, 10, 1.5, 0, None) == -1
    assert f["ccd_set_dec_attn_variant"](1) == 0 and f["ccd_set_seg_cls_variant"](1) == 0


def test_no_cpu_fallback():
    from Dino.model.dino_vision import ABIDINOModel
    from Dino.modules import vision_transformer as vits
    from Dino.modules.segmentor import SegHead
    from ccd_b200 import ops
    m = ABIDINOModel(vits.vit_tiny(patch_size=4), SegHead(in_channels=192), vits.DINOHead(192, 256))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 3, 32, 128), torch.eye(3).repeat(2, 1, 1), torch.zeros(2, 32, 128), 0)
    with pytest.raises(RuntimeError):
        ops.layernorm_fwd(torch.zeros(4, 192), torch.ones(192), torch.zeros(192))
    with pytest.raises(NotImplementedError):
        vits.VisionTransformer(patch_size=16, embed_dim=768, num_heads=12, qkv_bias=True)
    # recognition path: same rule
    from Dino.decoder.nrtr_decoder import NRTRDecoder
    from Dino.model.dino_vision import DINO_Finetune
    from ccd_b200 import synthetic as S
    ft = DINO_Finetune(S.finetune_config("vit_tiny"))
    with pytest.raises(RuntimeError):
        ft(torch.zeros(2, 3, 32, 128), S.make_targets(2), return_loss=True)
    with pytest.raises(NotImplementedError):
        NRTRDecoder(d_model=256, d_embedding=256)


def test_param_groups_clip_and_schedules():
    import ccd_oracle as O
    from ccd_b200 import train_utils as U
    lin = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
    g = U.get_params_groups(lin)
    assert len(g[0]["params"]) == 1 and len(g[1]["params"]) == 3 and g[1]["weight_decay"] == 0.
    for p in lin.parameters():
        p.grad = torch.full_like(p, 5.0)
    with pytest.raises(RuntimeError):          # CUDA only, like every other op
        U.clip_gradients(lin, 3.0)
    assert np.allclose(U.cosine_iter_scheduler(1.0, 0.1, 50, 5), O.cosine_iter_schedule(1.0, 0.1, 50, 5))
    assert U.has_batchnorms(torch.nn.Sequential(torch.nn.BatchNorm2d(3))) and not U.has_batchnorms(lin)


def test_wgrad_split_plan_and_chunk_table():
    from ccd_b200 import ops
    assert ops.wgrad_splits(65536, 256, 4352) == 1
    s = ops.wgrad_splits(384, 1536, 131072)
    assert 1 < s <= 2048 and s * 18 >= 296 and ops.gemm_bn(1152) == 192 and ops.gemm_bn(1536) == 256 and ops.gemm_bn(200) == 128
    tab = ops.ChunkTable()
    a = [torch.zeros(70000), torch.zeros(5)]
    b = [torch.zeros(70000, dtype=torch.bfloat16), torch.zeros(5, dtype=torch.bfloat16)]
    t, n = tab.get(a, b, 2)
    assert n == 3 and t.shape == (3, 3) and int(t[1, 2]) == 70000 - 65536
    assert int(t[1, 0]) == a[0].data_ptr() + 4 * 65536 and int(t[1, 1]) == b[0].data_ptr() + 2 * 65536
    t2, _ = tab.get(a, b, 2)
    assert t2 is t


def test_synthetic_batch_contract():
    from ccd_b200 import synthetic as S
    import ccd_oracle as O
    x, masks, metrics = S.make_batch(16, seed=1)
    assert x.shape == (16, 3, 3, 32, 128) and masks.shape == (16, 32, 128) and metrics.shape == (16, 3, 3)
    assert torch.equal(metrics[0], torch.eye(3)) and torch.equal(metrics[3], torch.eye(3))
    rows = 0
    for b in range(16):
        n = int(O.label_cluster(masks[b].numpy())[1].max())
        assert n == 4 + b % 8
        rows += n + 1
    assert 2 * rows == 17 * 16            # 2R = 17 B
