"""Recognition / fine-tuning path: the oracle restatement against the committed reference outputs
(runs anywhere) and against the UNMODIFIED reference executed live (build container only); drop-in interface checks."""
import os
import warnings

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"finetune_vit_tiny_b4": ("vit_tiny", 4, 5, 0.05), "finetune_vit_small_b3": ("vit_small", 3, 6, 0.04)}


def _case(name):
    from ccd_b200 import synthetic as S
    from ccd_b200.finetune import DINO_Finetune
    arch, n, wseed, std = CASES[name]
    shapes = {k: v.shape for k, v in DINO_Finetune(S.finetune_config(arch)).state_dict().items()}
    sd = S.fill_state_dict(shapes, wseed, std)
    img = torch.randn(n, 3, 32, 128, generator=torch.Generator().manual_seed(100 + n))
    tgt = S.make_targets(n, seed=200 + n)
    return arch, sd, img, tgt, np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    import finetune_oracle as FO
    arch, sd, img, tgt, gold = _case(name)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "position_table" not in k) for k, v in sd.items()}
    loss, logits, _ = FO.finetune_forward_train(sd, arch, img, tgt)
    assert abs(loss.item() - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))
    assert np.abs(logits.detach().numpy() - gold["logits"]).max() < 2e-5
    loss.backward()
    checked = 0
    for k in gold.files:
        if k.startswith("grad/"):
            g = sd[k[5:]].grad.reshape(-1)[:16].numpy()
            assert np.abs(g - gold[k]).max() < 1e-5 + 1e-3 * np.abs(gold[k]).max(), k
            checked += 1
    assert checked > 200
    with torch.no_grad():
        probs = FO.finetune_forward_test({k: v.detach() for k, v in sd.items()}, arch, img)
    assert np.array_equal(probs.argmax(-1).numpy(), gold["greedy_argmax"])
    assert np.abs(probs.max(-1).values.numpy() - gold["greedy_maxprob"]).max() < 1e-5


def test_targets_generator_matches_convertor_framing():
    from ccd_b200 import synthetic as S
    from ccd_b200.finetune import AttnConvertor
    conv = AttnConvertor(dict_type="DICT90", max_seq_len=25, with_unknown=True)
    assert (conv.num_classes(), conv.start_idx, conv.end_idx, conv.padding_idx, conv.unknown_idx) == (93, 91, 91, 92, 90)
    t = conv.str2tensor(["hello", "B200~", "x" * 40, "café"])
    assert t.shape == (4, 25) and t[0, 0] == 91 and t[0, 6] == 91 and (t[0, 7:] == 92).all()
    assert t[2, -1] != 92 and t[3, 4] == 90                      # truncated to max_seq_len; unknown character
    assert conv.idx2str([t[0, 1:6].tolist()]) == ["hello"]
    g = S.make_targets(64, seed=1)
    assert (g[:, 0] == 91).all() and ((g == 91).sum(1) == 2).all() and g.max() == 92


@pytest.mark.needs_reference
def test_oracle_and_dropin_against_live_reference():
    import finetune_oracle as FO
    import ref_import
    from ccd_b200 import synthetic as S
    from ccd_b200.finetune import AttnConvertor, DINO_Finetune
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference()
    for arch in ("vit_tiny", "vit_base"):
        rm = ref.dv.DINO_Finetune(S.finetune_config(arch)).eval()
        mine = DINO_Finetune(S.finetune_config(arch))
        rsd, msd = rm.state_dict(), mine.state_dict()
        assert list(rsd.keys()) == list(msd.keys())                                   # names AND order (optimizer / EMA zips)
        assert all(rsd[k].shape == msd[k].shape and rsd[k].dtype == msd[k].dtype for k in rsd)
        assert [k for k, p in rm.named_parameters() if p.requires_grad] == [k for k, p in mine.named_parameters() if p.requires_grad]
        assert torch.equal(rsd["decoder.position_enc.position_table"], msd["decoder.position_enc.position_table"])
    sd = S.fill_state_dict({k: v.shape for k, v in rm.state_dict().items()}, 31, 0.05)
    rm.load_state_dict(sd)
    img = torch.randn(3, 3, 32, 128, generator=torch.Generator().manual_seed(7))
    tgt = S.make_targets(3, seed=8)
    with torch.no_grad():
        loss, _ = rm(img, tgt, return_loss=True)
        L, _, _ = FO.finetune_forward_train(sd, "vit_base", img, tgt)
        assert abs(loss.item() - L.item()) < 1e-5 * abs(loss.item())
        assert (rm(img, None, return_loss=False) - FO.finetune_forward_test(sd, "vit_base", img)).abs().max() < 1e-5
    rc = rm.label_convertor
    mc = AttnConvertor(dict_type="DICT90", max_seq_len=25, with_unknown=True)
    words = ["hello", "World!", "a" * 30, "~`_"]
    assert torch.equal(rc.str2tensor(words), mc.str2tensor(words)) and rc.idx2char == mc.idx2char
