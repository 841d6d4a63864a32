"""Generates tests/golden/finetune_*.npz by executing the UNMODIFIED reference DINO_Finetune (/root/reference) in the build
container (model.eval(): every dropout is the identity -- the parity configuration).

    python tests/golden/make_golden_finetune.py

Inputs and weights are regenerated from seeds by ccd_b200.synthetic, so only OUTPUTS are stored: the TFLoss value, the
logits, the leading elements of every parameter gradient and the greedy-decoding result.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_import  # noqa: E402
from ccd_b200 import synthetic as S  # noqa: E402

CASES = {"finetune_vit_tiny_b4": ("vit_tiny", 4, 5, 0.05), "finetune_vit_small_b3": ("vit_small", 3, 6, 0.04)}
GRAD_HEAD = 16


def main():
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference()
    for name, (arch, n, wseed, std) in CASES.items():
        model = ref.dv.DINO_Finetune(S.finetune_config(arch)).eval()
        sd = S.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, wseed, std)
        model.load_state_dict(sd)
        img = torch.randn(n, 3, 32, 128, generator=torch.Generator().manual_seed(100 + n))
        tgt = S.make_targets(n, seed=200 + n)
        loss, _ = model(img, tgt, return_loss=True)
        loss.backward()
        feat = model.extract_feat(img)
        logits, _ = model.decoder(feat, model.encoder(feat), {"padded_targets": tgt}, train_mode=True)
        with torch.no_grad():
            probs = model(img, None, return_loss=False)
        out = {"loss": np.float64(loss.item()), "logits": logits.detach().numpy().astype(np.float32),
               "greedy_argmax": probs.argmax(-1).numpy().astype(np.int16), "greedy_maxprob": probs.max(-1).values.numpy().astype(np.float32)}
        for k, p in model.named_parameters():
            if p.grad is not None:
                out["grad/" + k] = p.grad.reshape(-1)[:GRAD_HEAD].numpy().astype(np.float32)
                out["gnorm/" + k] = np.float64(p.grad.norm().item())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "loss", loss.item(), "keys", len(out))


if __name__ == "__main__":
    main()
