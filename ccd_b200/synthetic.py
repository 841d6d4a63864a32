"""Synthetic inputs and deterministic weights for parity tests and bench.py.

Shapes follow the reference's data contract (Dino/dataset/datasetsupervised_kmeans.py:48-87):
image_tensors [B,3,3,32,128] fp32 (view 0 is never forwarded, Dino/model/dino_vision.py:52-54),
masks [B,32,128] fp32 in {0,1}, metrics [B,3,3] affine thetas.
"""
import zlib

import torch


def make_batch(batch, seed=1234, device="cpu"):
    """n_b = 4 + (b mod 8) character boxes, rows 6..25, cols 4+11c .. 4+11c+7;
    identity theta when b mod 3 == 0 else a mild affine."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, 3, 32, 128, generator=g)
    masks = torch.zeros(batch, 32, 128)
    for b in range(batch):
        for c in range(4 + (b % 8)):
            masks[b, 6:26, 4 + 11 * c: 4 + 11 * c + 8] = 1.0
    u = torch.rand(batch, 5, generator=g)
    metrics = torch.zeros(batch, 3, 3)
    metrics[:, 0, 0] = 0.9 + 0.2 * u[:, 0]
    metrics[:, 0, 1] = -0.3 + 0.6 * u[:, 1]
    metrics[:, 0, 2] = -0.05 + 0.1 * u[:, 2]
    metrics[:, 1, 1] = 0.9 + 0.2 * u[:, 3]
    metrics[:, 1, 2] = -0.05 + 0.1 * u[:, 4]
    metrics[:, 2, 2] = 1.0
    ident = torch.eye(3)
    for b in range(0, batch, 3):
        metrics[b] = ident
    return x.to(device), masks.to(device), metrics.to(device)


def random_masks(batch, seed=0, density=0.5, blobs=True):
    """Irregular masks for the connected-component tests: random rectangles / noise."""
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(batch, 32, 128)
    for b in range(batch):
        if blobs:
            n = int(torch.randint(0, 40, (1,), generator=g))
            for _ in range(n):
                y0 = int(torch.randint(0, 30, (1,), generator=g)); x0 = int(torch.randint(0, 124, (1,), generator=g))
                hh = int(torch.randint(1, 14, (1,), generator=g)); ww = int(torch.randint(1, 14, (1,), generator=g))
                m[b, y0:y0 + hh, x0:x0 + ww] = 1.0
            # punch some holes so components get non-convex
            holes = torch.rand(32, 128, generator=g) < 0.08
            m[b][holes] = 0.0
        else:
            m[b] = (torch.rand(32, 128, generator=g) < density).float()
    return m


def fill_state_dict(shapes, seed=0, std=0.02):
    """Deterministic weights that depend only on (name, shape, seed) -- independent of module construction
    order, so the reference modules, the oracle and the product can all be loaded with identical values.
    `shapes` = {name: torch.Size}."""
    out = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=torch.long)
        elif leaf == "running_var":
            out[name] = 1.0 + 0.1 * torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            out[name] = 0.01 * torch.randn(shape, generator=g)
        elif leaf == "position_table":                      # PositionalEncoding buffer (transformer_module.py:138-149): a constant
            import numpy as np
            d = shape[-1]
            den = torch.Tensor([1.0 / np.power(10000, 2 * (j // 2) / d) for j in range(d)]).view(1, -1)
            tab = torch.arange(shape[-2]).unsqueeze(-1).float() * den
            tab[:, 0::2] = torch.sin(tab[:, 0::2])
            tab[:, 1::2] = torch.cos(tab[:, 1::2])
            out[name] = tab.reshape(shape)
        elif leaf == "weight_g":
            out[name] = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:          # LayerNorm / BatchNorm scale
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            out[name] = 0.02 * torch.randn(shape, generator=g)
        else:
            out[name] = std * torch.randn(shape, generator=g)
    return out


def finetune_config(arch="vit_small", drop_path_rate=0.0):
    """The attribute bag DINO_Finetune reads (Dino/configs/CCD_vision_model_ARD.yaml:62-85 flattened as train_finetune.py does)."""
    import types
    return types.SimpleNamespace(arch=arch, patch_size=4, drop_path_rate=drop_path_rate, decoder_max_seq_len=25, decoder_n_layers=6,
                                 decoder_d_embedding=512, decoder_n_head=8, decoder_d_k=64, decoder_d_v=64, decoder_d_model=512,
                                 decoder_d_inner=256)


def make_targets(n, seed=0, max_seq_len=25, start_idx=91, pad_idx=92):
    """Synthetic label tensors framed like AttnConvertor.str2tensor (Dino/convertor/attn.py:87-105): BOS, 1..max_seq_len-2
    characters in 0..89, EOS (= BOS index), PAD up to max_seq_len  (SURVEY.md section 8d, cfg5)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.full((n, max_seq_len), pad_idx, dtype=torch.long)
    for i in range(n):
        ln = int(torch.randint(1, max_seq_len - 1, (1,), generator=g))
        t[i, 0] = start_idx
        t[i, 1:1 + ln] = torch.randint(0, 90, (ln,), generator=g)
        t[i, 1 + ln] = start_idx
    return t
