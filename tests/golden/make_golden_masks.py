"""Generates tests/golden/kmeans_masks.npz by running the UNMODIFIED reference function clusterpixels
(Dino/utils/kmeans.py:8-24, the same code as mask_create/generate_mask.py:13-29) on synthetic grey text crops.

    python tests/golden/make_golden_masks.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import mask_oracle as MO  # noqa: E402
from test_mask_oracle import _reference_clusterpixels  # noqa: E402


def main():
    clusterpixels = _reference_clusterpixels()
    imgs = np.concatenate([MO.synthetic_text_crops(40, seed=7), MO.synthetic_text_crops(8, h=48, w=160, seed=8)[:, :32, :128],
                           MO.synthetic_text_crops(8, seed=9, noise=30.0)])
    masks, votes = [], []
    for i, im in enumerate(imgs):
        np.random.seed(i)                                   # scipy's kmeans draws its initial centroids from numpy's global RNG
        code_final = np.asarray(clusterpixels(im, 2)).astype(np.uint8)
        masks.append(code_final)
        # border votes of the code BEFORE the flip decide whether the reference's answer was seed dependent (num == 2)
        v = MO.border_votes(code_final)
        votes.append(2 if v == 2 else 0)
    np.savez_compressed(os.path.join(HERE, "kmeans_masks.npz"), images=imgs, masks=np.stack(masks), votes=np.array(votes, dtype=np.uint8))
    print("golden masks:", len(imgs), "seed-dependent:", int(sum(v == 2 for v in votes)))


if __name__ == "__main__":
    main()
