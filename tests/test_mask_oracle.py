"""CPU: the mask-generation oracle (oracle/mask_oracle.py) against the UNMODIFIED reference function clusterpixels
(Dino/utils/kmeans.py:8-24 = mask_create/generate_mask.py:13-29, scipy k-means) and its committed golden masks."""
import os
import sys

import numpy as np
import pytest

import mask_oracle as MO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_masks.npz")


def _reference_clusterpixels():
    """Imports the reference's clusterpixels behind the import stubs of oracle/ref_import.py (pylab / mkl are missing here;
    `np.float`, which the function uses, was removed from numpy 1.24: restored as the alias it was)."""
    import importlib
    import ref_import
    ref_import._install_stubs()
    if not hasattr(np, "float"):
        np.float = float
    saved = list(sys.path)
    for k in [k for k in sys.modules if k == "Dino" or k.startswith("Dino.")]:
        del sys.modules[k]
    sys.path[:] = [ref_import.REFERENCE_ROOT] + [e for e in saved if not os.path.isfile(os.path.join(e or os.getcwd(), "Dino", "__init__.py"))]
    try:
        mod = importlib.import_module("Dino.utils.kmeans")
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k == "Dino" or k.startswith("Dino.")]:
            del sys.modules[k]
    return mod.clusterpixels


def _compare(im, ref, got):
    """The reference mask must be THE SAME ALGORITHM's output: a threshold on the grey level (then the polarity rule), which differs
    from the exact 2-means optimum in at most two grey levels -- scipy stops its Lloyd iterations once the distortion
    changes by less than 1e-5, which on low-contrast crops is a level or two short of the fixed point.  Returns pixel agreement."""
    if MO.border_votes(ref) == 2 or MO.border_votes(1 - ref) == 2:      # the reference's answer depends on its random centroid order
        assert np.array_equal(got, ref) or np.array_equal(got, 1 - ref) or (got == ref).mean() > 0.98 or (got == 1 - ref).mean() > 0.98
        return None
    bright = ref if im[ref == 1].mean() > im[ref == 0].mean() else 1 - ref
    t_ref = int(im[bright == 0].max())
    assert im[bright == 1].min() > t_ref                                  # a threshold partition
    differ = np.unique(im[got != ref])                                    # grey levels the two partitions classify differently
    assert len(differ) <= 2, differ
    assert (got.sum() > got.size / 2) == (ref.sum() > ref.size / 2)       # same polarity decision
    return float((got == ref).mean())


def test_oracle_matches_committed_reference_masks():
    z = np.load(GOLD)
    agree = [_compare(im, w, MO.cluster_pixels(im)) for im, w in zip(z["images"], z["masks"])]
    agree = [a for a in agree if a is not None]
    assert min(agree) >= 0.985, min(agree)
    assert sum(a == 1.0 for a in agree) >= 0.85 * len(agree)              # bit-exact wherever the two clusters are separated


def test_constant_and_two_level_images():
    assert MO.cluster_pixels(np.full((32, 128), 77, np.uint8)).sum() == 0
    im = np.full((32, 128), 200, np.uint8)
    im[8:24, 30:50] = 20                              # dark text on a bright background -> text is 1 after the polarity flip
    m = MO.cluster_pixels(im)
    assert m[10, 40] == 1 and m[0, 0] == 0 and m.sum() == 16 * 20
    m2 = MO.cluster_pixels(255 - im)                  # bright text on dark: no flip needed
    assert np.array_equal(m, m2)


@pytest.mark.needs_reference
def test_oracle_matches_reference_live():
    clusterpixels = _reference_clusterpixels()
    imgs = MO.synthetic_text_crops(24, seed=123)
    agree = []
    for i, im in enumerate(imgs):
        np.random.seed(1000 + i)
        ref = np.asarray(clusterpixels(im, 2)).astype(np.uint8)
        agree.append(_compare(im, ref, MO.cluster_pixels(im)))
    agree = [a for a in agree if a is not None]
    assert min(agree) >= 0.985 and sum(a == 1.0 for a in agree) >= 0.85 * len(agree)


def test_threshold_scan_minimises_the_two_means_distortion():
    """The histogram scan maximises s0^2/n0 + s1^2/n1, i.e. minimises the k-means distortion sum |v - c(v)|^2 over all threshold
    partitions -- checked by brute force -- and the partition is a Lloyd fixed point (nearest-centroid assignment reproduces it)."""
    g = np.random.default_rng(0)
    for trial in range(40):
        n = int(g.integers(20, 400))
        vals = np.clip(np.rint(np.concatenate([g.normal(g.uniform(20, 120), g.uniform(3, 25), n),
                                               g.normal(g.uniform(130, 240), g.uniform(3, 25), int(g.integers(5, 300)))])), 0, 255).astype(np.uint8)
        im = vals.reshape(1, -1)
        t, c0, c1 = MO.two_means_threshold(im)
        v = vals.astype(np.float64)

        def sse(th):
            a, b = v[v <= th], v[v > th]
            return ((a - a.mean()) ** 2).sum() + ((b - b.mean()) ** 2).sum() if len(a) and len(b) else np.inf
        best = min(sse(th) for th in range(255))
        assert abs(sse(t) - best) <= 1e-9 * max(1.0, best)
        mid = 0.5 * (c0 + c1)
        assert np.array_equal(v > mid, v > t)            # a fixed point of Lloyd's iteration


def test_affine_theta_identity_and_scaling():
    th = MO.affine_theta(np.eye(3), 64, 300)
    assert np.allclose(th, np.eye(3), atol=1e-7)
    # a pure translation by (dx, dy) source pixels becomes 2 dx / ((w_src / 128) * 127) in normalised units
    m = np.eye(3); m[0, 2], m[1, 2] = 30.0, -4.0
    th = MO.affine_theta(m, 64, 256)
    assert np.allclose(th[0, 2], 30.0 / 2.0 * 2 / 127, atol=1e-6) and np.allclose(th[1, 2], -4.0 / 2.0 * 2 / 31, atol=1e-6)
