"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's text-mask generation,
`clusterpixels(im, 2)`  (mask_create/generate_mask.py:13-29, identical copy in Dino/utils/kmeans.py:8-24), the checker for the
CUDA kernel `kmeans_mask_kernel` (ccd_b200/csrc/charseg.cu, C ABI ccd_kmeans_mask).

What the reference does: scipy.cluster.vq.kmeans(grey levels, k = 2) -- Lloyd iterations from 20 random initialisations, the
codebook with the lowest distortion is kept (scipy 1.x `kmeans`, `iter=20`, `thresh=1e-5`; third-party, not under
/root/reference; requirement.txt pins scipy==1.7.3, this image has scipy 1.16) -- then `vq` assigns every pixel to the nearest
centroid, and the code is inverted when at least three of the four border lines are mostly 1 (:21-29).

Restated here: in one dimension every Lloyd fixed point is a threshold on the grey level, and the minimum-distortion one is
found exactly from the 256-bin histogram (`two_means_threshold`).  The centroid ORDER is canonical (dark 0, bright 1); scipy's
depends on its random initialisation, which changes the final mask only when exactly two border sums exceed half.

Pinning: tests/test_mask_oracle.py runs the unmodified reference function (seeded) on synthetic text crops and compares.
"""
import numpy as np


def two_means_threshold(grey_u8):
    """-> (t, c0, c1): class 0 = levels <= t maximising s0^2/n0 + s1^2/n1 (ties -> smallest t), its two centroids; None for a
    constant image.  Same arithmetic (float64, same operation order) as the kernel."""
    hist = np.bincount(grey_u8.reshape(-1), minlength=256).astype(np.int64)
    n = np.cumsum(hist)
    s = np.cumsum(hist * np.arange(256, dtype=np.int64))
    N, S = int(n[255]), int(s[255])
    best_j, best_t = -1.0, -1
    for t in range(255):
        n0 = int(n[t])
        if n0 <= 0 or n0 >= N:
            continue
        s0, s1 = float(s[t]), float(S - int(s[t]))
        j = s0 * s0 / float(n0) + s1 * s1 / float(N - n0)
        if j > best_j:
            best_j, best_t = j, t
    if best_t < 0:
        return None
    c0 = float(s[best_t]) / float(n[best_t])
    c1 = float(S - int(s[best_t])) / float(N - int(n[best_t]))
    return best_t, c0, c1


def cluster_pixels(grey_u8):
    """clusterpixels(im, 2) -> uint8 {0,1} [H,W] (text = 1)."""
    h, w = grey_u8.shape
    r = two_means_threshold(grey_u8)
    if r is None:
        return np.zeros((h, w), dtype=np.uint8)                 # scipy drops the empty cluster: every code is 0
    _, c0, c1 = r
    mid = np.float32(0.5 * (c0 + c1))
    code = (grey_u8.astype(np.float32) > mid).astype(np.int64)  # vq: nearest centroid, ties -> index 0        (:19)
    fc, lc = code[:, 0].sum(), code[:, -1].sum()                # (:21-24)
    fr, lr = code[0, :].sum(), code[-1, :].sum()
    num = int(fr > w // 2) + int(lr > w // 2) + int(fc > h // 2) + int(lc > h // 2)                             # (:25)
    return (1 - code if num >= 3 else code).astype(np.uint8)    # (:26-29)


def border_votes(code):
    """The `num` of generate_mask.py:25 for a given code image (tests use it to skip the seed-dependent num == 2 cases)."""
    h, w = code.shape
    return int(code[0, :].sum() > w // 2) + int(code[-1, :].sum() > w // 2) + int(code[:, 0].sum() > h // 2) + \
        int(code[:, -1].sum() > h // 2)


def synthetic_text_crops(n, h=32, w=128, seed=0, noise=12.0):
    """Grey crops that look like the reference's inputs: a background level, a few character-like strokes of a contrasting
    level (dark on bright or bright on dark), smooth illumination gradient, gaussian noise."""
    g = np.random.default_rng(seed)
    out = np.zeros((n, h, w), dtype=np.uint8)
    for i in range(n):
        bg, fg = g.uniform(120, 235), g.uniform(10, 90)
        if g.random() < 0.5:
            bg, fg = fg, bg
        img = np.full((h, w), bg, dtype=np.float64)
        img += np.linspace(-1, 1, w)[None, :] * g.uniform(0, 15) + np.linspace(-1, 1, h)[:, None] * g.uniform(0, 8)
        nch = int(g.integers(2, max(3, w // 14)))
        for c in range(nch):
            x0 = 4 + c * (w - 8) // nch + int(g.integers(0, 3))
            cw = max(3, (w - 8) // nch - int(g.integers(3, 7)))
            y0, y1 = int(g.integers(3, h // 4)), int(g.integers(3 * h // 4, h - 3))
            sw = int(g.integers(2, 4))
            img[y0:y1, x0:x0 + sw] = fg                          # vertical stroke
            img[y0:y0 + sw, x0:x0 + cw] = fg                     # top bar
            if g.random() < 0.6:
                img[(y0 + y1) // 2:(y0 + y1) // 2 + sw, x0:x0 + cw] = fg
            if g.random() < 0.5:
                img[y0:y1, x0 + cw - sw:x0 + cw] = fg
        img += g.normal(0, noise, size=(h, w))
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def affine_theta(m_inv, src_h, src_w, img_h=32, img_w=128):
    """datasetsupervised_kmeans.py:63-71, line for line (numpy float64, cast to float32 as :87 does): `m_inv` is imgaug's
    `matric[0]._inv_matrix` of the warp applied to the source-size image."""
    W_scale = src_w / img_w
    H_scale = src_h / img_h
    W_inv = np.array([[1 / W_scale, 0, 0], [0, 1 / H_scale, 0], [0, 0, 1]])
    W = np.array([[W_scale, 0, 0], [0, H_scale, 0], [0, 0, 1]])
    metric = np.matmul(np.matmul(W_inv, m_inv), W)
    W_ = np.array([[2 / (img_w - 1), 0, -1], [0, 2 / (img_h - 1), -1], [0, 0, 1]])
    theta = np.matmul(np.matmul(W_, metric), np.linalg.inv(W_))
    return np.array(theta, dtype=np.float32)
