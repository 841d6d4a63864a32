"""The golden cases of the pretraining step: ONE table shared by the generator (make_golden.py, runs the unmodified reference)
and the tests that read the fixtures (test_oracle_golden.py on the CPU, test_pretrain_parity_gpu.py on the B200)."""

CASES = {
    # name: (arch, E, batch, out_dim, student weight seed, teacher weight seed, std, norm_last_layer)
    "cfg1_tiny_b4": ("vit_tiny", 192, 4, 65536, 1, 2, 0.05, False),
    "small_b3": ("vit_small", 384, 3, 8192, 3, 4, 0.04, True),
    "base_b2": ("vit_base", 512, 2, 4096, 5, 6, 0.03, False),
    # BASELINE cfg 2's architecture and out_dim (ViT-Small, K = 65536, norm_last_layer False as the shipped YAML) at a batch the
    # reference finishes in seconds on the host; the full batch 256 is checked on the GPU against the reference itself
    # (tests/test_full_size_reference_gpu.py)
    "small_b32_k65536": ("vit_small", 384, 32, 65536, 7, 8, 0.04, False),
    # epoch >= 30: the component labelling runs on the student's own thresholded segmentation (dino_vision.py:64-70)
    "tiny_b6_epoch30": ("vit_tiny", 192, 6, 4096, 9, 10, 0.05, False),
}
EPOCH = {"tiny_b6_epoch30": 30}
COL_STRIDE = 61                                # logits columns kept: 0, 61, 122, ...
COL_STRIDE_OF = {"small_b32_k65536": 509}      # 544 rows x 65536 columns: keep every 509th column
SEG_KEEP = {"small_b32_k65536": 4}             # segmentation logits kept for the first 4 images of each view only
GRAD_HEAD = 24                                 # leading elements of every gradient kept


def col_stride(name):
    return COL_STRIDE_OF.get(name, COL_STRIDE)
