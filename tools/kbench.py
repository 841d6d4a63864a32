"""Per-kernel timings at the shapes of the ViT-Small batch-256 step (CUDA events on the launching stream, L2 flushed
by writing a 512 MB buffer before every timed launch, median of --iters).  Prints one JSON object; used for A/B
runs of kernel variants (CCD_* environment switches / ccd_set_option) without paying for a full bench.

    python tools/kbench.py [--only gemm,mhsa] [--iters 7] [--tag name]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccd_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=7)
ap.add_argument("--tag", default="kbench")
ap.add_argument("--seqs", type=int, default=512)
ap.add_argument("--heads", type=int, default=6)
ap.add_argument("--shape", default="", help="substring filter on the gemm shape names")
a = ap.parse_args()
only = set(x for x in a.only.split(",") if x)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    ts = []
    for i in range(a.iters + 2):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


S, H = a.seqs, a.heads
E = 64 * H
T = S * 256
res = {}
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


if not only or "gemm" in only:
    # (name, M, N, K, a_mn, b_mn, epi)
    shapes = [("qkv_fwd", T, 3 * E, E, 0, 0, ops.EPI_BF16), ("proj_fwd_resid", T, E, E, 0, 0, ops.EPI_RESID),
              ("fc1_fwd_gelu_save", T, 4 * E, E, 0, 0, ops.EPI_GELU), ("fc1_fwd_gelu_nosave", T, 4 * E, E, 0, 0, ops.EPI_GELU),
              ("fc2_fwd_resid", T, E, 4 * E, 0, 0, ops.EPI_RESID), ("fc2_dgrad_dgelu", T, 4 * E, E, 0, 1, ops.EPI_DGELU),
              ("fc1_dgrad", T, E, 4 * E, 0, 1, ops.EPI_BF16), ("qkv_dgrad", T, E, 3 * E, 0, 1, ops.EPI_BF16),
              ("proj_dgrad", T, E, E, 0, 1, ops.EPI_BF16), ("fc1_wgrad", 4 * E, E, T, 1, 1, ops.EPI_F32),
              ("fc2_wgrad", E, 4 * E, T, 1, 1, ops.EPI_F32), ("qkv_wgrad", 3 * E, E, T, 1, 1, ops.EPI_F32),
              ("proj_wgrad", E, E, T, 1, 1, ops.EPI_F32)]
    for name, M, N, K, amn, bmn, epi in shapes:
        if a.shape and a.shape not in name:
            continue
        A = rnd(K, M) if amn else rnd(M, K)
        B = rnd(K, N, scale=0.05) if bmn else rnd(N, K, scale=0.05)
        bias = torch.randn(N, device=dev, generator=g) if epi != ops.EPI_F32 else None
        out0 = out1 = aux = None
        splits = 1
        if epi == ops.EPI_BF16:
            out0 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        elif epi == ops.EPI_GELU:
            out0 = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if "nosave" not in name else None
            out1 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        elif epi == ops.EPI_RESID:
            out0 = torch.empty(M, N, dtype=torch.float32, device=dev)
            aux = torch.randn(M, N, device=dev, generator=g)
        elif epi == ops.EPI_DGELU:
            out0 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            aux = rnd(M, N)
            bias = None
        elif epi == ops.EPI_F32:
            out0 = torch.zeros(M, N, dtype=torch.float32, device=dev)
            splits = ops.wgrad_splits(M, N, K)
        ms = timeit(lambda: ops.gemm(A, B, M, N, K, amn, bmn, epi, bias, out0, out1, aux, 0, splits))
        res["gemm_" + name] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        del A, B, out0, out1, aux

if not only or "mhsa" in only:
    qkv = rnd(T, 3 * E)
    ms = timeit(lambda: ops.mhsa_fwd(qkv, S, H))
    res["mhsa_fwd"] = {"ms": round(ms, 4), "tflops": round(4.0 * 256 * 256 * 64 * S * H / ms / 1e9, 1)}
    o, lse = ops.mhsa_fwd(qkv, S, H)
    d_o = rnd(T, E)
    ms = timeit(lambda: ops.mhsa_bwd(qkv, o, d_o, lse, S, H))
    res["mhsa_bwd(+delta)"] = {"ms": round(ms, 4), "tflops": round(10.0 * 256 * 256 * 64 * S * H / ms / 1e9, 1)}
    del qkv, o, lse, d_o

if not only or "row" in only:
    x = torch.randn(T, E, device=dev, generator=g)
    gm, bt = torch.ones(E, device=dev), torch.zeros(E, device=dev)
    ms = timeit(lambda: ops.layernorm_fwd(x, gm, bt))
    res["layernorm_fwd"] = {"ms": round(ms, 4), "gbs": round(T * E * 6 / ms / 1e6, 1)}
    dy = rnd(T, E)
    resid = torch.randn(T, E, device=dev, generator=g)
    dg, db, dbn = torch.zeros(E, device=dev), torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    ms = timeit(lambda: ops.layernorm_bwd(x, gm, dy, resid, dg, db, dbias_next=dbn))
    res["layernorm_bwd"] = {"ms": round(ms, 4), "gbs": round(T * E * 16 / ms / 1e6, 1)}
    dh = rnd(T, 4 * E)
    cs = torch.zeros(4 * E, device=dev)
    ms = timeit(lambda: ops.colsum_bf16(dh, cs))
    res["colsum_bf16_4E"] = {"ms": round(ms, 4), "gbs": round(T * 4 * E * 2 / ms / 1e6, 1)}

res["env"] = {k: v for k, v in os.environ.items() if k.startswith("CCD_")}
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/{a.tag}.json", "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
