"""Timeline of one CTA of the persistent GEMM (diagnostic build, see csrc/gemm_umma.cu: CCD_GEMM_TRACE).

    python tools/build_variants.py gemmtrace:CCD_GEMM_TRACE=1                                   # CPU container
    CCD_LIB=ccd_b200/libccd_b200_gemmtrace.so python tools/gemm_trace.py [shape ...]            # on a B200

For every requested shape of the ViT-Small step (default: the epilogue-bound ones) CTA 0 records its producers' waits for free
ring slots, the issuer's waits for operands / a free accumulator, and the epilogue warpgroups' waits for a finished accumulator
and their per-tile store time.  Prints cycles per tile by role; gpurun_out/gemm_trace_<shape>.json keeps the first tiles raw.
"""
import collections
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccd_b200 import lib, ops

ROLE = {0: "tma_A", 1: "mma_issuer", 2: "tma_B", 3: "epilogue_wg0", 7: "epilogue_wg1"}
BEGIN = {1: (2, "empty slot"), 10: (11, "tmem_empty"), 12: (13, "full slot"), 20: (21, "tmem_full")}
T, E = 131072, 384
SHAPES = {  # name: (M, N, K, a_mn, b_mn, epi)
    "fc1_fwd_gelu_nosave": (T, 4 * E, E, 0, 0, ops.EPI_GELU), "fc1_fwd_gelu_save": (T, 4 * E, E, 0, 0, ops.EPI_GELU),
    "qkv_fwd": (T, 3 * E, E, 0, 0, ops.EPI_BF16), "proj_fwd_resid": (T, E, E, 0, 0, ops.EPI_RESID),
    "fc2_fwd_resid": (T, E, 4 * E, 0, 0, ops.EPI_RESID), "fc2_dgrad_dgelu": (T, 4 * E, E, 0, 1, ops.EPI_DGELU),
    "fc1_dgrad": (T, E, 4 * E, 0, 1, ops.EPI_BF16), "fc2_wgrad": (E, 4 * E, T, 1, 1, ops.EPI_F32),
}
L = lib.load(build_if_missing=False)
if not hasattr(L, "ccd_debug_gemm_trace"):
    raise SystemExit("this library was built without -DCCD_GEMM_TRACE=1 (see the docstring)")
L.ccd_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
CAP, ROLES = 4096, 11                    # csrc/gemm_umma.cu: GEMM_TRACE_CAP records per role region + 11 counters at the end
buf = torch.zeros(ROLES * CAP + ROLES, dtype=torch.int64, device=dev)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


for name in (sys.argv[1:] or ["fc1_fwd_gelu_nosave", "fc2_dgrad_dgelu", "qkv_fwd", "fc2_fwd_resid"]):
    M, N, K, amn, bmn, epi = SHAPES[name]
    A = rnd(K, M) if amn else rnd(M, K)
    B = rnd(K, N, scale=0.05) if bmn else rnd(N, K, scale=0.05)
    bias = torch.randn(N, device=dev, generator=g) if epi not in (ops.EPI_F32, ops.EPI_DGELU) else None
    out0 = out1 = aux = None
    splits = 1
    if epi == ops.EPI_BF16:
        out0 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    elif epi == ops.EPI_GELU:
        out0 = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if "nosave" not in name else None
        out1 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    elif epi == ops.EPI_RESID:
        out0, aux = torch.empty(M, N, device=dev), torch.randn(M, N, device=dev, generator=g)
    elif epi == ops.EPI_DGELU:
        out0, aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev), rnd(M, N)
    else:
        out0, splits = torch.zeros(M, N, device=dev), ops.wgrad_splits(M, N, K)
    run = lambda: ops.gemm(A, B, M, N, K, amn, bmn, epi, bias, out0, out1, aux, 0, splits)
    run()
    torch.cuda.synchronize()
    buf.zero_()
    assert L.ccd_debug_gemm_trace(buf.data_ptr(), CAP) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    L.ccd_debug_gemm_trace(None, 0)
    host = buf.cpu().numpy().view("uint64")
    rec = []
    for role in ROLE:
        cnt = int(host[ROLES * CAP + role])
        for v in host[role * CAP: role * CAP + min(cnt, CAP)]:
            lo, clk = int(v) & 0xFFFFFFFF, int(v) >> 32
            rec.append(((role << 32) | (lo & 0xFF), lo >> 16, (lo >> 8) & 0xFF, clk))
    # keep the tiles every role covered (the regions fill at different rates)
    covered = min(max(r[1] for r in rec if (r[0] >> 32) == role) for role in ROLE if any((r[0] >> 32) == role for r in rec))
    rec = [r for r in rec if r[1] <= covered]
    n = ctypes.c_uint(len(rec))
    rec.sort(key=lambda r: r[3])
    t0 = rec[0][3]
    tiles = len({r[1] for r in rec})
    per_role, open_wait, last = collections.defaultdict(lambda: collections.defaultdict(int)), {}, {}
    for tag, item, aux_, clk in rec:
        role, ev = tag >> 32, tag & 0xFFFFFFFF
        if ev in BEGIN:
            open_wait[(role, ev)] = clk
        else:
            for b, (e, nm) in BEGIN.items():
                if e == ev and (role, b) in open_wait:
                    per_role[role]["wait " + nm] += clk - open_wait.pop((role, b))
        if role in last:
            per_role[role]["_span"] += clk - last[role]
        last[role] = clk
    total = rec[-1][3] - t0
    print(json.dumps({"shape": name, "kernel_us": round(e0.elapsed_time(e1) * 1e3, 1), "cta0_cycles": total, "tiles_of_cta0": tiles,
                      "cycles_per_tile": round(total / max(1, tiles)), "records": len(rec), "dropped": max(0, n.value - CAP)}))
    for role in sorted(per_role):
        d = per_role[role]
        waits = {k: round(v / tiles) for k, v in d.items() if k.startswith("wait")}
        busy = round((d["_span"] - sum(v for k, v in d.items() if k.startswith("wait"))) / tiles)
        print(f"  {ROLE.get(role, role):13} per tile: " + ", ".join(f"{k} {v}" for k, v in sorted(waits.items(), key=lambda kv: -kv[1])) + f", other (work) {busy}")
    os.makedirs("gpurun_out", exist_ok=True)
    first = sorted({r[1] for r in rec})[:3]
    with open(f"gpurun_out/gemm_trace_{name}.json", "w") as f:
        json.dump({"records": [[r[0] >> 32, r[0] & 0xFFFFFFFF, r[1], r[2], r[3] - t0] for r in rec if r[1] in first]}, f)
    del A, B, out0, out1, aux
