"""Timeline of the pipelined attention-backward kernel (diagnostic build, see csrc/mhsa_bwd.cu: CCD_ATT_TRACE).

    python tools/build_variants.py atttrace:CCD_ATT_TRACE=1                       # CPU container: cross-compiles the variant
    CCD_LIB=ccd_b200/libccd_b200_atttrace.so python tools/att_trace.py           # on a B200 (inside a gpurun call)

CTA 0 records (role, event, item, sub-tile, clock64) for every barrier wait of its TMA producer, MMA issuer and the two
softmax-gradient warpgroups.  The script runs the kernel once at the ViT-Small batch-256 shape and prints, per role, how many
cycles of an average item go to which wait and to the work in between; gpurun_out/att_trace.json keeps the raw records of
the first items.  (Never a bench value: the records perturb the kernel slightly.)
"""
import collections
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccd_b200 import lib, ops

ROLE = {0: "tma_producer", 1: "mma_issuer", 2: "softmax_wg0", 3: "softmax_wg1"}
WAITS = {(1, 2): "bar_done", (10, 11): "bar_qk+bar_vdo", (12, 13): "bar_pd", (14, 15): "bar_epi", (20, 21): "bar_free", (22, 23): "bar_s",
         (25, 26): "bar_acc"}
BEGIN = {b: (e, n) for (b, e), n in WAITS.items()}

S, H, CAP = int(os.environ.get("ATT_S", "512")), 6, 1 << 16
L = lib.load(build_if_missing=False)
if not hasattr(L, "ccd_debug_att_trace"):
    raise SystemExit("this library was built without -DCCD_ATT_TRACE=1 (see the docstring)")
L.ccd_debug_att_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint]
L.ccd_debug_att_trace_count.argtypes = [ctypes.POINTER(ctypes.c_uint)]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
E = 64 * H
qkv = torch.randn(S * 256, 3 * E, device=dev, generator=g).to(torch.bfloat16)
d_o = torch.randn(S * 256, E, device=dev, generator=g).to(torch.bfloat16)
o, lse = ops.mhsa_fwd(qkv, S, H)
ops.mhsa_bwd(qkv, o, d_o, lse, S, H)                       # warm-up (untraced: the buffer pointer is still NULL)
torch.cuda.synchronize()
buf = torch.zeros(CAP, 4, dtype=torch.int64, device=dev)
assert L.ccd_debug_att_trace(buf.data_ptr(), CAP) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.mhsa_bwd(qkv, o, d_o, lse, S, H)
e1.record()
torch.cuda.synchronize()
n = ctypes.c_uint(0)
L.ccd_debug_att_trace_count(ctypes.byref(n))
L.ccd_debug_att_trace(None, 0)
rec = buf[: min(n.value, CAP)].cpu().tolist()
# the per-role record buffers (shared memory) hold a different number of items per role: keep the items every role covered
items = min(max(r[1] for r in rec if (r[0] >> 32) == role) for role in {r[0] >> 32 for r in rec})
rec = [r for r in rec if r[1] < items]
rec.sort(key=lambda r: r[3])
t0 = rec[0][3]
per_role = collections.defaultdict(lambda: collections.defaultdict(int))
open_wait = {}
last = {}
for tag, item, sub, clk in rec:
    role, ev = tag >> 32, tag & 0xFFFFFFFF
    if ev in BEGIN:
        open_wait[(role, ev)] = clk
    else:
        for b, (e, name) in BEGIN.items():
            if e == ev and (role, b) in open_wait:
                per_role[role]["wait " + name] += clk - open_wait.pop((role, b))
    if role in last:
        per_role[role]["_span"] += clk - last[role]
    last[role] = clk
total = rec[-1][3] - t0
print(json.dumps({"kernel_us": round(e0.elapsed_time(e1) * 1e3, 1), "cta0_cycles": total, "items_of_cta0": items,
                  "cycles_per_item": round(total / items), "records": len(rec), "dropped": max(0, n.value - CAP)}))
for role in sorted(per_role):
    d = per_role[role]
    waits = {k: round(v / items) for k, v in d.items() if k.startswith("wait")}
    busy = round((d["_span"] - sum(v for k, v in d.items() if k.startswith("wait"))) / items)
    print(f"{ROLE[role]:14s} per item: " + ", ".join(f"{k} {v}" for k, v in sorted(waits.items(), key=lambda kv: -kv[1])) + f", other (work) {busy}")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/att_trace.json", "w") as f:
    json.dump({"t0": t0, "records": [[r[0] >> 32, r[0] & 0xFFFFFFFF, r[1], r[2], r[3] - t0] for r in rec if r[1] < 3]}, f)
