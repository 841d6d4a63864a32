"""Per-shape DRAM traffic of the dominant kernel (the persistent tcgen05 GEMM) over ONE pretraining step.

On the GPU box, under ncu (metrics only -- every GEMM launch of the step is captured, 241 of them):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_umma \\
        --profile-from-start off --csv --log-file gpurun_out/gemm_traffic_r02.csv python tools/gemm_traffic.py --run
  -> also writes gpurun_out/gemm_shapes_r02.json: the (M, N, K, layouts, epilogue) of every GEMM launch of that step, in order.
In the CPU container:
    python tools/gemm_traffic.py --merge r02      -> profiles/gemm_traffic_r02.json (+ .md): per shape launches, measured bytes,
                                                     algorithmic bytes, kernel time; bench.py's roofline.traffic reads it.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EPI = {0: "bf16", 1: "gelu", 2: "resid", 3: "f32", 4: "dgelu", 5: "pos"}


def algorithmic_bytes(shape, saved_preact=True):
    """Operands once + outputs once (+ auxiliary operands of the fused epilogue) -- what a launch MUST move."""
    if shape[0] == "conv":
        _, M, N, K = shape
        return 2 * (M * K / 9.0 + N * K) + 2 * M * N          # activation taps overlap: ~1/9 of the im2col extent is unique (3x3)
    M, N, K, a_mn, b_mn, epi = shape
    b = 2 * (M * K + N * K)
    name = EPI.get(epi, str(epi))
    if name == "bf16":
        b += 2 * M * N
    elif name == "gelu":
        b += 2 * M * N * (2 if saved_preact else 1)
    elif name == "resid":
        b += 4 * M * N * 2
    elif name == "f32":
        b += 4 * M * N
    elif name == "dgelu":
        b += 2 * M * N * 2
    elif name == "pos":
        b += 4 * M * N + 4 * 256 * N
    return b


def run():
    import torch
    from ccd_b200 import ops, synthetic as S
    from ccd_b200.trainer import PretrainStep
    t = PretrainStep(arch="vit_small", batch_per_gpu=256, device=torch.device("cuda", 0))
    t.student.train()
    batch = tuple(v.cuda() for v in S.make_batch(256, seed=1234))
    for _ in range(2):
        t.step(*batch, sync_loss=False)
    torch.cuda.synchronize()
    prof = {"names": {"ccd_gemm_bf16", "ccd_conv_gemm"}, "events": []}
    ops.PROFILE = prof
    torch.cuda.cudart().cudaProfilerStart()
    t.step(*batch, sync_loss=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    ops.PROFILE = None
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_shapes_r02.json", "w") as f:
        json.dump([list(w[1]) if isinstance(w[1], (list, tuple)) else w[1] for _, w, _, _ in prof["events"]], f)
    print("recorded", len(prof["events"]), "GEMM launches")


def merge(tag):
    shapes = json.load(open(os.path.join(ROOT, "gpurun_out", f"gemm_shapes_{tag}.json")))
    with open(os.path.join(ROOT, "gpurun_out", f"gemm_traffic_{tag}.csv")) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if "gemm_umma" not in row["Kernel Name"]:
            continue
        d = per_launch.setdefault(row["ID"], {})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        else:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        d[row["Metric Name"]] = v
    launches = list(per_launch.values())
    assert len(launches) == len(shapes), (len(launches), len(shapes))
    agg = collections.OrderedDict()
    for sh, m in zip(shapes, launches):
        key = json.dumps(sh)
        a = agg.setdefault(key, {"shape": sh, "launches": 0, "dram_bytes": 0.0, "us": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        a["us"] += m["gpu__time_duration.sum"]
    rows = []
    for a in agg.values():
        sh = tuple(a["shape"])
        alg = algorithmic_bytes(sh)
        rows.append({"shape": a["shape"], "launches_per_step": a["launches"], "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
                     "algorithmic_bytes_per_launch": alg, "ratio": a["dram_bytes"] / a["launches"] / alg,
                     "us_per_launch_under_ncu": a["us"] / a["launches"]})
    tot = sum(r["dram_bytes_per_launch"] * r["launches_per_step"] for r in rows)
    n = sum(r["launches_per_step"] for r in rows)
    out = {"tag": tag, "launches_per_step": n, "dram_bytes_per_step": tot, "dram_bytes_per_launch_avg": tot / n,
           "algorithmic_bytes_per_launch_avg": sum(r["algorithmic_bytes_per_launch"] * r["launches_per_step"] for r in rows) / n,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over EVERY gemm_umma launch of one pretraining step "
                     "(ViT-Small, batch 256, 1 x B200), shapes recorded in launch order by tools/gemm_traffic.py", "shapes": rows}
    with open(os.path.join(ROOT, "profiles", f"gemm_traffic_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
    md = [f"# DRAM traffic of the tcgen05 GEMM per shape ({tag}): one pretraining step, ViT-Small batch 256", "",
          f"{n} launches, {tot / 1e9:.2f} GB of DRAM traffic per step = {tot / n / 1e6:.1f} MB per launch on average "
          f"(algorithmic: {out['algorithmic_bytes_per_launch_avg'] / 1e6:.1f} MB).", "",
          "| shape (M, N, K, a_mn, b_mn, epilogue) | launches | measured MB / launch | algorithmic MB / launch | ratio |", "|---|---:|---:|---:|---:|"]
    for r in sorted(rows, key=lambda r: -r["dram_bytes_per_launch"] * r["launches_per_step"]):
        sh = r["shape"]
        label = f"conv {sh[1:]}" if sh[0] == "conv" else f"{sh[0]} x {sh[1]} x {sh[2]}, {'MN' if sh[3] else 'K'}/{'MN' if sh[4] else 'K'}, {EPI.get(sh[5], sh[5])}"
        md.append(f"| {label} | {r['launches_per_step']} | {r['dram_bytes_per_launch'] / 1e6:.1f} | {r['algorithmic_bytes_per_launch'] / 1e6:.1f} | {r['ratio']:.2f} |")
    with open(os.path.join(ROOT, "profiles", f"gemm_traffic_{tag}.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    if "--run" in sys.argv:
        run()
    elif "--merge" in sys.argv:
        merge(sys.argv[sys.argv.index("--merge") + 1])
