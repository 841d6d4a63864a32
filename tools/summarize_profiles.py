"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/*.ncu-rep, launches_*.csv) into the tracked
summaries under profiles/:   python tools/summarize_profiles.py r01
Runs in the CPU container (ncu -i reads reports without a GPU)."""
import collections
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu.sum",
       "lts__t_sector_hit_rate.pct"]


def launch_list(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v; a[1] += 1; tot += v
    out = [f"# ncu launch list of ONE pretraining step ({tag}) -- ViT-Small, batch 256, 1 x B200",
           "", "`ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/profile_step.py` (cold-cache, serialised:",
           "compare SHARES, not absolutes).", "", f"total {tot / 1e3:.2f} ms over {sum(a[1] for a in agg.values())} launches", "",
           "| ms | share | launches | kernel |", "|---:|---:|---:|---|"]
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if t / tot < 0.001:
            continue
        out.append(f"| {t / 1e3:.3f} | {100 * t / tot:.1f}% | {n} | `{k[:110]}` |")
    return "\n".join(out) + "\n"


def full_report(path):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")][:100]}
        for m in RAW:
            if m in hdr:
                d[m] = row[hdr.index(m)] + " " + units[hdr.index(m)]
        out.append(d)
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    ll = launch_list(tag)
    if ll:
        with open(os.path.join(OUT, f"launch_list_{tag}.md"), "w") as f:
            f.write(ll)
    summary = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"full_*_{tag}.ncu-rep"))):
        rep = full_report(path)
        if rep:
            summary[os.path.basename(path)] = rep
    if summary:
        with open(os.path.join(OUT, f"ncu_full_{tag}.json"), "w") as f:
            json.dump(summary, f, indent=1)
        lines = [f"# ncu --set full captures ({tag}), key metrics per captured launch", ""]
        for k, rep in summary.items():
            lines.append(f"## {k}")
            for d in rep:
                lines.append("")
                lines.append(f"* `{d['kernel']}`")
                for m in RAW:
                    if m in d:
                        lines.append(f"  * {m}: {d[m]}")
            lines.append("")
        with open(os.path.join(OUT, f"ncu_full_{tag}.md"), "w") as f:
            f.write("\n".join(lines))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
