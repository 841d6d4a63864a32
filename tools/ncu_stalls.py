"""Summarise the warp-stall sampling of an `ncu --set full --import-source on` report (runs in the CPU container):
top SASS instructions by samples with their dominant stall reasons, and stall-reason totals.
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
r = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True)
lines = r.stdout.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[0][:160])
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: 0 for c in stall_cols}
total_samples = 0
recs = []
for i, row in enumerate(rows):
    try:
        n = int(row["# Samples"] or 0)
    except ValueError:
        n = 0
    total_samples += n
    st = {}
    for c in stall_cols:
        try:
            v = int(row[c] or 0)
        except ValueError:
            v = 0
        tot[c] += v
        if v:
            st[c[6:]] = v
    recs.append((n, i, row["Source"], st, row.get("Instructions Executed", "")))
print("total samples", total_samples)
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(1,total_samples):.1f}%)" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
for n, i, src, st, ex in sorted(recs, key=lambda r: -r[0])[:top]:
    s = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:6d} {100*n/max(1,total_samples):5.1f}%  #{i:4d} ex={ex:>8}  {src[:70]:70s} | {s}")
