"""Top stall sites of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).  usage: ncu_stalls.py rep [launch#] [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
print(rows[0][1][:80] if rows[0] else "")
si, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) > si and r[si].isdigit():
        data.append((int(r[si]), r[src], r))
tot = sum(d[0] for d in data) or 1
agg = {}
for n, s, r in data:
    for i in cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("samples", tot, sorted(agg.items(), key=lambda x: -x[1])[:7])
for n, s, r in sorted(data, key=lambda x: -x[0])[:topn]:
    extra = {hdr[i][6:]: r[i] for i in cols if r[i] not in ("0", "")}
    top = sorted(extra.items(), key=lambda x: -int(x[1]))[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}% ex={r[ex]:>9s} {s.strip()[:62]:62s} {top}")
