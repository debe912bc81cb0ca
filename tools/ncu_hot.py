"""Top SASS instructions by warp-stall samples of an ncu report:  python tools/ncu_hot.py report.ncu-rep [n]
(reads `ncu -i report --page source --csv`; prints sample share, dominant stall reasons and the instruction)."""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
si = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[start + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body) or 1
agg = {}
for i, h in stall_cols:
    agg[h] = sum(int(r[i] or 0) for r in body)
print("total samples", tot, " stall mix:", ", ".join(f"{h[6:]} {100 * v / tot:.0f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:n]:
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print(f"{100 * int(r[si]) / tot:5.1f}%  #{idx:5d}  {', '.join(f'{h}:{v}' for v, h in st if v):28s} {r[1].strip()[:110]}")
