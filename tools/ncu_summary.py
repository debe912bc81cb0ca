"""Compact summary of an ncu report for profiles/:  python tools/ncu_summary.py report.ncu-rep out.summary.csv
Keeps the raw-page columns the design discussion uses (duration, DRAM / L2 / crossbar bytes and throughput, tensor / XU / FMA pipe
utilisation, issue slots, launch geometry) as header / unit / value rows -- the format bench.py reads (`roofline.traffic`,
`attn_tensor_pipe_pct`)."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit = rows[0], rows[1]
KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration", "dram__", "lts__throughput", "lts__t_bytes.sum", "lts__t_sector_hit_rate",
        "l1tex__m_xbar2l1tex_read_bytes", "sm__pipe_tensor", "smsp__pipe_tensor", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fma",
        "sm__inst_executed_pipe_tensor", "sm__issue_active", "smsp__issue_active", "smsp__inst_executed.sum", "sm__throughput", "sm__cycles_active",
        "sm__warps_active", "launch__")
cols = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in cols])
    w.writerow([unit[i] for i in cols])
    for r in rows[2:]:
        if len(r) == len(hdr):
            w.writerow([r[i] for i in cols])
print(f"{out}: {len(cols)} columns, {len(rows) - 2} kernel(s)")
