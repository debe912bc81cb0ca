"""Join the C-ABI call log of tools/profile_step.py (gpurun_out/launch_shapes.json) with an ncu launch list (csv with
gpu__time_duration.sum) and print time per call shape -- which GEMM / conv / attention shapes the step spends its time in."""
import collections
import csv
import json
import re
import sys

calls = json.load(open(sys.argv[1]))
rows = [l for l in open(sys.argv[2]) if not l.startswith("==")]
kern = [(re.sub(r"\(.*", "", r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(rows)
        if "hcp::" in r["Kernel Name"]]


def expand(c):
    f = c["fn"]
    if f == "hcp_gemm_bf16":
        return 2 if c.get("split") else 1
    if f == "hcp_conv3x3_bf16":
        return 4 if c.get("mode") == 1 else (2 if c.get("split") else 1)
    if f == "hcp_attn_bwd_bf16":
        return 4 if c.get("d", 0) > 128 else 3
    if f in ("hcp_groupnorm_fwd_bf16", "hcp_groupnorm_bwd_bf16", "hcp_adamw_flat"):
        return 2
    return 1


agg = collections.defaultdict(lambda: [0, 0.0])
i = 0
for c in calls:
    n = expand(c)
    t = sum(k[1] for k in kern[i:i + n])
    i += n
    key = json.dumps({k: v for k, v in c.items()}, sort_keys=True)
    agg[key][0] += 1
    agg[key][1] += t
print("kernels consumed", i, "of", len(kern))
tot = sum(v[1] for v in agg.values())
print("total us", round(tot))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    c = json.loads(k)
    fl = None
    if c["fn"] == "hcp_gemm_bf16":
        fl = 2.0 * c["M"] * c["N"] * sum(c["K"])
    elif c["fn"] == "hcp_conv3x3_bf16":
        s = c["stride"]
        fl = 2.0 * c["B"] * c["H"] * c["W"] * c["Cout"] * 9 * c["Cin"] / (s * s if c["mode"] == 0 else 1) * (1 if c["mode"] == 0 else 1)
    elif c["fn"].startswith("hcp_attn"):
        fl = (4.0 if "fwd" in c["fn"] else 10.0) * c["B"] * c["H"] * c["Lq"] * c["Lkv"] * c["d"]
    tf = f"{fl / (t / n) * 1e-6:7.0f} TF/s" if fl else " " * 12
    print(f"{t:9.1f} us {100 * t / tot:5.1f}%  n={n:3d} avg={t / n:7.1f} {tf}  {k[:150]}")
