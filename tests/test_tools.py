"""Host-side tooling around the measurements (CPU only): the launch-list join and the roofline traffic lookup of bench.py must keep
working on the committed profiles, otherwise the numbers DESIGN.md quotes cannot be regenerated."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def test_join_launches_consumes_every_kernel_of_the_committed_launch_list():
    shapes = os.path.join(PROFILES, "r01_launch_shapes_v23.json")
    launches = os.path.join(PROFILES, "r01_ncu_launches_step_v23.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "join_launches.py"), shapes, launches],
                         capture_output=True, text=True, check=True).stdout
    first = out.splitlines()[0].split()                      # "kernels consumed N of N"
    assert first[:2] == ["kernels", "consumed"] and first[2] == first[4], out[:200]
    calls = json.load(open(shapes))
    assert sum(c["fn"] == "hcp_attn_fwd_bf16" for c in calls) == 32      # 16 transformer blocks x (self + cross)
    assert "hcp_attn_bwd_bf16" in out and "hcp_conv3x3_bf16" in out


def test_roofline_traffic_reads_the_committed_ncu_summary():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_dram_traffic()
    assert t is not None and t["kernel"].startswith("attn_bwd_kernel")
    # q, k, v, o, dO in and dq, dk, dv out are 8 x 10.5 MB of algorithmic bf16 traffic at B4 H8 L4096 d40; measured DRAM bytes of one
    # launch must be the same order (fp32 dQ partials stay in L2)
    assert 40e6 < t["bytes_per_launch"] < 160e6
    assert bench.ncu_dram_traffic("profiles/does_not_exist.csv") is None


def test_round2_profiles_regenerate_and_feed_the_bench_line():
    """The round-2 launch list joins with its call log (17 rank-projection launches left of round 1's 121), the
    per-kernel totals tool runs on it, and bench.py picks the round-2 `ncu --set full` summaries for `roofline.traffic` and
    `attn_tensor_pipe_pct` (SM-average tensor-pipe utilisation over the kernel's duration)."""
    shapes = os.path.join(PROFILES, "r02_launch_shapes_final.json")
    launches = os.path.join(PROFILES, "r02_ncu_launches_step_final.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "join_launches.py"), shapes, launches],
                         capture_output=True, text=True, check=True).stdout
    first = out.splitlines()[0].split()
    assert first[:2] == ["kernels", "consumed"] and int(first[4]) - int(first[2]) <= 1, out[:200]
    calls = json.load(open(shapes))
    # T / U ride the layers' own GEMMs: the only rank-projection launches left are U of the layers whose input carries no gradient
    # (text-embedding k/v projections, the very first qkv) -- 17 instead of round 1's 121
    skinny = [c for c in calls if c["fn"] == "hcp_gemm_bf16" and c.get("N") == 64]
    assert len(skinny) <= 17 and all(c["M"] in (308, 16384) for c in skinny)
    assert any(c["fn"] == "hcp_lora_merge" for c in calls)
    tot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_totals.py"), launches, "5"],
                         capture_output=True, text=True, check=True).stdout
    assert tot.startswith("total ") and "attn_bwd_kernel" in tot
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_dram_traffic()
    assert t["source"].startswith("profiles/r02_") and 40e6 < t["bytes_per_launch"] < 160e6
    pct = bench.attn_tensor_pipe_pct()
    assert pct["fwd"]["source"].startswith("profiles/r02_") and 10.0 < pct["fwd"]["pct"] < 60.0
    assert pct["bwd"]["metric"] == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" and 15.0 < pct["bwd"]["pct"] < 60.0
