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
