"""One launch of a GEMM / conv between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full --import-source on`:
  python tools/probe_one.py gemm M K N [bias] [res]   |   python tools/probe_one.py conv B H Cin Cout"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hcp_diffusion_b200.models  # noqa: E402,F401
from hcp_diffusion_b200 import ops  # noqa: E402
from hcp_diffusion_b200.ops import ConvPack, LinearPack  # noqa: E402

BF = torch.bfloat16
kind = sys.argv[1]
nums = [int(a) for a in sys.argv[2:] if a.isdigit()]
flags = [a for a in sys.argv[2:] if not a.isdigit()]
if kind == "conv":
    B, H, Cin, Cout = nums
    pack = ConvPack(torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin), torch.randn(Cout, device="cuda") * 0.1, 1)
    x = torch.randn(B, H * H, Cin, device="cuda").to(BF)
    run = lambda: ops.conv3x3(pack, x, (B, H, H))      # noqa: E731
else:
    M, K, N = nums
    pack = LinearPack(torch.randn(N, K, device="cuda") / math.sqrt(K), torch.randn(N, device="cuda") * 0.1 if "bias" in flags else None)
    x = torch.randn(M, K, device="cuda").to(BF)
    res = torch.randn(M, N, device="cuda").to(BF) if "res" in flags else None
    run = lambda: ops.fused_linear(pack, [x], res)     # noqa: E731
with torch.no_grad():
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
