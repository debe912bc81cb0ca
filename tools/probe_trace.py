"""Where does a GEMM / conv launch spend its time?  Per-CTA clock64 stamps recorded by the kernel itself (bring-up hook
hcp_debug_gemm_trace in gemm.cu):  python tools/probe_trace.py conv 4 8 1280 1280 | gemm 256 1280 1280
Slots: 0 start, 1 first TMA issued, 2 last TMA issued, 3 first full barrier seen by the MMA thread, 4 last commit issued,
5 epilogue saw the first accumulator, 6 epilogue done, 7 globaltimer at start; 8 MMA cycles waiting on full barriers,
9 producer cycles waiting on empty barriers, 10 MMA cycles waiting for a drained accumulator, 11 k-blocks loaded."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hcp_diffusion_b200.models  # noqa: E402,F401
from hcp_diffusion_b200 import _lib, ops  # noqa: E402
from hcp_diffusion_b200.ops import ConvPack, LinearPack  # noqa: E402

BF = torch.bfloat16
kind, args = sys.argv[1], [int(a) for a in sys.argv[2:]]
if kind == "conv":
    B, H, Cin, Cout = args
    pack = ConvPack(torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin), torch.randn(Cout, device="cuda") * 0.1, 1)
    x = torch.randn(B, H * H, Cin, device="cuda").to(BF)
    run = lambda: ops.conv3x3(pack, x, (B, H, H))      # noqa: E731
else:
    M, K, N = args
    pack = LinearPack(torch.randn(N, K, device="cuda") / math.sqrt(K), torch.randn(N, device="cuda") * 0.1)
    x = torch.randn(M, K, device="cuda").to(BF)
    run = lambda: ops.fused_linear(pack, [x])          # noqa: E731
lib = _lib.lib()
lib.hcp_debug_gemm_trace.argtypes = [C.c_void_p]
with torch.no_grad():
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(296 * 16, dtype=torch.int64, device="cuda")
    lib.hcp_debug_gemm_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    lib.hcp_debug_gemm_trace(None)
t = buf.view(296, 16).cpu()
t = t[t[:, 0] != 0]
g0 = int(t[:, 7].min())
print(f"{kind} {args}: {t.shape[0]} CTAs traced, event time {e0.elapsed_time(e1) * 1e3:.1f} us (eager launch, incl. finalize)")
print("cta  start_ns | first_tma last_tma first_full last_commit epi_first epi_done (cycles from CTA start) | wait_full wait_empty wait_acc kblocks")
for i in list(range(min(6, t.shape[0]))) + list(range(max(6, t.shape[0] - 3), t.shape[0])):
    r = t[i]
    rel = [int(r[j] - r[0]) if r[j] else -1 for j in range(1, 7)]
    epi = [int(r[j] - r[0]) if r[j] else -1 for j in range(12, 16)]
    print(f"{i:3d} {int(r[7]) - g0:9d} | " + " ".join(f"{v:9d}" for v in rel) + f" | {int(r[8]):9d} {int(r[9]):9d} {int(r[10]):9d} {int(r[11]):5d}"
          f" | part0: ld {epi[0]} conv {epi[1]} sync {epi[2]} stored {epi[3]}")
span = int((t[:, 7].max() - g0))
print(f"start skew over CTAs {span} ns; median epi_done {int((t[:, 6] - t[:, 0]).median())} cycles; median wait_full {int(t[:, 8].median())}, wait_empty {int(t[:, 9].median())}")
