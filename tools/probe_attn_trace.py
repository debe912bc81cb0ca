"""Where does a q tile of the attention backward go?  (hcp_debug_attn_trace: cycle sums per phase of the lean softmax loop,
thread 0 of CTA 0.)  Needs a library built with the trace compiled in:
    HCP_EXTRA_NVCC_FLAGS=-DHCP_ATTN_TRACE python -c "import hcp_diffusion_b200 as p, os; os.remove(p._lib.LIB_PATH) if os.path.exists(p._lib.LIB_PATH) else None; p.build()"
    python tools/probe_attn_trace.py [B H L d]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hcp_diffusion_b200.models  # noqa: E402,F401
from hcp_diffusion_b200 import _lib, ops  # noqa: E402

B, H, L, d = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (4, 8, 4096, 40)
Cc = H * d
qkv = (torch.randn(B, L, 3 * Cc, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
o = ops.attention(H, Cc, (0, Cc, 2 * Cc), qkv, None, None)
do = torch.randn_like(o)
o.backward(do, retain_graph=True)
torch.cuda.synchronize()
lib = _lib.lib()
lib.hcp_debug_attn_trace.argtypes = [C.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.hcp_debug_attn_trace(buf.data_ptr())
qkv.grad = None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
o.backward(do)
e1.record()
torch.cuda.synchronize()
lib.hcp_debug_attn_trace(None)
t = buf.cpu().tolist()
names = ["wait S/dP", "exp (ld S, ex2, pack)", "wait dK/dQ(i-1)", "P store+fence+arrive", "dS (ld dP, fma, hmul2)", "dS store+fence+arrive", "dQ drain(i-1)"]
nq = max(t[1], 1)
print(f"attention bwd B{B} H{H} L{L} d{d}: {e0.elapsed_time(e1) * 1e3:.1f} us (eager, incl. prep/post); CTA 0: {t[0]} cycles for {nq} q tiles = {t[0] / nq:.0f} per tile")
for n, v in zip(names, t[2:9]):
    print(f"  {n:28s} {v / nq:8.0f} cycles / tile  ({100.0 * v / max(t[0], 1):4.1f} %)")
