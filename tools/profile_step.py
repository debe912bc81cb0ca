"""One eager (no CUDA graph) LoRA training step of the benchmark workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 1 -o gpurun_out/attn_bwd \
      python tools/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hcp_diffusion_b200.engine import LoraTrainStep  # noqa: E402
from hcp_diffusion_b200.models import UNet2DConditionModel  # noqa: E402
from hcp_diffusion_b200.utils.cfg_net_tools import make_hcpdiff  # noqa: E402

B = int(os.environ.get("HCP_BATCH", "4"))
torch.manual_seed(0)
unet = UNet2DConditionModel().cuda().requires_grad_(False).eval()      # random init: timing only
if os.environ.get("HCP_PROFILE_CONFIG", "2") == "3":      # BASELINE config 3: full fine-tune (run with HCP_BATCH=16)
    groups, lora = make_hcpdiff(unet, [{"lr": 1e-6, "layers": [""]}], None)
else:
    groups, lora = make_hcpdiff(unet, None, [{"rank": 8, "dropout": 0.0, "layers": [r"re:.*\.attn.?$"]}])
step = LoraTrainStep(unet, [p for g in groups for p in g["params"]], use_cuda_graph=False)
lat, noise = torch.randn(B, 4, 64, 64), torch.randn(B, 4, 64, 64)
t, ehs = torch.randint(0, 1000, (B,)), torch.randn(B, 77, 768)
for _ in range(2):
    step.step(lat, noise, t, ehs)

# log the shape of every C-ABI call of the profiled step, in launch order (joined with the ncu launch list offline)
import ctypes, json
from hcp_diffusion_b200 import _lib
calls = []
_orig_call = _lib.call


def _logged(name, *args):
    info = {"fn": name}
    try:
        a0 = args[0]
        obj = a0._obj if hasattr(a0, "_obj") else None
        if name == "hcp_gemm_bf16":
            info.update(M=obj.M, N=obj.N, K=[obj.k[i] for i in range(obj.nseg)], split=bool(obj.workspace))
        elif name == "hcp_conv3x3_bf16":
            info.update(B=obj.B, H=obj.Hin, W=obj.Win, Cin=obj.Cin, Cout=obj.Cout, stride=obj.stride, mode=obj.mode, split=bool(obj.workspace))
        elif name in ("hcp_attn_fwd_bf16", "hcp_attn_bwd_bf16"):
            info.update(B=obj.B, H=obj.H, Lq=obj.Lq, Lkv=obj.Lkv, d=obj.d)
        elif name in ("hcp_groupnorm_fwd_bf16", "hcp_groupnorm_bwd_bf16"):
            info.update(B=obj.B, HW=obj.HW, C=obj.C1 + obj.C2)
    except Exception as e:  # noqa: BLE001
        info["err"] = str(e)
    calls.append(info)
    return _orig_call(name, *args)


_lib.call = _logged
import hcp_diffusion_b200.ops as _ops, hcp_diffusion_b200.engine as _eng, hcp_diffusion_b200.runtime as _rt
_ops.call = _logged; _eng.call = _logged
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.step(lat, noise, t, ehs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(calls, open("gpurun_out/launch_shapes.json", "w"))
print("loss", float(step.loss.cpu()), "calls", len(calls))
