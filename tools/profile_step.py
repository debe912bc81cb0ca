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
groups, lora = make_hcpdiff(unet, None, [{"rank": 8, "dropout": 0.0, "layers": [r"re:.*\.attn.?$"]}])
step = LoraTrainStep(unet, [p for g in groups for p in g["params"]], use_cuda_graph=False)
lat, noise = torch.randn(B, 4, 64, 64), torch.randn(B, 4, 64, 64)
t, ehs = torch.randint(0, 1000, (B,)), torch.randn(B, 77, 768)
for _ in range(2):
    step.step(lat, noise, t, ehs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.step(lat, noise, t, ehs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(step.loss.cpu()))
