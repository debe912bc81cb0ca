"""Worker of tests/test_gpu_more.py::test_two_rank_nccl_step_equals_single_rank_on_concatenated_batch (launched by torchrun, 2 ranks).

Each rank builds the TINY UNet + LoRA with a DIFFERENT seed on purpose (the engine's sync_params must make the replicas equal, like
DDP's construction-time broadcast), steps 3 times on its own half of a 2B batch; rank 0 then repeats the run alone on the
concatenated batch.  Results go to <out>/rank{r}.pt and <out>/single.pt."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hcp_diffusion_b200.engine import LoraTrainStep  # noqa: E402
from hcp_diffusion_b200.models import UNet2DConditionModel  # noqa: E402
from hcp_diffusion_b200.utils.cfg_net_tools import make_hcpdiff  # noqa: E402
from oracle import unet_ref as U  # noqa: E402


def build(seed, dev):
    spec = U.TINY
    unet = UNet2DConditionModel(sample_size=spec.sample_size, block_out_channels=spec.block_out_channels,
                                attention_head_dim=spec.num_heads, cross_attention_dim=spec.cross_attention_dim)
    unet.load_state_dict(U.init_params(spec))
    unet = unet.to(dev).requires_grad_(False).eval()
    torch.manual_seed(seed)
    groups, group = make_hcpdiff(unet, None, [{"rank": 4, "lr": 1e-3, "layers": [r"re:.*\.attn.?$"]}])
    for blk in group.plugin_dict.values():
        torch.nn.init.normal_(blk.layer.W_up, std=0.02)
    return unet, groups


def run(step, batches):
    losses = []
    init = step.flat.data.clone()
    for lat, noise, t, ehs in batches:
        losses.append(float(step.step(lat, noise, t, ehs).cpu()))
    return {"params": step.flat.data.cpu(), "m": step.m.cpu(), "init": init.cpu(), "loss": losses}


def main():
    out = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    B = 2
    full = [U.synthetic_batch(B * world, U.TINY, seed=500 + i) for i in range(3)]
    mine = [tuple(x[rank * B:(rank + 1) * B] for x in b) for b in full]
    unet, groups = build(seed=10 + rank, dev=dev)                  # different LoRA init per rank ...
    step = LoraTrainStep(unet, groups, use_cuda_graph=True)
    step.sync_params(src=0)                                        # ... made equal here
    torch.save(run(step, mine), os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        unet, groups = build(seed=10, dev=dev)
        single = LoraTrainStep(unet, groups, use_cuda_graph=True)
        assert single.world == 1
        torch.save(run(single, full), os.path.join(out, "single.pt"))


if __name__ == "__main__":
    main()
