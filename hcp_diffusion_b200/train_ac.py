"""`python -m hcp_diffusion_b200.train_ac --cfg cfgs/train/lora_sd15_synthetic.yaml key=value ...`

The reference entrypoint (hcpdiff/train_ac.py:559-566: `load_config_with_cli` -> `Trainer(conf)` -> `trainer.train()`), reduced to
the hot path: build the UNet (cfg `model.unet`, the reference's injection seam train_ac.py:220), apply `lora_unet` through
`make_hcpdiff` (train_ac.py:324-359), then run `train.train_steps` LoRA steps with the B200 engine and save
`ckpts/unet-<step>.safetensors` every `train.save_step` in the reference checkpoint format (train_ac.py:523-544).

Out of the hot path and therefore NOT here: datasets / buckets / captions, the CLIP text encoder, VAE, EMA, loggers,
DeepSpeed / Colossal-AI trainers.  Inputs are the synthetic latents / text embeddings of SURVEY.md 8d (`data.synthetic`), or
tensors saved in a .pt file (`data.path`: {'latents': [N,4,h,w], 'encoder_hidden_states': [N,L,768]}).

Launch data-parallel with torchrun (one process per GPU); gradients are all-reduced over NCCL, the learning rate is
scaled by batch x world x accumulation when `train.scale_lr` is set (reference train_ac.py:192-197).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

from .ckpt_manager import CkptManagerPKL, CkptManagerSafe
from .engine import LoraTrainStep
from .utils.cfg_net_tools import HCPModelLoader, make_hcpdiff
from .utils.config import instantiate, load_config_with_cli


class Trainer:
    def __init__(self, cfgs):
        self.cfgs = cfgs
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
        torch.manual_seed(int(cfgs.get("seed", 114514)) + self.local_rank)          # reference train_ac.py:128

        unet = cfgs.model.get("unet")
        unet = instantiate(unet) if isinstance(unet, dict) else unet
        if unet is None:
            raise ValueError("cfg `model.unet` must instantiate a UNet (e.g. _target_: hcp_diffusion_b200.models.UNet2DConditionModel)")
        init = cfgs.model.get("init")
        if init and init != "random":            # "random"/absent: keep the constructor's initialisation (no weights on disk here)
            sd = CkptManagerSafe().load_ckpt(init) if init.endswith(".safetensors") else torch.load(init, map_location="cpu")
            unet.load_state_dict(sd.get("base", sd), strict=False)
        self.unet = unet.to(self.device).requires_grad_(False).eval()
        self.unet.enable_xformers_memory_efficient_attention()                      # no-ops kept for config compatibility
        if cfgs.model.get("gradient_checkpointing", False):
            self.unet.enable_gradient_checkpointing()

        tr = cfgs.train
        bs = int(cfgs.data.get("batch_size", 4))
        lr_scale = bs * self.world * int(tr.get("gradient_accumulation_steps", 1)) if tr.get("scale_lr", False) else 1
        if cfgs.get("unet"):
            raise NotImplementedError("`unet:` full-layer training needs the wgrad kernels, which are not built yet")
        groups, self.lora = make_hcpdiff(self.unet, None, cfgs.get("lora_unet"), default_lr=float(tr.optimizer.get("lr", 1e-4)))
        resume = tr.get("resume")
        if resume and resume.get("ckpt_path", {}).get("unet"):
            HCPModelLoader(self.unet).load_lora([{"path": p, "alpha": 1.0} for p in resume.ckpt_path.unet])
        params = [p for g in groups for p in g["params"]]
        lr = float(groups[0]["lr"]) * lr_scale
        self.step_fn = LoraTrainStep(self.unet, params, lr=lr, weight_decay=float(tr.optimizer.get("weight_decay", 1e-2)),
                                     max_grad_norm=float(tr.get("max_grad_norm", 1.0)), use_cuda_graph=bool(tr.get("cuda_graph", True)))
        self.bs = bs
        self.ckpt = CkptManagerSafe() if cfgs.get("ckpt_type", "safetensors") == "safetensors" else CkptManagerPKL()
        self.exp_dir = cfgs.get("exp_dir", "exps/run")
        if self.rank == 0:
            self.ckpt.set_save_dir(os.path.join(self.exp_dir, "ckpts"))
        self._load_data()

    def _load_data(self):
        d = self.cfgs.data
        g = torch.Generator().manual_seed(1234 + self.rank)
        if d.get("path"):
            blob = torch.load(d.path, map_location="cpu")
            self.latents, self.ehs = blob["latents"].float(), blob["encoder_hidden_states"].float()
        else:
            n = int(d.get("num_samples", 64))
            s = int(self.unet.config.sample_size)
            self.latents = torch.randn((n, self.unet.config.in_channels, s, s), generator=g)
            self.ehs = torch.randn((n, int(d.get("tokens", 77)), self.unet.config.cross_attention_dim), generator=g)
        self.gen = g

    def next_batch(self):
        idx = torch.randint(0, self.latents.shape[0], (self.bs,), generator=self.gen)
        lat, ehs = self.latents[idx], self.ehs[idx]
        noise = torch.randn(lat.shape, generator=self.gen)
        t = torch.randint(0, 1000, (self.bs,), generator=self.gen, dtype=torch.int64)
        return [x.pin_memory() for x in (lat, noise, t, ehs)]

    def train(self):
        tr = self.cfgs.train
        steps, save_step, log_step = int(tr.train_steps), int(tr.get("save_step", 0)), int(tr.get("log_step", 20))
        t0, seen = time.time(), 0
        for step in range(1, steps + 1):
            loss = self.step_fn.step(*self.next_batch())
            seen += self.bs * self.world
            if step % log_step == 0 or step == steps:
                val = float(loss.cpu())
                if self.rank == 0:
                    print(f"step {step}/{steps}  loss {val:.5f}  {seen / (time.time() - t0):.1f} img/s", flush=True)
                t0, seen = time.time(), 0
            if save_step and step % save_step == 0 and self.rank == 0:
                self.ckpt.save_model_with_lora(None, self.lora, "unet", step)
        if self.world > 1:
            dist.barrier()


def main(argv=None):
    ap = argparse.ArgumentParser(description="HCP-Diffusion LoRA training on the B200 hot path")
    ap.add_argument("--cfg", type=str, required=True)
    args, overrides = ap.parse_known_args(argv)
    conf = load_config_with_cli(args.cfg, args_list=overrides)
    Trainer(conf).train()


if __name__ == "__main__":
    main(sys.argv[1:])
