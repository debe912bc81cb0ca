"""Forward-only reuse of the UNet hot path: the classifier-free-guidance denoising loop of the reference's text-to-image pipe
(hcpdiff/utils/pipe_hook.py:115-150, `HookPipe_T2I.__call__` step 7) around `UNet2DConditionModel`, latent in -> latent out.

The text encoders and the VAE are outside the hot path: the caller supplies prompt / negative-prompt embeddings
(and, for SDXL, `added_cond_kwargs`) and decodes the returned latent itself.  Per step the UNet runs ONCE on the doubled batch
[negative | positive] (the order the reference concatenates them in, and the order DreamArtist++ adapters expect); with
`use_cuda_graph` that forward is captured once and replayed, so a step is one graph launch plus a handful of elementwise ops on the
4-channel latent.

Scheduler: DDIM with eta = 0, `leading` timestep spacing and `steps_offset = 1`, `set_alpha_to_one = False` -- the scheduler config
Stable Diffusion 1.x ships with; the beta schedule is the training one (`engine.ddpm_alphas_cumprod`).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .engine import ddpm_alphas_cumprod


def ddim_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, steps_offset: int = 1) -> torch.Tensor:
    ratio = num_train_timesteps // num_inference_steps
    return (torch.arange(0, num_inference_steps) * ratio).flip(0) + steps_offset


class CFGDenoiser:
    def __init__(self, unet: nn.Module, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
        self.unet = unet
        self.T = num_train_timesteps
        self.acp = ddpm_alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self._graph = None
        self._static = None

    def _eps(self, x2: torch.Tensor, t: torch.Tensor, ehs2: torch.Tensor, added: Optional[Dict[str, torch.Tensor]]) -> torch.Tensor:
        if added is not None:
            return self.unet(x2, t, ehs2, added_cond_kwargs=added).sample
        return self.unet(x2, t, ehs2).sample

    @torch.no_grad()
    def sample(self, latents: torch.Tensor, prompt_embeds: torch.Tensor, negative_embeds: torch.Tensor, num_inference_steps: int = 30,
               guidance_scale: float = 7.5, added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
               use_cuda_graph: bool = True) -> torch.Tensor:
        """latents [B,4,H,W] ~ N(0,1) on the UNet's device; prompt / negative embeddings [B,L,ctx]; `added_cond_kwargs` already
        doubled ([negative | positive]) like the reference does for `crop_info` (pipe_hook.py:113-114).  Returns x_0 latents."""
        dev = latents.device
        B = latents.shape[0]
        acp = self.acp.to(dev)
        x = latents.float().clone()
        ehs2 = torch.cat([negative_embeds, prompt_embeds], 0).float().contiguous()
        x2 = torch.empty((2 * B, *x.shape[1:]), dtype=torch.float32, device=dev)
        t_dev = torch.zeros(2 * B, dtype=torch.int64, device=dev)
        added = None if added_cond_kwargs is None else {k: v.to(dev).float().contiguous() for k, v in added_cond_kwargs.items()}
        graph, eps2 = None, None
        if use_cuda_graph:
            x2.copy_(torch.cat([x, x], 0))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._eps(x2, t_dev, ehs2, added)           # warm-up: builds packed weights / tensor maps outside the capture
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eps2 = self._eps(x2, t_dev, ehs2, added)
        steps = ddim_timesteps(num_inference_steps, self.T)
        for i, t in enumerate(steps.tolist()):
            t = min(t, self.T - 1)
            x2.copy_(torch.cat([x, x], 0))
            t_dev.fill_(t)
            if graph is not None:
                graph.replay()
            else:
                eps2 = self._eps(x2, t_dev, ehs2, added)
            e_u, e_c = eps2.chunk(2)
            eps = e_u + guidance_scale * (e_c - e_u)
            a_t = acp[t]
            t_prev = t - self.T // num_inference_steps
            a_prev = acp[t_prev] if t_prev >= 0 else acp[0]
            x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
            x = a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps
        return x
