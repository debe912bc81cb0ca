"""The LoRA training step around the UNet call, B200-native.

What the reference does per step (hcpdiff/train_ac.py:467-504): H2D copies, `make_noise` (:437-447), UNet forward through
accelerate's autocast + DDP, MSE loss in fp32 (:506-515), `accelerator.backward` (DDP bucketed all-reduce of the LoRA
gradients), `clip_grad_norm_` (:485-490), AdamW, `zero_grad`, and a `loss.item()` sync.

Here: every trainable LoRA tensor is a view into ONE flat fp32 buffer (same for gradients and Adam moments), so
  * zero_grad is one memset, clip + AdamW are two kernels over the flat buffer (no host sync: the clip factor is computed
    on the device),
  * the data-parallel exchange is ONE NCCL all-reduce of the flat gradient (6.4 MB for rank-8 attention LoRA) over
    NVLink/NVSwitch -- the only collective in the job, exactly as the reference's DDP,
  * forward + loss + backward is captured once in a CUDA graph and replayed (about 1.5k kernel launches per step).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from . import _lib, ops
from ._lib import call, stream_ptr


def ddpm_alphas_cumprod(num_steps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    """SD1.5 DDPMScheduler(beta_schedule='scaled_linear') alphas_cumprod (constants as in reference tools/gen_from_ptlist.py:14-16)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class FlatParams:
    """Re-homes a list of parameters into one contiguous fp32 buffer (+ gradient buffer); names/shapes are untouched, so
    state_dict(), the optimizer param groups and hcpdiff's checkpoint code keep working."""

    def __init__(self, params: Sequence[nn.Parameter]):
        params = [p for p in dict.fromkeys(params)]
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        self.params = params
        self.offsets, n = [], 0
        for p in params:
            if p.dtype != torch.float32:
                raise TypeError("trainable LoRA parameters must be fp32 master weights")
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4          # keep every tensor 16-byte aligned inside the flat buffer
        self.numel = n
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, self.offsets):
            view = self.data[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad.zero_()


class LoraTrainStep:
    """One optimisation step of LoRA training on a `UNet2DConditionModel` (eps-prediction MSE, AdamW, grad-norm clip)."""

    def __init__(self, unet: nn.Module, params: Iterable[nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 1.0, use_cuda_graph: bool = True,
                 process_group: Optional[dist.ProcessGroup] = None, side_stream: bool = True):
        self.unet = unet
        self.flat = FlatParams(list(params))
        dev = self.flat.data.device
        self.m = torch.zeros_like(self.flat.data)
        self.v = torch.zeros_like(self.flat.data)
        self.lr = torch.tensor([lr], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.betas, self.eps, self.wd, self.max_norm = betas, eps, weight_decay, max_grad_norm
        self.acp = ddpm_alphas_cumprod().to(dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.gsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.use_graph = use_cuda_graph
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self._static = None
        self._graph_fb = None
        self._graph_opt = None
        # work off the critical path (LoRA-gradient kernels, text-embedding k/v projections) goes to a side stream inside
        # _forward_backward and is joined there, before anything reads the gradients (HCP_SIDE_STREAM=0 keeps a single stream)
        self.side_stream = side_stream

    def set_lr(self, lr: float):
        self.lr.fill_(lr)

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def _forward_backward(self, latents, noise, t, ehs, added=None):
        """zero_grad, x_t = add_noise, pred = unet(x_t, t, ehs), loss = mse(pred, noise), backward."""
        self.flat.grad.zero_()
        self.loss.zero_()
        ops.set_side_stream(self.side_stream)
        try:
            self._fb_body(latents, noise, t, ehs, added)
        finally:
            ops.join_side()
            ops.set_side_stream(False)

    def _fb_body(self, latents, noise, t, ehs, added=None):
        B = latents.shape[0]
        per_image = latents[0].numel()
        x_t = torch.empty_like(latents)
        call("hcp_add_noise", latents.data_ptr(), noise.data_ptr(), t.data_ptr(), self.acp.data_ptr(), B, per_image, x_t.data_ptr(),
             stream_ptr())
        pred = (self.unet(x_t, t, ehs, added_cond_kwargs=added) if added is not None else self.unet(x_t, t, ehs)).sample
        dpred = torch.empty_like(pred)
        call("hcp_mse_loss", pred.data_ptr(), noise.data_ptr(), pred.numel(), 1.0, self.loss.data_ptr(), dpred.data_ptr(), stream_ptr())
        pred.backward(dpred)

    def _optimizer(self):
        self.gsq.zero_()
        n = self.flat.numel
        call("hcp_sumsq", self.flat.grad.data_ptr(), n, self.gsq.data_ptr(), stream_ptr())
        call("hcp_adamw_flat", self.flat.data.data_ptr(), self.flat.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), n,
             self.lr.data_ptr(), self.betas[0], self.betas[1], self.eps, self.wd, 1.0 / self.world, self.gsq.data_ptr(),
             float(self.max_norm or 0.0), self.step_count.data_ptr(), stream_ptr())

    def _all_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.pg)   # averaged by grad_scale = 1/world in AdamW

    # ---- public ------------------------------------------------------------------------------------------------------
    def step(self, latents: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, ehs: torch.Tensor,
             added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """latents/noise fp32 [B,4,H,W], t int64 [B], ehs fp32 [B,L,ctx] (host-pinned or device); `added_cond_kwargs`
        ({'text_embeds' [B,P], 'time_ids' [B,6]}) for SDXL UNets (reference wrapper.py:66).  Returns the device loss tensor
        (shape [1]); reading it is the caller's D2H."""
        dev = self.flat.data.device
        if not self.use_graph:
            added = None if added_cond_kwargs is None else {k: v.to(dev, non_blocking=True) for k, v in added_cond_kwargs.items()}
            self._forward_backward(latents.to(dev, non_blocking=True), noise.to(dev, non_blocking=True), t.to(dev, non_blocking=True),
                                   ehs.to(dev, non_blocking=True), added)
            self._all_reduce()
            self._optimizer()
            return self.loss
        if self._static is None:
            self._capture(latents, noise, t, ehs, added_cond_kwargs)
        s = self._static
        s["latents"].copy_(latents, non_blocking=True)
        s["noise"].copy_(noise, non_blocking=True)
        s["t"].copy_(t, non_blocking=True)
        s["ehs"].copy_(ehs, non_blocking=True)
        if s["added"] is not None:
            for k, v in s["added"].items():
                v.copy_(added_cond_kwargs[k], non_blocking=True)
        self._graph_fb.replay()
        self._all_reduce()
        self._graph_opt.replay()
        return self.loss

    def step_resident(self) -> torch.Tensor:
        """Replay on the inputs already resident in the static device buffers (kernel-only timing in bench.py)."""
        if self._static is None:
            raise RuntimeError("call step() once before step_resident()")
        self._graph_fb.replay()
        self._all_reduce()
        self._graph_opt.replay()
        return self.loss

    def _capture(self, latents, noise, t, ehs, added=None):
        dev = self.flat.data.device
        self._static = {
            "latents": latents.to(dev).float().contiguous().clone(), "noise": noise.to(dev).float().contiguous().clone(),
            "t": t.to(dev).long().contiguous().clone(), "ehs": ehs.to(dev).float().contiguous().clone(),
            "added": None if added is None else {k: v.to(dev).float().contiguous().clone() for k, v in added.items()},
        }
        s = self._static
        # warm-up on a side stream (builds the packed weights, tensor maps, autograd metadata) -- parameters are restored
        saved = (self.flat.data.clone(), self.m.clone(), self.v.clone(), self.step_count.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._forward_backward(s["latents"], s["noise"], s["t"], s["ehs"], s["added"])
                self._optimizer()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.flat.data.copy_(saved[0]); self.m.copy_(saved[1]); self.v.copy_(saved[2]); self.step_count.copy_(saved[3])
        before = _lib.launch_count
        self._graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_fb):
            self._forward_backward(s["latents"], s["noise"], s["t"], s["ehs"], s["added"])
        self._graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_opt):
            self._optimizer()
        self.launches_per_step = _lib.launch_count - before
