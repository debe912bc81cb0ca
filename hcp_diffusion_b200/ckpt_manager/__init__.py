"""LoRA / base checkpoint I/O in hcpdiff's on-disk format (interchangeable with the reference).

File layout (reference hcpdiff/ckpt_manager/ckpt_pkl.py:55-71, ckpt_safetensor.py:18-63):
  * `.ckpt`         torch.save({'base': {...}, 'lora': {'<layer>.___.layer.W_down': t, '<layer>.___.layer.W_up': t,
                                                       '<layer>.___.alpha': t}})
  * `.safetensors`  the same nested dict flattened with ':' -> keys 'lora:<layer>.___.layer.W_down', ...
`auto_manager(path)` picks the manager from the file extension (reference ckpt_manager/__init__.py:4-5).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
from torch import nn

from ..models.plugin import BasePluginBlock, PluginGroup


class CkptManagerPKL:
    def __init__(self, plugin_from_raw: bool = False):
        self.plugin_from_raw = plugin_from_raw
        self.save_dir = "."

    def set_save_dir(self, save_dir: str, emb_dir: Optional[str] = None):
        os.makedirs(save_dir, exist_ok=True)
        self.save_dir = save_dir
        self.emb_dir = emb_dir

    @staticmethod
    def exclude_state(state, key):
        if key is None:
            return state
        return {k: v for k, v in state.items() if key not in k}

    def save_model_with_lora(self, model: Optional[nn.Module], lora_blocks: Optional[PluginGroup], name: str, step: int,
                             model_ema=None, exclude_key=None, ema_state: Optional[Dict[nn.Parameter, torch.Tensor]] = None):
        """Reference ckpt_pkl.py:55-71.  The EMA parts ('base_ema' / 'lora_ema', same keys as 'base' / 'lora') come from
        `ema_state` = {parameter: EMA tensor} (`LoraTrainStep.ema_state()`: the EMA lives in the engine's flat buffer, not in a
        ModelEMA object); buffers (LoRA `alpha`) are copied as they are, like ModelEMA.update does (utils/ema.py:29-31)."""
        sd_model: Dict[str, Dict[str, torch.Tensor]] = {}
        if model is not None:
            sd_model["base"] = self.exclude_state(BasePluginBlock.extract_state_without_plugin(model, trainable=True), exclude_key)
        if lora_blocks is not None and not lora_blocks.empty():
            sd_model["lora"] = lora_blocks.state_dict(model if self.plugin_from_raw else None)
        if model_ema is not None:
            raise NotImplementedError("pass the engine's `ema_state()` instead of a ModelEMA object")
        if ema_state is not None:
            def ema_of(t):
                for p, e in ema_state.items():
                    if p.data_ptr() == t.data_ptr() and p.shape == t.shape:
                        return e
                return t
            if "base" in sd_model:
                sd_model["base_ema"] = {k: ema_of(v) for k, v in sd_model["base"].items()}
            if "lora" in sd_model:
                sd_model["lora_ema"] = {k: ema_of(v) for k, v in sd_model["lora"].items()}
        return self._save_ckpt(sd_model, name, step)

    def _save_ckpt(self, sd_model, name=None, step=None, save_path=None):
        if save_path is None:
            save_path = os.path.join(self.save_dir, f"{name}-{step}.ckpt")
        torch.save({k: {kk: vv.detach().cpu() for kk, vv in v.items()} for k, v in sd_model.items()}, save_path)
        return save_path

    def load_ckpt(self, ckpt_path, map_location="cpu"):
        return torch.load(ckpt_path, map_location=map_location)

    def load_ckpt_to_model(self, model: nn.Module, ckpt_path, model_ema=None):
        sd = self.load_ckpt(ckpt_path)
        for part in ("base", "lora", "plugin"):
            if part in sd:
                model.load_state_dict(sd[part], strict=False)


class CkptManagerSafe(CkptManagerPKL):
    def _save_ckpt(self, sd_model, name=None, step=None, save_path=None):
        from safetensors.torch import save_file
        if save_path is None:
            save_path = os.path.join(self.save_dir, f"{name}-{step}.safetensors")
        flat = {k: v.detach().cpu().contiguous().clone() for k, v in self.unfold_dict(sd_model).items()}
        save_file(flat, save_path)
        return save_path

    def load_ckpt(self, ckpt_path, map_location="cpu"):
        from safetensors import safe_open
        with safe_open(ckpt_path, framework="pt", device=map_location) as f:
            return self.fold_dict(f)

    @staticmethod
    def unfold_dict(data, split_key=":"):
        """{'lora': {'a.___.alpha': t}} -> {'lora:a.___.alpha': t}; lists/tuples are indexed."""
        flat = {}

        def walk(prefix, node):
            for k, v in node.items():
                key = f"{k}" if prefix == "" else f"{prefix}{split_key}{k}"
                if isinstance(v, dict):
                    walk(key, v)
                elif isinstance(v, (list, tuple)):
                    walk(key, dict(enumerate(v)))
                else:
                    flat[key] = v

        walk("", data)
        return flat

    @staticmethod
    def fold_dict(safe_f, split_key=":"):
        out = {}
        for k in safe_f.keys():
            *parents, leaf = k.split(split_key)
            node = out
            for p in parents:
                node = node.setdefault(p, {})
            node[leaf] = safe_f.get_tensor(k)
        return out


def auto_manager(ckpt_path: str):
    return CkptManagerSafe() if ckpt_path.endswith(".safetensors") else CkptManagerPKL()
