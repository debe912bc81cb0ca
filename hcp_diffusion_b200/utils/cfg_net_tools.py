"""cfg -> network surgery for the LoRA hot path: regex layer selection and LoRA injection.

Restates the behaviour of the reference's hcpdiff/utils/cfg_net_tools.py: `get_match_layers` (:30-75, the `re:` / `pre_hook:` /
`cls(...)` name prefixes), `get_lora_rank_and_cls` (:77-88), `make_hcpdiff` (:90-128) and `HCPModelLoader.load_lora`
(:249-292).  `make_plugin` (ControlNet etc.) is outside the hot path and not provided.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Tuple, Union

import torch
from torch import nn

from ..models.lora import LoraBlock, LoraGroup, lora_layer_map
from ..models.plugin import split_module_name


def net_path_join(*args: str) -> str:
    """'.'-join that skips empty components (reference hcpdiff/utils/utils.py:118-119)."""
    return ".".join(a for a in args if a is not None and len(a) > 0)


def _class_matches(class_name: str, block: nn.Module) -> List[str]:
    if type(block).__name__ == class_name:
        return [""]
    return ["." + n for n, m in block.named_modules() if type(m).__name__ == class_name]


def get_match_layers(layers, all_layers: Dict[str, nn.Module], return_metas: bool = False) -> Union[List[str], List[Dict[str, Any]]]:
    """Resolve a yaml `layers:` list against `dict(model.named_modules())`.
    'name' selects that module; 're:<regex>' every module whose name the regex `match`es (anchored at the start);
    'pre_hook:' marks the entry for pre-hook plugins; 'cls(<ClassName>):' narrows to descendants of that class.
    Duplicates are dropped, first occurrence wins, order preserved."""
    found: List[Tuple[str, bool]] = []
    for entry in layers:
        *metas, name = entry.split(":")
        use_re = "re" in metas
        pre_hook = "pre_hook" in metas
        cls_filter = next((m[4:-1] for m in metas if m.startswith("cls(")), None)
        if use_re:
            rx = re.compile(name)
            matched = [k for k in all_layers.keys() if rx.match(k) is not None]
        else:
            matched = [name]
        if cls_filter is not None:
            matched = [layer + suffix for layer in matched for suffix in _class_matches(cls_filter, all_layers[layer])]
        found += [(m, pre_hook) for m in matched]
    seen, out = set(), []
    for layer, pre_hook in found:
        if layer in seen:
            continue
        seen.add(layer)
        out.append({"layer": layer, "pre_hook": pre_hook} if return_metas else layer)
    return out


def get_lora_rank_and_cls(lora_state: Dict[str, torch.Tensor]):
    if "layer.W_down" in lora_state:
        return lora_layer_map["lora"], lora_state["layer.W_down"].shape[0], False
    if "layer.lora_down.weight" in lora_state:      # old format (reference cfg_net_tools.py:78-82, tools/convert_old_lora.py)
        return lora_layer_map["lora"], lora_state["layer.lora_down.weight"].shape[0], True
    raise ValueError("Unknown lora format.")


def _get(item, key, default=None):
    if isinstance(item, dict):
        return item.get(key, default)
    return getattr(item, key, default)


def _items(item):
    return item.items() if isinstance(item, dict) else vars(item).items()


def make_hcpdiff(model: nn.Module, cfg_model, cfg_lora, default_lr: float = 1e-5):
    """Apply the `unet:` (full-layer training) and `lora_unet:` config lists to `model`.
    Returns (optimizer param groups, LoraGroup) like reference cfg_net_tools.py:90-128."""
    named_modules = dict(model.named_modules())
    train_params: List[Dict[str, Any]] = []
    all_lora_blocks: Dict[str, LoraBlock] = {}
    if cfg_model is not None:
        for item in cfg_model:
            group = []
            for layer_name in get_match_layers(_get(item, "layers"), named_modules):
                layer = named_modules[layer_name]
                layer.requires_grad_(True)
                layer.train()
                group.extend(LoraBlock.extract_param_without_lora(layer).values())
            train_params.append({"params": list(dict.fromkeys(group)), "lr": _get(item, "lr", default_lr)})
    if cfg_lora is not None:
        for lora_id, item in enumerate(cfg_lora):
            group = []
            for layer_name in get_match_layers(_get(item, "layers"), named_modules):
                parent_name, host_name = split_module_name(layer_name)
                layer = named_modules[layer_name]
                args = {k: v for k, v in _items(item) if k != "layers"}
                cls = lora_layer_map[args.get("type", "lora")]
                blocks = cls.wrap_model(lora_id, layer, parent_block=named_modules[parent_name], host_name=host_name, **args)
                for k, blk in blocks.items():
                    all_lora_blocks[net_path_join(layer_name, k)] = blk
                    blk.requires_grad_(True)
                    blk.train()
                    group.extend(blk.parameters())
            train_params.append({"params": group, "lr": _get(item, "lr", default_lr)})
    return train_params, LoraGroup(all_lora_blocks)


class HCPModelLoader:
    """Loads `{'lora': {<layer>.___.<key>: tensor}}` checkpoints back into a model (reference cfg_net_tools.py:227-292)."""

    def __init__(self, host: nn.Module):
        self.host = host
        self.named_modules = dict(host.named_modules())

    @torch.no_grad()
    def load_lora(self, cfg, base_model_alpha: float = 1.0, load_ema: bool = False, lora_id_offset: int = 0) -> LoraGroup:
        """cfg: list of {path | state_dict, alpha, alpha_auto_scale, dropout, layers}.  Every checkpoint becomes one more stacked
        LoraBlock per layer, exactly like the reference (cfg_net_tools.py:249-292): the `alpha` STORED in the checkpoint is dropped and
        the block is rebuilt with `item.alpha` (default 1.0) and `alpha_auto_scale` (default True, i.e. alpha / rank); `dropout` comes
        from the item; old-format (`.lora_block.`) keys and `lora_ema` parts are accepted.  `lora_id_offset` (not in the reference)
        shifts the block ids when the model already carries blocks from `make_hcpdiff` -- the reference would collide on
        `lora_block_0`."""
        from ..ckpt_manager import auto_manager
        all_blocks: Dict[str, LoraBlock] = {}
        for ck_id, item in enumerate(cfg or []):
            sd = _get(item, "state_dict")
            if sd is None:
                path = _get(item, "path")
                sd = auto_manager(path).load_ckpt(path, map_location="cpu")
            part = "lora_ema" if load_ema else "lora"
            sd = sd.get(part, sd)
            per_layer: Dict[str, Dict[str, torch.Tensor]] = {}
            for k, v in sd.items():
                layer, key = k.split(".___." if k.rfind("lora_block.") == -1 else ".lora_block.", 1)
                per_layer.setdefault(layer, {})[key] = v
            only = _get(item, "layers", "all")
            if only != "all":
                keep = get_match_layers(only, self.named_modules)
                per_layer = {k: v for k, v in per_layer.items() if any(k.startswith(p) for p in keep)}
            for layer_name, state in per_layer.items():
                cls, rank, old_format = get_lora_rank_and_cls(state)
                state = {k: v for k, v in state.items() if k != "alpha"}
                if old_format:
                    state = {"layer.W_down": state["layer.lora_down.weight"], "layer.W_up": state["layer.lora_up.weight"]}
                parent_name, host_name = split_module_name(layer_name)
                modules = dict(self.host.named_modules())
                blk = cls.wrap_layer(lora_id_offset + ck_id, modules[layer_name], rank=rank, dropout=_get(item, "dropout", 0.0),
                                     alpha=_get(item, "alpha", 1.0), bias="layer.bias" in state,
                                     alpha_auto_scale=_get(item, "alpha_auto_scale", True), parent_block=modules[parent_name],
                                     host_name=host_name)
                dev = blk.layer.W_down.device
                blk.layer.W_down.copy_(state["layer.W_down"].to(dev, torch.float32))
                blk.layer.W_up.copy_(state["layer.W_up"].to(dev, torch.float32))
                all_blocks[f"{layer_name}.{blk.name}"] = blk
        return LoraGroup(all_blocks)


@torch.no_grad()
def load_lora_state(group: LoraGroup, state: Dict[str, torch.Tensor], strict: bool = True) -> int:
    """Copy a `{'<layer>.___.<key>': tensor}` checkpoint part INTO the existing blocks of `group` (training resume).  The reference
    resumes with `model.load_state_dict(sd['lora'], strict=False)` (train_ac.py:281-288, ckpt_pkl.py:81-88), which only matches
    checkpoints written with `plugin_from_raw` keys; here both key styles land in the blocks that are being trained."""
    n = 0
    for k, v in state.items():
        if ".___." not in k:
            continue
        layer, key = k.split(".___.", 1)
        if layer not in group.plugin_dict:
            if strict:
                raise KeyError(f"checkpoint layer {layer!r} has no LoRA block in the model being resumed")
            continue
        blk = group[layer]
        target = blk
        for part in key.split("."):
            target = getattr(target, part)
        target.copy_(v.to(target.device, target.dtype).reshape(target.shape))
        n += 1
    return n
