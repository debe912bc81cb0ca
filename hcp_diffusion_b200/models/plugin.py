"""Plugin surface of the hot path: the patch-style plugin container that REPLACES a host layer in its parent.

API-compatible restatement of the part of the reference plugin framework the LoRA hot path uses
(hcpdiff/models/plugin.py: BasePluginBlock :20, WrapablePlugin :58-105, PatchPluginContainer :223-262,
PatchPluginBlock :264-315, PluginGroup :317-348).  The hook-based plugin kinds (SinglePluginBlock, PluginBlock,
MultiPluginBlock -- ControlNet etc.) are outside the hot path and are not provided.

Contract kept from the reference (checked by tests/test_plugin_surface.py against golden data generated from the real
reference classes):
  * the container stores the host as `_host`, so base weights appear as `<layer>._host.weight` in `state_dict()`;
  * plugins are attributes of the container named by `plugin.name` (`lora_block_<id>`), listed in `plugin_names`;
  * `PluginGroup.state_dict()` keys are `<layer path>.___.<plugin state key>`.
"""
from __future__ import annotations

import re
import weakref
from typing import Dict, Iterable, Optional

import torch
from torch import nn


def split_module_name(layer_name: str):
    """'a.b.c' -> ('a.b', 'c'); 'c' -> ('', 'c')   (reference hcpdiff/utils/net_utils.py:219-225)."""
    parent, _, name = layer_name.rpartition(".")
    return parent, name


class BasePluginBlock(nn.Module):
    def __init__(self, name: str):
        super().__init__()
        self.name = name

    def remove(self):
        pass

    def set_hyper_params(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @staticmethod
    def extract_state_without_plugin(model: nn.Module, trainable: bool = False) -> Dict[str, torch.Tensor]:
        """state_dict of `model` minus every plugin subtree (optionally trainable tensors only)."""
        plugin_prefixes = [k for k, v in model.named_modules() if isinstance(v, BasePluginBlock)]
        keep = {k for k, p in model.named_parameters() if p.requires_grad} if trainable else None
        out = {}
        for k, v in model.state_dict().items():
            if keep is not None and k not in keep:
                continue
            if any(k.startswith(p) for p in plugin_prefixes):
                continue
            out[k] = v
        return out

    def get_trainable_parameters(self) -> Iterable[nn.Parameter]:
        return self.parameters()


class WrapablePlugin:
    wrapable_classes = ()

    @classmethod
    def wrap_layer(cls, name, layer: nn.Module, **kwargs):
        return cls(name, layer, **kwargs)

    @classmethod
    def named_modules_with_exclude(cls, root: nn.Module, prefix: str = "", exclude_key: Optional[str] = None,
                                   exclude_classes=tuple(), _memo=None):
        """named_modules() that does not descend into names matching `exclude_key` / instances of `exclude_classes`."""
        if _memo is None:
            _memo = set()
        if root in _memo:
            return
        _memo.add(root)
        if exclude_key is not None and re.search(exclude_key, prefix):
            return
        if isinstance(root, exclude_classes):
            return
        yield prefix, root
        for name, child in root._modules.items():
            if child is None:
                continue
            sub = f"{prefix}.{name}" if prefix else name
            yield from cls.named_modules_with_exclude(child, sub, exclude_key, exclude_classes, _memo)


class PatchPluginContainer(nn.Module):
    """Takes the place of `host` inside `parent_block` (attribute `host_name`) and owns the plugins patched onto it."""

    def __init__(self, host_name: str, host: nn.Module, parent_block: nn.Module):
        super().__init__()
        self._host = host
        self.host_name = host_name
        self.parent_block = weakref.ref(parent_block)
        self.plugin_names = []
        delattr(parent_block, host_name)
        setattr(parent_block, host_name, self)

    def add_plugin(self, name: str, plugin: "PatchPluginBlock"):
        if name in self.plugin_names:
            # the reference silently overwrites the attribute and lists the name twice (plugin.py:236-238): the first block is
            # orphaned but its delta is applied twice.  Loading a checkpoint INTO existing blocks is `load_lora_state` / resume.
            raise ValueError(f"plugin {name!r} is already patched onto this layer: use another lora_id (lora_id_offset) or load the "
                             "checkpoint into the existing blocks")
        setattr(self, name, plugin)
        self.plugin_names.append(name)

    def remove_plugin(self, name: str):
        delattr(self, name)
        self.plugin_names.remove(name)
        if not self.plugin_names:
            self.remove()

    def remove(self):
        parent = self.parent_block()
        delattr(parent, self.host_name)
        setattr(parent, self.host_name, self._host)

    def __iter__(self):
        for name in self.plugin_names:
            yield name, self[name]

    def __getitem__(self, name):
        return getattr(self, name)


class PatchPluginBlock(BasePluginBlock, WrapablePlugin):
    container_cls = PatchPluginContainer

    def __init__(self, name: str, host: nn.Module, host_model=None, parent_block: nn.Module = None, host_name: str = None):
        super().__init__(name)
        real_host = host._host if isinstance(host, self.container_cls) else host
        self.host = weakref.ref(real_host)
        self.parent_block = weakref.ref(parent_block)
        self.host_name = host_name
        container = host if isinstance(host, self.container_cls) else self.container_cls(host_name, host, parent_block)
        container.add_plugin(name, self)
        self.container = weakref.ref(container)

    def remove(self):
        self.container().remove_plugin(self.name)

    @classmethod
    def wrap_model(cls, name, host: nn.Module, exclude_key=None, exclude_classes=tuple(), **kwargs):
        """Patch every wrapable layer below `host` (or `host` itself).  Returns {relative layer name: plugin}.
        Mirrors reference plugin.py:297-315: `_host` subtrees are skipped, existing containers are re-used."""
        out = {}
        if isinstance(host, cls.wrapable_classes):
            out[""] = cls.wrap_layer(name, host, **kwargs)
            return out
        named = dict(cls.named_modules_with_exclude(host, exclude_key=exclude_key or "_host", exclude_classes=exclude_classes))
        for layer_name, layer in named.items():
            if isinstance(layer, cls.wrapable_classes) or isinstance(layer, cls.container_cls):
                if "parent_block" in kwargs:
                    parent_name, host_name = split_module_name(layer_name)
                    kwargs["parent_block"] = named[parent_name]
                    kwargs["host_name"] = host_name
                out[layer_name] = cls.wrap_layer(name, layer, **kwargs)
        return out


class PluginGroup:
    """{host layer path: plugin}; the unit checkpoints are saved in (reference plugin.py:317-348)."""

    def __init__(self, plugin_dict: Dict[str, BasePluginBlock]):
        self.plugin_dict = plugin_dict

    def __setitem__(self, k, v):
        self.plugin_dict[k] = v

    def __getitem__(self, k):
        return self.plugin_dict[k]

    @property
    def plugin_name(self):
        return None if self.empty() else next(iter(self.plugin_dict.values())).name

    def remove(self):
        for plugin in self.plugin_dict.values():
            plugin.remove()

    def state_dict(self, model: nn.Module = None):
        if model is None:
            return {f"{k}.___.{ks}": vs for k, v in self.plugin_dict.items() for ks, vs in v.state_dict().items()}
        sd = model.state_dict()
        return {f"{k}.___.{ks}": sd[f"{k}.{v.name}.{ks}"] for k, v in self.plugin_dict.items() for ks in v.state_dict().keys()}

    def state_keys_raw(self):
        return [f"{k}.{v.name}.{ks}" for k, v in self.plugin_dict.items() for ks in v.state_dict().keys()]

    def empty(self):
        return len(self.plugin_dict) == 0
