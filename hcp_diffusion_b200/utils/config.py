"""Minimal loader for hcpdiff-style yaml configs (omegaconf / hydra are not available offline).

Implements the subset the training entrypoint needs, with the reference's semantics:
  * `_base_: [file, ...]` recursive inheritance: the reference folds `cfg = merge(load(base), cfg)` over the list
    (hcpdiff/utils/utils.py:56-64), so the file itself overrides every base and an EARLIER base overrides a LATER one;
  * the `'---'` sentinel travels through the merges like any value and the keys still holding it are deleted once, after
    the command-line overrides (utils.py:43-54, 66-72);
  * `key.sub=value` dot-list overrides from the command line (utils.py:66-72);
  * `${hcp.eval:...}`, `${hcp.time:}`, `${hcp.dtype:...}` resolvers (hcpdiff/utils/cfg_resolvers.py:11-16) and plain
    `${a.b}` interpolation;
  * `_target_` / `_partial_` instantiation of objects (hydra.utils.instantiate as used at hcpdiff/train_ac.py:55).
"""
from __future__ import annotations

import functools
import importlib
import os
import re
import time
from typing import Any, Dict, List

import yaml


class Cfg(dict):
    """dict with attribute access (`cfg.train.lr`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def get_path(self, path: str, default=None):
        node = self
        for part in path.split("."):
            if isinstance(node, list):
                part = int(part)
                if part >= len(node):
                    return default
                node = node[part]
            elif isinstance(node, dict) and part in node:
                node = node[part]
            else:
                return default
        return node


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(base, new):
    """OmegaConf.merge(base, new) for plain containers: dicts merge recursively, everything else (lists, scalars, the '---'
    sentinel) is replaced by `new`."""
    if isinstance(base, dict) and isinstance(new, dict):
        out = dict(base)
        for k, v in new.items():
            out[k] = _merge(out[k], v) if k in out else v
        return out
    return new


def _remove_undefined(node):
    """Delete every key / list item whose value is the '---' sentinel (reference remove_config_undefined, utils.py:43-54)."""
    if isinstance(node, dict):
        return {k: _remove_undefined(v) for k, v in node.items() if not (isinstance(v, str) and v == "---")}
    if isinstance(node, list):
        return [_remove_undefined(v) for v in node if not (isinstance(v, str) and v == "---")]
    return node


def _load_with_base(path: str) -> Dict[str, Any]:
    """reference load_config(path, remove_undefined=False), utils.py:56-64."""
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    for b in cfg.get("_base_", None) or []:
        bp = b if os.path.isabs(b) or os.path.exists(b) else os.path.join(os.path.dirname(path), b)
        cfg = _merge(_load_with_base(bp), cfg)          # what is already in cfg (the file, earlier bases) wins
    cfg.pop("_base_", None)
    return cfg


def _parse_scalar(s: str):
    if s.strip() == "---":          # the undefined sentinel, not a yaml document marker
        return "---"
    try:
        return yaml.safe_load(s)
    except yaml.YAMLError:
        return s


def _set_path(cfg: Dict[str, Any], path: str, value):
    node = cfg
    parts = path.split(".")
    for i, p in enumerate(parts[:-1]):
        if isinstance(node, list):
            node = node[int(p)]
        else:
            node = node.setdefault(p, {})
    last = parts[-1]
    if isinstance(node, list):
        node[int(last)] = value
    else:
        node[last] = value


_DTYPES = {"fp32": "torch.float32", "amp": "torch.float32", "fp16": "torch.float16", "bf16": "torch.bfloat16"}
_RX = re.compile(r"\$\{([^${}]+)\}")


def _resolve(value, root: Cfg):
    if isinstance(value, dict):
        return Cfg({k: _resolve(v, root) for k, v in value.items()})
    if isinstance(value, list):
        return [_resolve(v, root) for v in value]
    if not isinstance(value, str) or "${" not in value:
        return value

    def one(expr: str):
        if expr.startswith("hcp.eval:"):
            code = expr[len("hcp.eval:"):].strip().strip("\"'")
            return eval(code, {"__builtins__": {}}, {})
        if expr.startswith("hcp.time:"):
            fmt = expr[len("hcp.time:"):] or "%Y-%m-%d-%H-%M-%S"
            return time.strftime(fmt)
        if expr.startswith("hcp.dtype:"):
            return _DTYPES.get(expr[len("hcp.dtype:"):], "torch.float32")
        v = root.get_path(expr)
        return _resolve(v, root)

    cur = value
    for _ in range(8):                       # nested interpolations
        m = _RX.fullmatch(cur.strip())
        if m:
            cur = one(m.group(1))
            if not isinstance(cur, str) or "${" not in cur:
                return cur
            continue
        new = _RX.sub(lambda mm: str(one(mm.group(1))), cur)
        if new == cur:
            break
        cur = new
    return cur


def load_config_with_cli(path: str, args_list: List[str] = None) -> Cfg:
    cfg = _load_with_base(path)
    for item in args_list or []:
        key, _, val = item.partition("=")
        _set_path(cfg, key, _parse_scalar(val))
    root = _wrap(_remove_undefined(cfg))
    return _resolve(root, root)


def _locate(name: str):
    mod, _, attr = name.rpartition(".")
    obj = importlib.import_module(mod)
    return getattr(obj, attr)


def instantiate(node):
    """Build objects for `_target_` nodes (depth first); `_partial_: true` yields functools.partial."""
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return Cfg({k: instantiate(v) for k, v in node.items()})
    target = _locate(node["_target_"])
    partial = bool(node.get("_partial_", False))
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_", "_args_")}
    args = [instantiate(v) for v in node.get("_args_", [])]
    if partial:
        return functools.partial(target, *args, **kwargs)
    return target(*args, **kwargs)
