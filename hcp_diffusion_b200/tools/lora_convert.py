"""LoRA checkpoint interchange with the webui / kohya naming (SD1.x UNet and CLIP text encoder).

Behavioural restatement of the reference converter hcpdiff/tools/lora_convert.py (LoraConverter :9-209, CLI :192-230); pinned to
it key-by-key by tests/golden/lora_webui_keys.json (generated from the real class by tests/golden/make_golden.py).

  hcpdiff key                                              webui key
  <module.path>.___.layer.W_down   [r, in(,kh,kw)]  <->    lora_unet_<module_path>.lora_down.weight
  <module.path>.___.layer.W_up     [out, r(,1,1)]   <->    lora_unet_<module_path>.lora_up.weight
  <module.path>.___.alpha          []               <->    lora_unet_<module_path>.alpha

The webui name flattens '.' to '_'; going back, every '_' becomes '.' except inside the module names that contain an underscore
themselves (`down_blocks`, `to_q`, ...).  `auto_scale_alpha` multiplies W_down by sqrt(rank) and W_up by sqrt(rank) in either
direction, exactly like the reference (:170-189) -- hcpdiff stores alpha/rank in the `alpha` buffer, webui divides by rank at load.
"""
from __future__ import annotations

import argparse
import math
import os
from typing import Dict, Iterable, Tuple

import torch

# module names that contain '_' and must survive the '_' -> '.' expansion (reference :10-11)
UNET_NAMES = ("down_blocks", "up_blocks", "mid_block", "transformer_blocks", "to_q", "to_k", "to_v", "to_out", "proj_in", "proj_out",
              "input_blocks", "middle_block", "output_blocks")
TE_NAMES = ("self_attn", "q_proj", "v_proj", "k_proj", "out_proj", "text_model")
PREFIX_UNET = "lora_unet_"
PREFIX_TE = "lora_te_"
_W = {"lora_down.weight": "W_down", "lora_up.weight": "W_up"}


def _expand(flat: str, protected: Iterable[str]) -> str:
    """'down_blocks_0_attentions_0_..._to_q' -> 'down_blocks.0.attentions.0....to_q'."""
    for name in protected:                       # same left-to-right substitution order as the reference
        flat = flat.replace(name, name.replace("_", "%"))
    return flat.replace("_", ".").replace("%", "_")


class LoraConverter:
    def convert_to_webui(self, sd_unet: Dict[str, torch.Tensor], sd_te: Dict[str, torch.Tensor] = None, auto_scale_alpha: bool = False,
                         sdxl: bool = False) -> Dict[str, torch.Tensor]:
        if sdxl:
            raise NotImplementedError("SDXL key maps are outside the SD1.x hot path")
        out = self._to_webui(sd_unet, PREFIX_UNET)
        out.update(self._to_webui(sd_te or {}, PREFIX_TE))
        if auto_scale_alpha:
            out = {k: self._scale(k, v, "lora_up", "lora_down") for k, v in out.items()}
        return out

    def convert_from_webui(self, state: Dict[str, torch.Tensor], auto_scale_alpha: bool = False, sdxl: bool = False
                           ) -> Tuple[Dict[str, Dict[str, torch.Tensor]], Dict[str, Dict[str, torch.Tensor]]]:
        """-> ({'lora': text-encoder part}, {'lora': unet part}), the order the reference returns them in (:36)."""
        if sdxl:
            raise NotImplementedError("SDXL key maps are outside the SD1.x hot path")
        sd_unet = self._from_webui(state, PREFIX_UNET, UNET_NAMES)
        sd_te = self._from_webui(state, PREFIX_TE, TE_NAMES)
        if auto_scale_alpha:
            sd_unet = {k: self._scale(k, v, "W_up", "W_down") for k, v in sd_unet.items()}
            sd_te = {k: self._scale(k, v, "W_up", "W_down") for k, v in sd_te.items()}
        return {"lora": sd_te}, {"lora": sd_unet}

    @staticmethod
    def _to_webui(state, prefix):
        out = {}
        for k, v in state.items():
            path, key = k.split(".___.", 1)
            if key.endswith("W_down"):
                key = "lora_down.weight"
            elif key.endswith("W_up"):
                key = "lora_up.weight"
            out[f"{prefix}{path.replace('.', '_')}.{key}"] = v
        return out

    @staticmethod
    def _from_webui(state, prefix, protected):
        out = {}
        for k, v in state.items():
            if not k.startswith(prefix):
                continue
            flat, key = k[len(prefix):].split(".", 1)
            path = _expand(flat, protected)
            out[f"{path}.___.{key}" if key == "alpha" else f"{path}.___.layer.{_W[key]}"] = v
        return out

    @staticmethod
    def _scale(key, v, up, down):
        if up in key:
            return v * math.sqrt(v.shape[1])
        if down in key:
            return v * math.sqrt(v.shape[0])
        return v


def main(argv=None):
    from ..ckpt_manager import auto_manager
    ap = argparse.ArgumentParser(description="hcpdiff <-> webui LoRA checkpoint conversion (same flags as the reference tool)")
    ap.add_argument("--lora_path", required=True)
    ap.add_argument("--lora_path_TE", default=None)
    ap.add_argument("--dump_path", required=True)
    ap.add_argument("--from_webui", action="store_true")
    ap.add_argument("--to_webui", action="store_true")
    ap.add_argument("--auto_scale_alpha", action="store_true")
    ap.add_argument("--sdxl", action="store_true")
    args = ap.parse_args(argv)
    conv = LoraConverter()
    mgr = auto_manager(args.lora_path)
    name = os.path.basename(args.lora_path)
    if args.from_webui:
        sd_te, sd_unet = conv.convert_from_webui(mgr.load_ckpt(args.lora_path), args.auto_scale_alpha, args.sdxl)
        os.makedirs(args.dump_path, exist_ok=True)
        mgr._save_ckpt(sd_te, save_path=os.path.join(args.dump_path, "TE-" + name))
        mgr._save_ckpt(sd_unet, save_path=os.path.join(args.dump_path, "unet-" + name))
    elif args.to_webui:
        sd_unet = mgr.load_ckpt(args.lora_path)
        sd_te = mgr.load_ckpt(args.lora_path_TE) if args.lora_path_TE else {"lora": {}}
        mgr._save_ckpt(conv.convert_to_webui(sd_unet["lora"], sd_te["lora"], args.auto_scale_alpha, args.sdxl), save_path=args.dump_path)
    else:
        ap.error("one of --from_webui / --to_webui is required")


if __name__ == "__main__":
    main()
