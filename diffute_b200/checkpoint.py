"""diffusers-format checkpoint folders: <root>/<subfolder>/config.json + diffusion_pytorch_model.{safetensors,bin}.

This is the on-disk format the reference writes with `save_pretrained` (train_diffute_v1.py:664-669) and reads with
`from_pretrained(path, subfolder=...)` (app.ipynb:545-553).  Only tensors + a JSON config: no diffusers import.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import torch

WEIGHT_NAMES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin",
                "diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.fp16.bin")
_LEGACY_VAE = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def remap_legacy_vae_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """diffusers <= 0.16 named the VAE attention projections query/key/value/proj_attn."""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts and len(parts) >= 2 and parts[-2] in _LEGACY_VAE:
            parts[-2:-1] = _LEGACY_VAE[parts[-2]].split(".")
        out[".".join(parts)] = v
    return out


def load_diffusers_folder(path: str, subfolder: Optional[str] = None) -> Tuple[dict, Dict[str, torch.Tensor]]:
    root = os.path.join(path, subfolder) if subfolder else path
    cfg_path = os.path.join(root, "config.json")
    if not os.path.isfile(cfg_path):
        raise FileNotFoundError(f"{cfg_path} not found")
    with open(cfg_path) as f:
        cfg = json.load(f)
    for name in WEIGHT_NAMES:
        p = os.path.join(root, name)
        if os.path.isfile(p):
            if name.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd = load_file(p)
            else:
                sd = torch.load(p, map_location="cpu", weights_only=True)
            return cfg, {k: v.float() for k, v in sd.items()}
    raise FileNotFoundError(f"no weight file ({', '.join(WEIGHT_NAMES[:2])}) under {root}")


def save_diffusers_folder(path: str, subfolder: Optional[str], cfg: dict, sd: Dict[str, torch.Tensor],
                          class_name: str, safe_serialization: bool = True) -> None:
    root = os.path.join(path, subfolder) if subfolder else path
    os.makedirs(root, exist_ok=True)
    c = {"_class_name": class_name, "_diffusers_version": "0.16.0"}
    c.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()})
    with open(os.path.join(root, "config.json"), "w") as f:
        json.dump(c, f, indent=2)
    sd = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
    if safe_serialization:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(root, "diffusion_pytorch_model.safetensors"))
    else:
        torch.save(sd, os.path.join(root, "diffusion_pytorch_model.bin"))


def load_scheduler_config(path: str, subfolder: Optional[str] = "scheduler") -> dict:
    root = os.path.join(path, subfolder) if subfolder else path
    with open(os.path.join(root, "scheduler_config.json")) as f:
        return json.load(f)
