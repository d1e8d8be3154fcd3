"""Synthetic weights / conditioning / video for benchmarks and smoke tests (no checkpoints or datasets exist
offline, SURVEY.md 'Key facts').  Weights are drawn directly on the GPU: every matrix U(-b, b) with
b = 1/sqrt(fan_in), norm scales 1 + 0.1 N(0,1), biases 0.1 N(0,1) -- in particular the tensors the reference
zero-initialises get non-trivial values, otherwise the network would output exactly zero."""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from .config import instantiate_from_config, load_yaml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_CONFIG = os.path.join(REPO, "configs", "inference_geo4d.yaml")


@torch.no_grad()
def randomize_(module: torch.nn.Module, seed: int = 0, device: Optional[torch.device] = None):
    dev = device or next(module.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() == 1:
            p.copy_(torch.randn(p.shape, device=dev, generator=g) * 0.1)
            if name.endswith(".weight"):
                p.add_(1.0)
        else:
            fan_in = 1
            for d in p.shape[1:]:
                fan_in *= d
            p.copy_((torch.rand(p.shape, device=dev, generator=g) * 2 - 1) / math.sqrt(fan_in))
    return module


def build_model(config_path: str = DEFAULT_CONFIG, device="cuda", seed: int = 0, t: int = 16):
    """LatentVisualDiffusion + point-map VAE from the YAML with seeded synthetic weights and a seeded constant
    conditioning tensor ([1, 77, ctx] text + [1, 16 t, ctx] image tokens)."""
    cfg = load_yaml(config_path)
    with torch.device(device):
        model = instantiate_from_config(cfg["model"])
        pm_vae = instantiate_from_config(cfg["pointmap_vae_config"])
    randomize_(model.model, seed)
    randomize_(model.first_stage_model, seed + 1)
    randomize_(pm_vae, seed + 2)
    model.prepare()
    pm_vae.prepare()
    ctx_dim = cfg["model"]["params"]["unet_config"]["params"]["context_dim"]
    g = torch.Generator(device=device).manual_seed(seed + 3)
    model.set_cached_conditioning(torch.randn(1, 77, ctx_dim, device=device, generator=g),
                                  torch.randn(1, 16 * t, ctx_dim, device=device, generator=g))
    return model, pm_vae, cfg


def synthetic_video(T: int, H: int, W: int, device="cuda", seed: int = 123) -> torch.Tensor:
    """Smooth random field in [-1, 1], [1, 3, T, H, W]."""
    g = torch.Generator(device=device).manual_seed(seed)
    low = torch.randn(1, 3, max(T // 4, 2), H // 16, W // 16, device=device, generator=g)
    vid = torch.nn.functional.interpolate(low, size=(T, H, W), mode="trilinear", align_corners=False)
    return torch.tanh(vid)
