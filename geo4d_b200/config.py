"""Config seam: `instantiate_from_config({'target': 'pkg.mod.Class', 'params': {...}})` as in the reference's
utils/utils.py:27-42, with the reference's dotted class paths (configs/inference_geo4d.yaml) resolved to the
B200-native classes of this package so the shipped YAML works unchanged."""
from __future__ import annotations

import importlib

TARGET_ALIASES = {
    "lvdm.models.ddpm3d.LatentVisualDiffusion": "geo4d_b200.diffusion.LatentVisualDiffusion",
    "lvdm.models.ddpm3d.DiffusionWrapper": "geo4d_b200.diffusion.DiffusionWrapper",
    "lvdm.modules.networks.openaimodel3d.UNetModel": "geo4d_b200.unet.UNetModel",
    "lvdm.models.autoencoder.AutoencoderKL": "geo4d_b200.vae.AutoencoderKL",
    "lvdm.models.samplers.ddim.DDIMSampler": "geo4d_b200.sampler.DDIMSampler",
    "lvdm.models.samplers.ddim_multiplecond.DDIMSampler": "geo4d_b200.sampler.DDIMSampler_multicond",
    "lvdm.modules.encoders.resampler.Resampler": "geo4d_b200.resampler.Resampler",
    "dust3r.cloud_opt.optimizer_group.LightPointCloudGroupOptimizer":
        "geo4d_b200.cloud_opt.LightPointCloudGroupOptimizer",
}


def get_obj_from_str(string: str):
    string = TARGET_ALIASES.get(string, string)
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**dict(config.get("params", dict()) or {}))


def load_yaml(path: str) -> dict:
    """The shipped YAML has CRLF line endings; omegaconf is not required."""
    import yaml
    with open(path, "r") as f:
        return yaml.safe_load(f.read().replace("\r\n", "\n"))
