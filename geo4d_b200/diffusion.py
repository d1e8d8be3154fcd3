"""Inference subset of the reference's `LatentVisualDiffusion` (lvdm/models/ddpm3d.py).

Keeps the operator surface the evaluation scripts touch:
schedule buffers with the reference's names, `apply_model`, `encode_first_stage`,
`decode_first_stage`, `decode_first_stage_confhead`, `predict_start_from_z_and_v`,
`predict_eps_from_z_and_v`, `get_learned_conditioning`, the attributes read by
`image_guided_synthesis` (modality, cross_attention, uncond_type, perframe_ae,
scale_factor, encoder_type, model.conditioning_key, ...), and the checkpoint
layout `model.diffusion_model.*` / `first_stage_model.*` + schedule buffers.

Out of scope ("next" row N3): the frozen OpenCLIP text/image
towers and the Resampler.  With the shipped settings their output is a constant
[1, 77+256, 1024] tensor; supply it with `set_cached_conditioning()` (or pass
`c_crossattn` explicitly).  Their checkpoint tensors are accepted and ignored.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import schedule as sched
from .config import instantiate_from_config
from .vae import AutoencoderKL, DiagonalGaussianDistribution

_IGNORED_PREFIXES = ("cond_stage_model.", "embedder.", "model_ema.")   # OpenCLIP towers: outside this port (N3)


class DiffusionWrapper(nn.Module):
    """ddpm3d.py:2523-2597 ('hybrid' / 'concat' / 'crossattn' branches)."""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat: Optional[list] = None, c_crossattn: Optional[list] = None, **kwargs):
        if self.conditioning_key is None:
            return self.diffusion_model(x, t)
        if self.conditioning_key == "concat":
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t, **kwargs)
        if self.conditioning_key == "crossattn":
            return self.diffusion_model(x, t, context=torch.cat(c_crossattn, 1), **kwargs)
        if self.conditioning_key == "hybrid":
            xc = torch.cat([x] + c_concat, dim=1)
            cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
            return self.diffusion_model(xc, t, context=cc, **kwargs)
        raise NotImplementedError(self.conditioning_key)


class LatentVisualDiffusion(nn.Module):
    def __init__(self, unet_config, first_stage_config, cond_stage_config=None, img_cond_stage_config=None,
                 image_proj_stage_config=None, timesteps=1000, beta_schedule="linear", linear_start=1e-4,
                 linear_end=2e-2, parameterization="eps", rescale_betas_zero_snr=False, conditioning_key=None,
                 channels=3, image_size=256, scale_factor=1.0, scale_by_std=False, use_ema=False,
                 uncond_type="empty_seq", use_dynamic_rescale=False, base_scale=0.7, turning_step=400,
                 perframe_ae=False, encoder_type="2d", modality="pc", cross_attention=False,
                 fps_condition_type="fs", first_stage_key="image", cond_stage_key="caption",
                 cond_stage_trainable=False, num_timesteps_cond=1, v_posterior=0.0, **ignored):
        super().__init__()
        assert parameterization in ("eps", "x0", "v")
        self.parameterization = parameterization
        self.channels = channels
        self.image_size = image_size
        self.use_ema = False
        self.uncond_type = uncond_type
        self.use_dynamic_rescale = use_dynamic_rescale
        self.perframe_ae = perframe_ae
        self.encoder_type = encoder_type
        self.modality = modality
        self.cross_attention = cross_attention
        self.fps_condition_type = fps_condition_type
        self.first_stage_key = first_stage_key
        self.cond_stage_key = cond_stage_key
        self.rescale_betas_zero_snr = rescale_betas_zero_snr
        self.scale_factor = scale_factor
        self.perchannel_vae = False
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.first_stage_model = instantiate_from_config(first_stage_config).eval()
        bufs = sched.register_schedule_buffers(timesteps, linear_start, linear_end, beta_schedule,
                                               rescale_betas_zero_snr, v_posterior)
        self.num_timesteps = int(timesteps)
        for k, v in bufs.items():
            self.register_buffer(k, torch.tensor(v))
        zeros = torch.zeros(self.num_timesteps)
        ac = bufs["alphas_cumprod"].astype(np.float64)
        if parameterization != "v":
            with np.errstate(divide="ignore"):
                self.register_buffer("sqrt_recip_alphas_cumprod", torch.tensor(np.sqrt(1.0 / ac), dtype=torch.float32))
                self.register_buffer("sqrt_recipm1_alphas_cumprod",
                                     torch.tensor(np.sqrt(1.0 / ac - 1), dtype=torch.float32))
        else:
            self.register_buffer("sqrt_recip_alphas_cumprod", zeros.clone())
            self.register_buffer("sqrt_recipm1_alphas_cumprod", zeros.clone())
        if use_dynamic_rescale:
            self.register_buffer("scale_arr", torch.tensor(sched.make_scale_arr(self.num_timesteps, base_scale,
                                                                                  turning_step)))
        self.cond_stage_model = None
        self.embedder = None
        # image-token Resampler (ddpm3d.py:2416-2421 `image_proj_model`): native (geo4d_b200/resampler.py); the OpenCLIP
        # towers that feed it are not part of this port -- their (constant) outputs are handed over, see below
        self.image_proj_model = None
        if image_proj_stage_config is not None and "target" in image_proj_stage_config:
            try:
                self.image_proj_model = instantiate_from_config(image_proj_stage_config)
            except (ImportError, AttributeError, NotImplementedError):
                self.image_proj_model = None
        self._cached_cond: Optional[Dict[str, torch.Tensor]] = None

    # ------------------------------------------------------------------ checkpoint
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the reference model.ckpt['state_dict'] as is: conditioning-tower tensors (out of scope)
        and non-persistent extras are dropped, everything else must match exactly when strict."""
        sd = {k: v for k, v in state_dict.items()
              if not k.startswith(_IGNORED_PREFIXES) and k not in ("logvar", "lvlb_weights")}
        if self.image_proj_model is None:
            sd = {k: v for k, v in sd.items() if not k.startswith("image_proj_model.")}
        elif not any(k.startswith("image_proj_model.") for k in sd):
            # a state dict without the Resampler (e.g. synthetic weights): keep the module's own tensors
            sd.update({"image_proj_model." + k: v for k, v in self.image_proj_model.state_dict().items()})
        return super().load_state_dict(sd, strict=strict, **kw)

    @property
    def device(self):
        return self.betas.device

    def prepare(self):
        self.model.diffusion_model.prepare()
        self.first_stage_model.prepare()
        return self

    # ------------------------------------------------------------------ conditioning
    def set_cached_conditioning(self, text_emb: torch.Tensor, img_emb: Optional[torch.Tensor] = None,
                                clip_image_tokens: Optional[torch.Tensor] = None):
        """text_emb [1, 77, 1024] (FrozenOpenCLIPEmbedder output for the fixed prompt) and either img_emb
        [1, 16*t, 1024] (Resampler output for the all-zero image, infer_geo4d.py:150-156) or clip_image_tokens
        [1, 257, 1280] (FrozenOpenCLIPImageEmbedderV2 output), which the native Resampler projects."""
        if img_emb is None and clip_image_tokens is not None:
            if self.image_proj_model is None:
                raise NotImplementedError("no image_proj_model was configured (image_proj_stage_config)")
            img_emb = self.image_proj_model(clip_image_tokens.to(self.device))
        self._cached_cond = {"text": text_emb, "img": img_emb}

    def get_learned_conditioning(self, c):
        if self._cached_cond is None:
            raise NotImplementedError(
                "the OpenCLIP conditioning towers are outside this port (SURVEY.md N3); call "
                "set_cached_conditioning() with the constant conditioning tensors")
        b = len(c) if isinstance(c, (list, tuple)) else 1
        return self._cached_cond["text"].to(self.device).expand(b, -1, -1)

    def get_image_conditioning(self, b: int):
        if self._cached_cond is None or self._cached_cond["img"] is None:
            raise NotImplementedError("no cached image conditioning; see set_cached_conditioning()")
        return self._cached_cond["img"].to(self.device).expand(b, -1, -1)

    # ------------------------------------------------------------------ v-parameterisation helpers
    @staticmethod
    def _extract(a, t, x_shape):
        out = a.gather(-1, t)
        return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

    def predict_start_from_z_and_v(self, x_t, t, v):  # ddpm3d.py:278-284
        return (self._extract(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t -
                self._extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def predict_eps_from_z_and_v(self, x_t, t, v):  # ddpm3d.py:286-290
        return (self._extract(self.sqrt_alphas_cumprod, t, x_t.shape) * v +
                self._extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * x_t)

    # ------------------------------------------------------------------ U-Net
    def apply_model(self, x_noisy, t, cond, **kwargs):  # ddpm3d.py:1002-1017
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = "c_concat" if self.model.conditioning_key == "concat" else "c_crossattn"
            cond = {key: cond}
        out = self.model(x_noisy, t, **cond, **kwargs)
        return out[0] if isinstance(out, tuple) else out

    # ------------------------------------------------------------------ first stage
    def _frames(self, x):
        if self.encoder_type == "2d" and x.dim() == 5:
            b, c, t, h, w = x.shape
            return x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), (b, t)
        return x, None

    @staticmethod
    def _unframes(y, bt):
        if bt is None:
            return y
        b, t = bt
        return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)

    def get_first_stage_encoding(self, encoder_posterior, noise=None):  # ddpm3d.py:674-681
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            z = encoder_posterior.sample(noise=noise)
        elif isinstance(encoder_posterior, torch.Tensor):
            z = encoder_posterior
        else:
            raise NotImplementedError(type(encoder_posterior))
        return self.scale_factor * z

    @torch.no_grad()
    def encode_first_stage(self, x, noise=None):
        """ddpm3d.py:683-707.  All frames are encoded in one batch (perframe_ae only trades memory for speed in
        the reference; per-frame GroupNorm statistics make the results identical)."""
        xf, bt = self._frames(x)
        z = self.get_first_stage_encoding(self.first_stage_model.encode(xf), noise=noise)
        return self._unframes(z, bt)

    @torch.no_grad()
    def decode_first_stage(self, z, **kwargs):  # ddpm3d.py:802-823,935-936
        zf, bt = self._frames(z)
        return self._unframes(self.first_stage_model.decode(zf * (1.0 / self.scale_factor)), bt)

    @torch.no_grad()
    def decode_first_stage_confhead(self, z, **kwargs):  # ddpm3d.py:849-870,926-928
        zf, bt = self._frames(z)
        return self._unframes(self.first_stage_model.decode_with_conf_adaptor(zf * (1.0 / self.scale_factor)), bt)
