"""CPU restatement of the reference noise schedule and DDIM sampler (test oracle).

Follows lvdm/models/utils_diffusion.py (make_beta_schedule:31-53,
make_ddim_timesteps:56-76, make_ddim_sampling_parameters:79-91,
rescale_zero_terminal_snr:112-144, rescale_noise_cfg:147-158),
lvdm/models/ddpm3d.py (register_schedule:162-225, scale_arr:585-590,
predict_start_from_z_and_v:278-284, predict_eps_from_z_and_v:286-290) and
lvdm/models/samplers/ddim.py (make_schedule:24-57, ddim_sampling:134-203,
p_sample_ddim:205-279) of jzr99/Geo4D.

Pinned by the known-answer values of SURVEY.md section 4 (computed with the
reference functions) and by oracle/gen_golden.py (reference DDIMSampler run on
CPU with a stub model).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch


def make_beta_schedule_linear(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """utils_diffusion.py:32-35 ('linear' = linspace of sqrt(beta), squared; fp64)."""
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2


def rescale_zero_terminal_snr(betas):
    """utils_diffusion.py:112-144."""
    alphas = 1.0 - betas
    abar_sqrt = np.sqrt(np.cumprod(alphas, axis=0))
    a0 = abar_sqrt[0].copy()
    aT = abar_sqrt[-1].copy()
    abar_sqrt = abar_sqrt - aT
    abar_sqrt = abar_sqrt * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    alphas = np.concatenate([abar[0:1], abar[1:] / abar[:-1]])
    return 1 - alphas


def make_ddim_timesteps(method: str, num_ddim: int, num_ddpm: int = 1000):
    """utils_diffusion.py:56-76."""
    if method == "uniform":
        c = num_ddpm // num_ddim
        return np.asarray(list(range(0, num_ddpm, c))) + 1
    if method == "uniform_trailing":
        c = num_ddpm / num_ddim
        return np.flip(np.round(np.arange(num_ddpm, 0, -c))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(num_ddpm * .8), num_ddim)) ** 2).astype(int) + 1
    raise NotImplementedError(method)


@dataclass
class Schedule:
    """Everything the sampler reads from the model (ddpm3d.py:162-225, 585-590)."""
    betas: np.ndarray
    alphas_cumprod: np.ndarray           # fp32, as registered
    alphas_cumprod_prev: np.ndarray
    sqrt_alphas_cumprod: np.ndarray
    sqrt_one_minus_alphas_cumprod: np.ndarray
    scale_arr: Optional[np.ndarray]

    @staticmethod
    def geo4d(timesteps=1000, linear_start=0.00085, linear_end=0.012, zero_snr=True,
              use_dynamic_rescale=True, base_scale=0.7, turning_step=400) -> "Schedule":
        betas = make_beta_schedule_linear(timesteps, linear_start, linear_end)
        if zero_snr:
            betas = rescale_zero_terminal_snr(betas)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        f32 = lambda a: np.asarray(a, dtype=np.float32)
        scale = None
        if use_dynamic_rescale:
            scale = f32(np.concatenate((np.linspace(1.0, base_scale, turning_step),
                                        np.full(timesteps, base_scale))))
        return Schedule(f32(betas), f32(ac), f32(ac_prev), f32(np.sqrt(ac)),
                        f32(np.sqrt(1.0 - ac)), scale)


@dataclass
class DDIMTables:
    timesteps: np.ndarray        # ascending DDPM indices, len S
    alphas: np.ndarray           # fp32 (sliced from the fp32 alphas_cumprod)
    alphas_prev: np.ndarray
    sigmas: np.ndarray
    scale: Optional[np.ndarray]
    scale_prev: Optional[np.ndarray]


def make_ddim_tables(sch: Schedule, S: int, spacing="uniform_trailing", eta=0.0) -> DDIMTables:
    """ddim.py:24-57 + utils_diffusion.py:79-91."""
    ts = make_ddim_timesteps(spacing, S, len(sch.alphas_cumprod))
    ac = sch.alphas_cumprod
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist(), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    sc = sc_prev = None
    if sch.scale_arr is not None:
        sc = sch.scale_arr[ts]
        sc_prev = np.concatenate([sc[0:1], sc[:-1]])
    return DDIMTables(ts, alphas.astype(np.float32), alphas_prev, np.asarray(sigmas, np.float32),
                      sc, sc_prev)


def ddim_step_v(x, v, sch: Schedule, tab: DDIMTables, index: int, noise=None):
    """One p_sample_ddim update for the v-parameterisation (ddim.py:231-277).

    Note the mixed tables: e_t / pred_x0 use the full-schedule sqrt tables
    gathered at the DDPM timestep (ddpm3d.py:278-290), x_prev uses the DDIM
    alphas_prev[index]."""
    t = int(tab.timesteps[index])
    sa = float(sch.sqrt_alphas_cumprod[t])
    s1 = float(sch.sqrt_one_minus_alphas_cumprod[t])
    e_t = sa * v + s1 * x
    pred_x0 = sa * x - s1 * v
    if tab.scale is not None:
        pred_x0 = pred_x0 * (float(tab.scale_prev[index]) / float(tab.scale[index]))
    a_prev = torch.tensor(float(tab.alphas_prev[index]), dtype=torch.float32)
    sigma = torch.tensor(float(tab.sigmas[index]), dtype=torch.float32)
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma * noise
    return x_prev, pred_x0


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """utils_diffusion.py:147-158."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


@torch.no_grad()
def ddim_sample(apply_model: Callable, x_T: torch.Tensor, sch: Schedule, S: int,
                spacing="uniform_trailing", eta=0.0, cfg_scale=1.0, apply_model_uncond=None,
                guidance_rescale=0.0, noise_fn=None, apply_model_uncond_img=None, cfg_img=None):
    """ddim.py:134-203 loop with the v-parameterisation; apply_model(x, t_long[b]) -> v.
    With apply_model_uncond_img the 3-way guidance of ddim_multiplecond.py:226-236 is used."""
    tab = make_ddim_tables(sch, S, spacing, eta)
    x = x_T
    b = x.shape[0]
    pred_x0 = x
    for i, step in enumerate(np.flip(tab.timesteps)):
        index = S - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        v = apply_model(x, ts)
        if apply_model_uncond is not None and cfg_scale != 1.0:
            v_u = apply_model_uncond(x, ts)
            v_c = v
            if apply_model_uncond_img is not None:
                v_ui = apply_model_uncond_img(x, ts)
                v = v_u + (cfg_scale if cfg_img is None else cfg_img) * (v_ui - v_u) + cfg_scale * (v_c - v_ui)
            else:
                v = v_u + cfg_scale * (v_c - v_u)
            if guidance_rescale > 0.0:
                v = rescale_noise_cfg(v, v_c, guidance_rescale)
        noise = noise_fn(x.shape) if (noise_fn is not None and eta > 0) else None
        x, pred_x0 = ddim_step_v(x, v, sch, tab, index, noise)
    return x, pred_x0
