"""Noise schedule and DDIM tables (host side, numpy) for the Geo4D sampler.

Same arithmetic as the reference helpers it stands in for:
lvdm/models/utils_diffusion.py make_beta_schedule:31-53 ('linear'),
rescale_zero_terminal_snr:112-144, make_ddim_timesteps:56-76,
make_ddim_sampling_parameters:79-91, and the buffers of
DDPM.register_schedule (lvdm/models/ddpm3d.py:162-225) / scale_arr (:585-590).
Everything is computed in float64 and stored as float32 exactly where the
reference stores float32, so the tables are bit-identical (tests/test_schedule.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def make_beta_schedule(schedule: str, n_timestep: int, linear_start=1e-4, linear_end=2e-2) -> np.ndarray:
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule '{schedule}' is not used by Geo4D")
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2


def rescale_zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    abar_sqrt = np.sqrt(np.cumprod(1.0 - betas, axis=0))
    first, last = abar_sqrt[0].copy(), abar_sqrt[-1].copy()
    abar_sqrt = (abar_sqrt - last) * (first / (first - last))
    abar = abar_sqrt ** 2
    alphas = np.concatenate([abar[0:1], abar[1:] / abar[:-1]])
    return 1.0 - alphas


def make_ddim_timesteps(method: str, num_ddim: int, num_ddpm: int) -> np.ndarray:
    if method == "uniform":
        return np.asarray(list(range(0, num_ddpm, num_ddpm // num_ddim))) + 1
    if method == "uniform_trailing":
        c = num_ddpm / num_ddim
        return np.flip(np.round(np.arange(num_ddpm, 0, -c))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(num_ddpm * .8), num_ddim)) ** 2).astype(int) + 1
    raise NotImplementedError(f'There is no ddim discretization method called "{method}"')


def register_schedule_buffers(timesteps=1000, linear_start=1e-4, linear_end=2e-2, beta_schedule="linear",
                              rescale_betas_zero_snr=False, v_posterior=0.0) -> Dict[str, np.ndarray]:
    """float32 buffers with the reference's names (ddpm3d.py:185-211)."""
    betas = make_beta_schedule(beta_schedule, timesteps, linear_start, linear_end)
    if rescale_betas_zero_snr:
        betas = rescale_zero_terminal_snr(betas)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    with np.errstate(divide="ignore", invalid="ignore"):
        post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
        bufs = {
            "betas": betas,
            "alphas_cumprod": ac,
            "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": np.sqrt(ac),
            "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
            "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
            "posterior_variance": post_var,
            "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
            "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
            "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        }
    return {k: np.asarray(v, dtype=np.float32) for k, v in bufs.items()}


def make_scale_arr(timesteps=1000, base_scale=0.7, turning_step=400) -> np.ndarray:
    return np.concatenate((np.linspace(1.0, base_scale, turning_step),
                           np.full(timesteps, base_scale))).astype(np.float32)


class DDIMTables:
    """What DDIMSampler.make_schedule (ddim.py:24-57) derives for one (S, spacing, eta)."""

    def __init__(self, alphas_cumprod: np.ndarray, scale_arr: Optional[np.ndarray], S: int,
                 spacing: str = "uniform", eta: float = 0.0):
        ac = np.asarray(alphas_cumprod, dtype=np.float32)
        self.timesteps = make_ddim_timesteps(spacing, S, ac.shape[0])
        ts = self.timesteps
        self.alphas = ac[ts]
        self.alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
        with np.errstate(divide="ignore", invalid="ignore"):
            self.sigmas = eta * np.sqrt((1 - self.alphas_prev) / (1 - self.alphas) *
                                        (1 - self.alphas / self.alphas_prev))
        self.sqrt_one_minus_alphas = np.sqrt(1.0 - self.alphas)
        self.scale = self.scale_prev = None
        if scale_arr is not None:
            self.scale = np.asarray(scale_arr, dtype=np.float32)[ts]
            self.scale_prev = np.concatenate([self.scale[0:1], self.scale[:-1]])

    def step_coefficients(self, sqrt_ac: np.ndarray, sqrt_1mac: np.ndarray) -> np.ndarray:
        """[S, 6] float32 rows {sqrt(abar_t), sqrt(1-abar_t), scale_prev/scale_t, sqrt(a_prev),
        sqrt(1-a_prev-sigma^2), sigma} in SAMPLING order (row 0 = first step = largest t), each value rounded
        to fp32 exactly where the reference materialises an fp32 tensor (ddim.py:244-271)."""
        S = len(self.timesteps)
        out = np.zeros((S, 6), dtype=np.float32)
        for i in range(S):
            index = S - 1 - i
            t = int(self.timesteps[index])
            a_prev = np.float32(self.alphas_prev[index])
            sigma = np.float32(self.sigmas[index])
            rescale = np.float32(1.0)
            if self.scale is not None:
                rescale = np.float32(self.scale_prev[index]) / np.float32(self.scale[index])
            out[i] = (np.float32(sqrt_ac[t]), np.float32(sqrt_1mac[t]), rescale, np.sqrt(a_prev),
                      np.sqrt(np.float32(1.0) - a_prev - sigma * sigma), sigma)
        return out
