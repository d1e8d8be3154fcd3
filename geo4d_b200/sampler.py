"""DDIM sampler behind the reference's `DDIMSampler` seam (lvdm/models/samplers/ddim.py).

Same constructor and `sample(...)` signature/returns as ddim.py:59-132.  Two execution paths:

* graph path (the shipped configuration: v-parameterisation, cfg scale 1 or no unconditional conditioning,
  eta 0, no mask, 'hybrid' conditioning): one U-Net step -- per-step embedding gather, input layout
  conversion, ~1500 kernel launches of the network, output layout conversion, the fused DDIM update
  (geo4d_ddim_step) and the step-counter bump -- is captured ONCE into a CUDA graph and replayed S times with
  no host involvement; all per-step scalars live in device tables indexed by a device counter.
* eager path (classifier-free guidance, eta > 0, masks, other parameterisations): the reference's
  control flow restated on top of `model.apply_model` (ddim.py:134-279).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .schedule import DDIMTables


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """utils_diffusion.py:147-158."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.counter = 0
        self._graphs: Dict[tuple, dict] = {}
        self.use_cuda_graph = kwargs.get("use_cuda_graph", True)

    # ------------------------------------------------------------------ schedule
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        m = self.model
        scale = m.scale_arr.detach().cpu().numpy() if m.use_dynamic_rescale else None
        self.tables = DDIMTables(m.alphas_cumprod.detach().cpu().numpy(), scale, ddim_num_steps, ddim_discretize,
                                 ddim_eta)
        t = self.tables
        self.ddim_timesteps = t.timesteps
        self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sigmas = t.alphas, t.alphas_prev, t.sigmas
        self.ddim_sqrt_one_minus_alphas = t.sqrt_one_minus_alphas
        if scale is not None:
            self.ddim_scale_arr = torch.tensor(t.scale)
            self.ddim_scale_arr_prev = torch.tensor(t.scale_prev)
        self.step_coef = t.step_coefficients(m.sqrt_alphas_cumprod.detach().cpu().numpy(),
                                             m.sqrt_one_minus_alphas_cumprod.detach().cpu().numpy())

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True,
               schedule_verbose=False, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, precision=None, fs=None, timestep_spacing='uniform',
               guidance_rescale=0.0, **kwargs):
        if conditioning is not None:
            c0 = conditioning[list(conditioning.keys())[0]] if isinstance(conditioning, dict) else conditioning
            cbs = (c0[0] if isinstance(c0, (list, tuple)) else c0).shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        # the evaluation script forwards inert extras (cfg_img=None, unconditional_conditioning_img_nonetext=None,
        # infer_geo4d.py:188-226) that the reference U-Net swallows in **kwargs
        kwargs = {k: v for k, v in kwargs.items() if v is not None}
        self.make_schedule(ddim_num_steps=S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=schedule_verbose)
        if len(shape) == 3:
            size = (batch_size, *shape)
        else:
            C, T, H, W = shape
            size = (batch_size, C, T, H, W)
        device = self.model.betas.device
        img = torch.randn(size, device=device) if x_T is None else x_T.to(device)
        no_cfg = unconditional_conditioning is None or unconditional_guidance_scale == 1.
        graph_ok = (self.use_cuda_graph and no_cfg and eta == 0. and mask is None and len(size) == 5
                    and self.model.parameterization == "v" and isinstance(conditioning, dict)
                    and self.model.model.conditioning_key == "hybrid" and score_corrector is None
                    and not quantize_x0 and callback is None and img_callback is None and not kwargs)
        if graph_ok:
            return self._sample_graph(img, conditioning, fs, log_every_t)
        return self._sample_eager(img, conditioning, fs, log_every_t, unconditional_guidance_scale,
                                  unconditional_conditioning, guidance_rescale, temperature, noise_dropout,
                                  mask, x0, callback, img_callback, **kwargs)

    # ------------------------------------------------------------------ graph path
    def _sample_graph(self, x_T, cond, fs, log_every_t):
        m = self.model
        unet = m.model.diffusion_model
        S = len(self.ddim_timesteps)
        b, C, T, H, W = x_T.shape
        dev = x_T.device
        cc = cond["c_crossattn"]
        cc = cc[0] if len(cc) == 1 else torch.cat(cc, 1)
        zc = cond["c_concat"]
        zc = (zc[0] if len(zc) == 1 else torch.cat(zc, 1)).float().contiguous()
        if unet._packed is None:
            unet.prepare()
        unet.set_context(cc, T)
        # per-step ResBlock embedding rows for all S steps at once: [S, b * sum(Cout)]
        steps = np.flip(self.ddim_timesteps).copy()
        ts_all = torch.as_tensor(np.repeat(steps, b), device=dev, dtype=torch.long)
        fs_all = None if fs is None else fs.to(dev).repeat(S)
        emb_table = unet.embed(ts_all, fs_all, S * b).reshape(S, -1).contiguous()
        key = (b, C, T, H, W, zc.shape[1], dev.index)
        st = self._graphs.get(key)
        if st is None:
            st = {
                "x": torch.empty((b, C, T, H, W), device=dev, dtype=torch.float32),
                "zc": torch.empty_like(zc),
                "emb": torch.empty((b, emb_table.shape[1] // b), device=dev, dtype=torch.float32),
                "pred_x0": torch.empty((b, C, T, H, W), device=dev, dtype=torch.float32),
                "idx": torch.zeros(1, device=dev, dtype=torch.int32),
                "coef": torch.empty((1024, 6), device=dev, dtype=torch.float32),
                "emb_table": None,
                "graph": None,
            }
            self._graphs[key] = st
        st["x"].copy_(x_T)
        st["zc"].copy_(zc)
        st["coef"][:S].copy_(torch.as_tensor(self.step_coef, device=dev))
        st["idx"].zero_()
        if st["emb_table"] is None or st["emb_table"].shape != emb_table.shape:
            st["emb_table"] = emb_table.clone()
            st["graph"] = None  # table pointer is baked into the graph
        else:
            st["emb_table"].copy_(emb_table)
        cin_pad = -(-unet.in_channels // 64) * 64

        def one_step():
            ops.gather_row(st["emb_table"], st["idx"], st["emb"])
            rows = ops.bcthw_to_rows(st["x"], st["zc"], cin_pad)
            v_rows = unet.forward_rows(rows, st["emb"], (b, T, H, W))
            v = ops.rows_to_bcthw(v_rows, C, b, T, H, W)
            ops.ddim_step(st["x"], v, st["coef"], st["idx"], pred_x0=st["pred_x0"])
            ops.advance_counter(st["idx"], 1)

        inter = {"x_inter": [x_T], "pred_x0": [x_T]}
        first = 0
        if st["graph"] is None:
            # the first step runs eagerly (also sets kernel attributes); the rest replay the graph
            one_step()
            first = 1
            inter["x_inter"].append(st["x"].clone())
            inter["pred_x0"].append(st["pred_x0"].clone())
            if S > 1:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                n0 = ops.raw_launch_count()
                with torch.cuda.stream(side):
                    with ops.capture_graph(g):
                        one_step()
                torch.cuda.current_stream().wait_stream(side)
                st["graph"] = g
                st["graph_kernels"] = ops.raw_launch_count() - n0
                ops.note_replay(st["graph_kernels"], -1)  # the capture pass recorded, it did not execute
                # capture does not execute: the counter still points at step 1
        for i in range(first, S):
            index = S - i - 1
            st["graph"].replay()
            ops.note_replay(st["graph_kernels"])
            if index % log_every_t == 0 or index == S - 1:
                inter["x_inter"].append(st["x"].clone())
                inter["pred_x0"].append(st["pred_x0"].clone())
        return st["x"].clone(), inter

    # ------------------------------------------------------------------ eager path
    def _sample_eager(self, img, cond, fs, log_every_t, cfg_scale, uc, guidance_rescale, temperature,
                      noise_dropout, mask, x0, callback, img_callback, **kwargs):
        m = self.model
        S = len(self.ddim_timesteps)
        b = img.shape[0]
        dev = img.device
        coef = torch.as_tensor(self.step_coef, device=dev)
        inter = {"x_inter": [img], "pred_x0": [img]}
        if mask is not None:
            raise NotImplementedError("masked sampling (ddim.py:174-181) is not part of the Geo4D inference path")
        img = img.float().contiguous().clone()
        pred_x0 = torch.empty_like(img)
        idx = torch.zeros(1, device=dev, dtype=torch.int32)
        for i, step in enumerate(np.flip(self.ddim_timesteps)):
            index = S - i - 1
            ts = torch.full((b,), int(step), device=dev, dtype=torch.long)
            kw = dict(kwargs)
            if fs is not None:
                kw["fs"] = fs
            if uc is None or cfg_scale == 1.:
                out = m.apply_model(img, ts, cond, **kw)
            else:
                e_c = m.apply_model(img, ts, cond, **kw)
                e_u = m.apply_model(img, ts, uc, **kw)
                out = e_u + cfg_scale * (e_c - e_u)
                if guidance_rescale > 0.0:
                    out = rescale_noise_cfg(out, e_c, guidance_rescale)
            if m.parameterization != "v":
                raise NotImplementedError("only the v-parameterisation is used by Geo4D")
            noise = None
            if float(self.ddim_sigmas[index]) != 0.0:
                noise = torch.randn(img.shape, device=dev) * temperature
                if noise_dropout > 0.:
                    noise = torch.nn.functional.dropout(noise, p=noise_dropout)
            idx.fill_(i)
            ops.ddim_step(img, out.float().contiguous(), coef, idx, pred_x0=pred_x0, noise=noise)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == S - 1:
                inter["x_inter"].append(img.clone())
                inter["pred_x0"].append(pred_x0.clone())
        return img, inter
