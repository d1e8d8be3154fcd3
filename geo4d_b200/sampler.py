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
    multicond = False   # DDIMSampler_multicond: 3-way guidance of ddim_multiplecond.py:226-236

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
        # guidance extras of the evaluation script (infer_geo4d.py:188-226): the multi-condition sampler consumes
        # them (ddim_multiplecond.py:214-236), the plain sampler forwards them to a U-Net that swallows them
        cfg_img = kwargs.pop("cfg_img", None)
        uc_img = kwargs.pop("unconditional_conditioning_img_nonetext", None)
        kwargs = {k: v for k, v in kwargs.items() if v is not None}
        self.make_schedule(ddim_num_steps=S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=schedule_verbose)
        if len(shape) == 3:
            size = (batch_size, *shape)
        else:
            C, T, H, W = shape
            size = (batch_size, C, T, H, W)
        device = self.model.betas.device
        img = torch.randn(size, device=device) if x_T is None else x_T.to(device)
        # guidance plan: conditionings to evaluate and how to mix their outputs
        #   plain : v = v_c
        #   cfg   : v = v_u + s (v_c - v_u)                                   (ddim.py:216-229)
        #   multi : v = v_u + s_img (v_ui - v_u) + s (v_c - v_ui)              (ddim_multiplecond.py:226-236)
        guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        conds, plan = [conditioning], ("plain",)
        if guided and self.multicond:
            if uc_img is None:
                raise ValueError("DDIMSampler_multicond needs unconditional_conditioning_img_nonetext "
                                 "(infer_geo4d.py:188-194: pass cfg_img != 1)")
            conds = [conditioning, unconditional_conditioning, uc_img]
            plan = ("multi", float(unconditional_guidance_scale),
                    float(unconditional_guidance_scale if cfg_img is None else cfg_img), float(guidance_rescale))
        elif guided:
            conds = [conditioning, unconditional_conditioning]
            plan = ("cfg", float(unconditional_guidance_scale), float(guidance_rescale))
        graph_ok = (self.use_cuda_graph and mask is None and len(size) == 5 and noise_dropout == 0.
                    and self.model.parameterization == "v" and all(isinstance(c, dict) for c in conds)
                    and self.model.model.conditioning_key == "hybrid" and score_corrector is None
                    and not quantize_x0 and callback is None and img_callback is None and not kwargs)
        if graph_ok:
            return self._sample_graph(img, conds, plan, fs, log_every_t, eta, temperature)
        if plan[0] == "multi":
            raise NotImplementedError("3-way guidance is implemented on the CUDA-graph path only")
        return self._sample_eager(img, conditioning, fs, log_every_t, unconditional_guidance_scale,
                                  unconditional_conditioning, guidance_rescale, temperature, noise_dropout,
                                  mask, x0, callback, img_callback, **kwargs)

    # ------------------------------------------------------------------ graph path
    @staticmethod
    def _mix(v, b, which, plan):
        """guidance mix of the per-conditioning U-Net outputs v [P*b, ...]; which[i] = slot of conditioning i"""
        if plan[0] == "plain" or len(set(which)) == 1:
            # identical conditionings give bit-identical outputs, for which every mix above returns v_c exactly
            return v[which[0] * b:(which[0] + 1) * b]
        e = [v[w * b:(w + 1) * b] for w in which]
        if plan[0] == "cfg":
            out = e[1] + plan[1] * (e[0] - e[1])
            rescale = plan[2]
        else:
            out = e[1] + plan[2] * (e[2] - e[1]) + plan[1] * (e[0] - e[2])
            rescale = plan[3]
        if rescale > 0.0:
            out = rescale_noise_cfg(out, e[0], rescale)
        return out.contiguous()

    def _sample_graph(self, x_T, conds, plan, fs, log_every_t, eta=0.0, temperature=1.0):
        m = self.model
        unet = m.model.diffusion_model
        S = len(self.ddim_timesteps)
        b, C, T, H, W = x_T.shape
        dev = x_T.device
        # unique conditionings (the shipped settings make the "unconditional" one identical to the conditional one:
        # empty prompt, zero image -- infer_geo4d.py:140-187 -- so guidance then costs ONE U-Net pass, not two)
        uniq, which = [], []
        for c in conds:
            cc = c["c_crossattn"]
            cc = cc[0] if len(cc) == 1 else torch.cat(cc, 1)
            zc = c["c_concat"]
            zc = (zc[0] if len(zc) == 1 else torch.cat(zc, 1)).float().contiguous()
            for j, (cc_j, zc_j) in enumerate(uniq):
                if cc_j.shape == cc.shape and zc_j.shape == zc.shape and torch.equal(cc_j, cc) and torch.equal(zc_j, zc):
                    which.append(j)
                    break
            else:
                which.append(len(uniq))
                uniq.append((cc, zc))
        P = len(uniq)
        cc_all = uniq[0][0] if P == 1 else torch.cat([u[0] for u in uniq], 0)
        zc_all = uniq[0][1] if P == 1 else torch.cat([u[1] for u in uniq], 0)
        if unet._packed is None:
            unet.prepare()
        unet.set_context(cc_all, T)
        # per-step ResBlock embedding rows for all S steps at once: [S, P*b * sum(Cout)]
        steps = np.flip(self.ddim_timesteps).copy()
        ts_all = torch.as_tensor(np.repeat(steps, P * b), device=dev, dtype=torch.long)
        fs_all = None if fs is None else fs.to(dev).repeat(S * P)
        emb_table = unet.embed(ts_all, fs_all, S * P * b).reshape(S, -1).contiguous()
        stochastic = bool(np.any(np.asarray(self.ddim_sigmas) != 0.0))
        key = (b, C, T, H, W, zc_all.shape[1], dev.index, P, tuple(which), plan, stochastic,
               unet._ctx_cache["sig"], id(unet._ctx_cache["kv"]))
        st = self._graphs.get(key)
        if st is None:
            st = {
                "x": torch.empty((b, C, T, H, W), device=dev, dtype=torch.float32),
                "xP": torch.empty((P * b, C, T, H, W), device=dev, dtype=torch.float32) if P > 1 else None,
                "zc": torch.empty_like(zc_all),
                "emb": torch.empty((P * b, emb_table.shape[1] // (P * b)), device=dev, dtype=torch.float32),
                "pred_x0": torch.empty((b, C, T, H, W), device=dev, dtype=torch.float32),
                "idx": torch.zeros(1, device=dev, dtype=torch.int32),
                "coef": torch.empty((1024, 6), device=dev, dtype=torch.float32),
                "noise": torch.empty((b, C, T, H, W), device=dev, dtype=torch.float32) if stochastic else None,
                "noise_table": None,
                "emb_table": None,
                "graph": None,
            }
            self._graphs[key] = st
        st["x"].copy_(x_T)
        st["zc"].copy_(zc_all)
        st["coef"][:S].copy_(torch.as_tensor(self.step_coef, device=dev))
        st["idx"].zero_()
        if st["emb_table"] is None or st["emb_table"].shape != emb_table.shape:
            st["emb_table"] = emb_table.clone()
            st["graph"] = None  # table pointer is baked into the graph
        else:
            st["emb_table"].copy_(emb_table)
        if stochastic:
            # the reference draws sigma_t * randn per step from the global CUDA generator (ddim.py:273-277); the same
            # S draws in the same order, made up front so that the captured step only gathers row `idx`
            noise = torch.stack([torch.randn((b, C, T, H, W), device=dev) for _ in range(S)]).reshape(S, -1)
            if temperature != 1.:
                noise = noise * temperature
            if st["noise_table"] is None or st["noise_table"].shape != noise.shape:
                st["noise_table"] = noise.contiguous()
                st["graph"] = None
            else:
                st["noise_table"].copy_(noise)
        cin_pad = -(-unet.in_channels // 64) * 64

        def one_step():
            ops.gather_row(st["emb_table"], st["idx"], st["emb"])
            xin = st["x"]
            if P > 1:
                st["xP"].view(P, b, C, T, H, W).copy_(st["x"].unsqueeze(0))
                xin = st["xP"]
            rows = ops.bcthw_to_rows(xin, st["zc"], cin_pad)
            v_rows = unet.forward_rows(rows, st["emb"], (P * b, T, H, W))
            v = self._mix(ops.rows_to_bcthw(v_rows, C, P * b, T, H, W), b, which, plan)
            if stochastic:
                ops.gather_row(st["noise_table"], st["idx"], st["noise"])
            ops.ddim_step(st["x"], v, st["coef"], st["idx"], pred_x0=st["pred_x0"], noise=st["noise"])
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


class DDIMSampler_multicond(DDIMSampler):
    """lvdm/models/samplers/ddim_multiplecond.py: text + image guidance with three U-Net evaluations per step
    (conditional, unconditional, image-only), batched into one pass of the captured step."""
    multicond = True
