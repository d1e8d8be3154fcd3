"""End-to-end Geo4D reconstruction glue: video -> per-window diffusion + decode -> per-window
post-processing -> sliding-window global alignment.

Mirrors scripts/evaluation/infer_geo4d.py: `image_guided_synthesis` (:117-273), the per-window block of
`run_evaluation` (:412-500), `raymap_to_camera_matrix` (:657-674) and `post_optimization` (:29-50), with the
same defaults (cfg scale 1, eta 0, 'uniform_trailing', sky eps 0.1, far value 1.99, alpha = beta = 2,
lr 0.03, 500 iterations, linear schedule).  Every array stays on the GPU between stages; the only host hops
are the ones the reference also has inside the alignment initialisation (scipy LM, cv2 PnP).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .cloud_opt import LightPointCloudGroupOptimizer
from .sampler import DDIMSampler, DDIMSampler_multicond

POSTPROCESS_DEFAULTS = dict(not_shared_focal=False, use_gt_focal=False, flow_loss_weight=0.0, flow_loss_fn="l1",
                            depth_regularize_weight=0.0, n_iter=500, temporal_smoothing_weight=0.015,
                            motion_mask_thre=0.35, flow_loss_start_epoch=0.1, flow_loss_thre=20,
                            translation_weight=1.0, eval_dataset="sintel", use_gt_mask=False,
                            sam2_mask_refine=False, pxl_thresh=50.0, pose_schedule="linear", silent=False)


def sliding_windows(T: int, stride: int, window: int = 16) -> List[slice]:
    """infer_geo4d.py:412-418."""
    out = [slice(s, s + window, 1) for s in range(0, T - window + 1, stride)]
    if slice(T - window, T, 1) not in out:
        out.append(slice(T - window, T, 1))
    return out


def raymap_to_camera_matrix(raymap: torch.Tensor, crossmap: torch.Tensor) -> torch.Tensor:
    """[1, 3, t, h, w] ray directions / moments -> cam-to-world [t, 4, 4] = [[R, c], [0, 1]]
    (cameras_from_plucker utils/rays.py:387-433 + infer_geo4d.py:657-674).  The per-pixel reductions run in
    geo4d_raymap_moments; the 3x3 solve / SVD per frame runs on the host in fp64 (16 tiny problems, like the Umeyama
    registrations of the alignment: no cuSOLVER / cuBLAS call on the path)."""
    _, _, T, H, W = raymap.shape
    mom = ops.raymap_moments(raymap[0].contiguous(), crossmap[0].contiguous(), T, H, W).cpu().numpy()  # [T, 18] fp64
    M = np.stack([mom[:, 0], mom[:, 1], mom[:, 2], mom[:, 1], mom[:, 3], mom[:, 4], mom[:, 2], mom[:, 4], mom[:, 5]],
                 -1).reshape(T, 3, 3)
    out = np.tile(np.eye(4), (T, 1, 1))
    for t in range(T):   # 16 independent 3x3 problems: least-squares line intersection + orthogonal Procrustes, fp64
        if not np.isfinite(mom[t]).all():
            continue     # non-finite ray maps: identity pose (LAPACK would raise; the reference would carry NaNs on)
        out[t, :3, 3] = np.linalg.lstsq(M[t], mom[t, 6:9], rcond=None)[0]
        U, _, Vh = np.linalg.svd(mom[t, 9:18].reshape(3, 3))
        sgn = np.sign(np.linalg.det(U @ Vh))
        out[t, :3, :3] = U @ np.diag([1.0, 1.0, sgn]) @ Vh
    return torch.from_numpy(out).to(device=raymap.device, dtype=torch.float32)


class Geo4DPipeline:
    def __init__(self, model, pointmap_vae=None, ddim_steps: int = 50, ddim_eta: float = 0.0,
                 unconditional_guidance_scale: float = 1.0, timestep_spacing: str = "uniform_trailing",
                 guidance_rescale: float = 0.0, postprocess: Optional[dict] = None, seed: int = 123,
                 multiple_cond_cfg: bool = False, cfg_img: Optional[float] = None):
        self.model = model
        self.pointmap_vae = pointmap_vae
        self.ddim_steps = ddim_steps
        self.ddim_eta = ddim_eta
        self.cfg_scale = unconditional_guidance_scale
        self.timestep_spacing = timestep_spacing
        self.guidance_rescale = guidance_rescale
        self.post = dict(POSTPROCESS_DEFAULTS)
        self.post.update(postprocess or {})
        self.seed = seed
        self.multiple_cond_cfg = multiple_cond_cfg      # infer_geo4d.py:119 (DDIMSampler_multicond) / :188-194
        self.cfg_img = cfg_img
        self.sampler = DDIMSampler_multicond(model) if multiple_cond_cfg else DDIMSampler(model)
        self.timings: Dict[str, float] = {}
        self.events: List = []

    # ------------------------------------------------------------------ phase timing (CUDA events)
    class _Phase:
        def __init__(self, pipe, name):
            self.pipe, self.name = pipe, name

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *a):
            self.e1.record()
            self.pipe.events.append((self.name, self.e0, self.e1))

    def phase(self, name):
        return Geo4DPipeline._Phase(self, name)

    def phase_ms(self) -> Dict[str, float]:
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for name, e0, e1 in self.events:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    # ------------------------------------------------------------------ one window
    @torch.no_grad()
    def image_guided_synthesis(self, videos: torch.Tensor, noise_shape: Sequence[int], fs: int = 24,
                               x_T: Optional[torch.Tensor] = None, z_noise: Optional[torch.Tensor] = None,
                               prompts: Optional[List[str]] = None) -> torch.Tensor:
        """videos [b, 3, t, H, W] in [-1, 1] -> decoded maps [b, 1, 11, t, H, W] (modality pc_ray_cross_depth)."""
        m = self.model
        if m.modality != "pc_ray_cross_depth":
            raise NotImplementedError(f"modality {m.modality!r}: Geo4D ships pc_ray_cross_depth")
        b = noise_shape[0]
        dev = m.device
        fs_t = torch.tensor([fs] * b, dtype=torch.long, device=dev)
        cond_emb = m.get_learned_conditioning(prompts or [""] * b)
        if m.cross_attention:
            raise NotImplementedError("per-frame image conditioning needs the OpenCLIP image tower (SURVEY.md N3)")
        img_emb = m.get_image_conditioning(b)  # embedding of the all-zero image, constant
        cond = {"c_crossattn": [torch.cat([cond_emb, img_emb], dim=1)]}
        if m.model.conditioning_key == "hybrid":
            with self.phase("encode"):
                cond["c_concat"] = [m.encode_first_stage(videos, noise=z_noise)]
        uc, uc_2 = None, None
        if self.cfg_scale != 1.0:
            # infer_geo4d.py:168-187: uncond_type 'empty_seq' re-embeds the prompt "" and the image branch embeds an
            # all-zero image again -- with text_input off and cross_attention off these are the SAME tensors as the
            # conditional ones; the sampler recognises identical conditionings and evaluates the U-Net once
            if m.uncond_type == "empty_seq":
                uc_emb = m.get_learned_conditioning([""] * b)
            elif m.uncond_type == "zero_embed":
                uc_emb = torch.zeros_like(cond_emb)
            else:
                raise NotImplementedError(f"uncond_type {m.uncond_type!r}")
            uc = {"c_crossattn": [torch.cat([uc_emb, img_emb], dim=1)]}
            if "c_concat" in cond:
                uc["c_concat"] = cond["c_concat"]
            if self.multiple_cond_cfg and self.cfg_img != 1.0:      # :188-194: image yes, text ""
                uc_2 = {"c_crossattn": [torch.cat([uc_emb, img_emb], dim=1)]}
                if "c_concat" in cond:
                    uc_2["c_concat"] = cond["c_concat"]
        with self.phase("ddim"):
            samples, _ = self.sampler.sample(S=self.ddim_steps, conditioning=cond, batch_size=b,
                                             shape=noise_shape[1:], verbose=False,
                                             unconditional_guidance_scale=self.cfg_scale,
                                             unconditional_conditioning=uc, eta=self.ddim_eta, mask=None, x0=None,
                                             x_T=x_T, fs=fs_t, timestep_spacing=self.timestep_spacing,
                                             guidance_rescale=self.guidance_rescale, cfg_img=self.cfg_img,
                                             unconditional_conditioning_img_nonetext=uc_2)
        with self.phase("decode"):
            out = self.decode_latents(samples).unsqueeze(1)
        return out

    @torch.no_grad()
    def decode_latents(self, samples: torch.Tensor) -> torch.Tensor:
        """infer_geo4d.py:247-257: point map + confidence from pointmap_vae, ray / cross / depth from the
        first-stage VAE (the three plain decodes share weights and run as one 3t-frame batch)."""
        m = self.model
        b, _, t, h, w = samples.shape
        if self.pointmap_vae is not None:
            z = samples[:, 0:4].permute(0, 2, 1, 3, 4).reshape(b * t, 4, h, w) * (1.0 / m.scale_factor)
            pc = self.pointmap_vae.decode_with_conf_adaptor(z)
            pc = pc.reshape(b, t, *pc.shape[1:]).permute(0, 2, 1, 3, 4)
        else:
            pc = m.decode_first_stage_confhead(samples[:, 0:4])
        rest = torch.cat([samples[:, 4:8], samples[:, 8:12], samples[:, 12:16]], dim=2)
        dec = m.decode_first_stage(rest)
        ray, cross, depth = dec[:, :, :t], dec[:, :, t:2 * t], dec[:, :, 2 * t:]
        depth = depth.mean(dim=1, keepdim=True)
        return torch.cat([pc, ray, cross, depth], dim=1)

    @torch.no_grad()
    def window_predictions(self, batch_samples: torch.Tensor, valid: Optional[torch.Tensor] = None) -> dict:
        """infer_geo4d.py:447-500 for one window (b = 1): pred dict for the aligner."""
        assert batch_samples.shape[0] == 1 and batch_samples.shape[1] == 11
        _, _, t, H, W = batch_samples.shape
        with self.phase("post"):
            traj = raymap_to_camera_matrix(batch_samples[:, 4:7], batch_samples[:, 7:10])
            pts, conf, invd = ops.postprocess_window(batch_samples[0].contiguous(), t, H, W, valid=valid,
                                                     sky_eps=0.1, far_value=1.99, has_conf=True)
        return {"pts3d": pts, "conf": conf, "inverse_depthmap": invd, "traj": traj}

    # ------------------------------------------------------------------ whole sequence
    def post_optimization(self, view_list, pred_list, lr: float = 0.03, init_method: str = "group"):
        """infer_geo4d.py:29-50."""
        a = self.post
        scene = LightPointCloudGroupOptimizer(
            view_list, pred_list, conf="id", conf_optimize=True, verbose=not a["silent"],
            shared_focal=not a["not_shared_focal"] and not a["use_gt_focal"], flow_loss_weight=a["flow_loss_weight"],
            flow_loss_fn=a["flow_loss_fn"], depth_regularize_weight=a["depth_regularize_weight"],
            num_total_iter=a["n_iter"], temporal_smoothing_weight=a["temporal_smoothing_weight"],
            motion_mask_thre=a["motion_mask_thre"], flow_loss_start_epoch=a["flow_loss_start_epoch"],
            flow_loss_thre=a["flow_loss_thre"], translation_weight=a["translation_weight"],
            sintel_ckpt=a["eval_dataset"] == "sintel", use_self_mask=not a["use_gt_mask"],
            sam2_mask_refine=a["sam2_mask_refine"], pxl_thre=a["pxl_thresh"], opt_raydir=False)
        with self.phase("align"):
            scene.compute_global_alignment(init=init_method, niter=a["n_iter"], schedule=a["pose_schedule"], lr=lr)
        return scene

    @torch.no_grad()
    def reconstruct(self, videos_all: torch.Tensor, stride: int = 8, windows: Optional[List[slice]] = None,
                    x_T_fn=None, z_noise_fn=None, align: bool = True, keep_images: bool = False):
        """videos_all [1, 3, T, H, W] -> (scene, pred_list).  x_T(w) = randn(seed + w) unless x_T_fn is given
        (per-window seeds make window sharding reproducible)."""
        B, C, T, H, W = videos_all.shape
        assert B == 1, "only support batch size = 1 (infer_geo4d.py:355)"
        dev = videos_all.device
        windows = windows if windows is not None else sliding_windows(T, stride)
        self.last_windows = windows
        h, w = H // 8, W // 8
        ch = self.model.model.diffusion_model.out_channels
        pred_list, view_list = [], []
        valid = torch.ones((T, H, W), dtype=torch.uint8, device=dev)
        t0 = time.time()
        for wi, sl in enumerate(windows):
            videos = videos_all[:, :, sl].contiguous()
            if x_T_fn is not None:
                x_T = x_T_fn(wi)
            else:
                g = torch.Generator(device=dev).manual_seed(self.seed + wi)
                x_T = torch.randn((1, ch, 16, h, w), device=dev, generator=g)
            zn = z_noise_fn(wi) if z_noise_fn is not None else None
            maps = self.image_guided_synthesis(videos, [1, ch, 16, h, w], fs=24 // sl.step, x_T=x_T, z_noise=zn)
            vslice = valid[sl].contiguous()
            pred_list.append(self.window_predictions(maps[:, 0], vslice))
            valid[sl] = vslice
            if keep_images:   # the reference's views carry the frames (save_rgb_imgs); costs a device -> host copy each
                view_list.append([{"img": videos_all[0, :, i], "idx": (i,)} for i in range(sl.start, sl.stop)])
            else:
                view_list.append([{"idx": (i,)} for i in range(sl.start, sl.stop)])
        torch.cuda.synchronize()
        self.timings["diffusion_s"] = time.time() - t0
        scene = None
        if align:
            t1 = time.time()
            with torch.enable_grad():
                scene = self.post_optimization(view_list, pred_list)
            torch.cuda.synchronize()
            self.timings["alignment_s"] = time.time() - t1
        self.timings["frames"] = T
        return scene, pred_list
