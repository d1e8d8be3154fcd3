#!/usr/bin/env python
"""Pin the oracle against the reference and (re)generate tests/golden/*.

Run in the BUILD container only (needs /root/reference, read-only):

    python oracle/gen_golden.py

It imports the reference's own modules (with import shims for packages that are
not installed offline: pytorch_lightning, ...), loads the oracle's seeded
synthetic state dicts into them with strict=True (pins key names and shapes),
runs reference and oracle on the same seeded inputs, asserts agreement, and
stores the REFERENCE outputs as small fixtures.  tests/ then check the oracle
(and, on the GPU box, the CUDA path) against these fixtures without needing
/root/reference.

Nothing from the reference is copied: only tensors it computed.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GEO4D_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def install_shims():
    """Stand-ins for packages missing offline so reference files import unmodified."""
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

    pl.LightningModule = LightningModule
    pl.LightningDataModule = object
    pl.seed_everything = lambda s: torch.manual_seed(s)
    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_only = lambda f: f
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util
    for name in ("ipdb",):
        sys.modules.setdefault(name, types.ModuleType(name))


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def install_alignment_shims():
    """roma / evo / plotting packages are not installed offline.  roma and the two evo calls are routed to the
    restatements in oracle/align.py (so those stay 'parity unpinned'); everything else in the reference's
    dust3r/cloud_opt runs unmodified on CPU."""
    from oracle import align as oa

    roma = types.ModuleType("roma")
    roma.rigid_points_registration = lambda x, y, weights=None, compute_scaling=False: \
        oa.rigid_points_registration(x, y, weights, compute_scaling)
    roma.rotmat_to_unitquat = oa.rotmat_to_unitquat

    class RigidUnitQuat:
        def __init__(self, Q, T):
            self.Q, self.T = Q, T

        def normalize(self):
            return RigidUnitQuat(self.Q / self.Q.norm(dim=-1, keepdim=True), self.T)

        def to_homogeneous(self):
            R = oa.unitquat_to_rotmat(self.Q)
            top = torch.cat([R, self.T.unsqueeze(-1)], -1)
            bot = torch.zeros(*self.Q.shape[:-1], 1, 4, dtype=self.Q.dtype)
            bot[..., 0, 3] = 1
            return torch.cat([top, bot], -2)

    roma.RigidUnitQuat = RigidUnitQuat
    sys.modules["roma"] = roma

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            m = _Any(self.__name__ + "." + k)
            setattr(self, k, m)
            return m

        def __call__(self, *a, **k):
            return _Any("x")

    for name in ["evo", "evo.core", "evo.core.sync", "evo.core.metrics", "evo.core.trajectory", "evo.tools",
                 "evo.tools.file_interface", "evo.tools.plot", "evo.main_ape", "evo.main_rpe", "evo.core.geometry",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "seaborn", "trimesh",
                 "pytorch3d", "pytorch3d.renderer", "pytorch3d.transforms", "imageio", "decord", "av", "gdown",
                 "timm", "open_clip", "kornia", "omegaconf"]:
        if name not in sys.modules:
            sys.modules[name] = _Any(name)


def pin_alignment(report):
    """Run the reference's LightPointCloudGroupOptimizer (CPU) and the oracle restatement on the same synthetic
    scene; store the reference's outputs."""
    from oracle import align as oa
    install_alignment_shims()
    import dust3r.cloud_opt.optimizer_group as og
    from scipy.spatial.transform import Rotation

    def tum_to_mats(tum):
        poses = np.tile(np.eye(4), (len(tum), 1, 1))
        for i, p in enumerate(tum):
            poses[i, :3, 3] = p[:3]
            qw, qx, qy, qz = p[3:]
            poses[i, :3, :3] = Rotation.from_quat([qx, qy, qz, qw]).as_matrix()
        return poses

    def align_traj(pred_traj, gt_traj, correct_scale=True, return_aligned_traj=False, align_origin=False):
        assert align_origin and not correct_scale
        est, ref = tum_to_mats(pred_traj[0]), tum_to_mats(gt_traj[0])
        P, rpe_rot = oa.align_origin_and_rpe_rot(est, ref)
        return 0.0, 0.0, rpe_rot, P, None

    og.align_trajectory_with_eval = align_traj
    torch.Tensor.cuda = lambda self, *a, **k: self  # depth_evaluation(use_gpu=True), depth_eval.py:180-182
    T, H, W, niter, start_b = 24, 32, 48, 60, 20
    groups, preds, gt = oa.synthetic_scene(T=T, H=H, W=W, noise=0.003)
    views = [[{"idx": (i,)} for i in g] for g in groups]
    torch.manual_seed(0)
    scene = og.LightPointCloudGroupOptimizer(
        views, [dict(p) for p in preds], conf="id", conf_optimize=True, verbose=False, shared_focal=True,
        flow_loss_weight=0.0, flow_loss_fn="l1", depth_regularize_weight=0.0, num_total_iter=niter,
        temporal_smoothing_weight=0.015, motion_mask_thre=0.35, flow_loss_start_epoch=0.1, flow_loss_thre=20,
        translation_weight=1.0, sintel_ckpt=True, use_self_mask=True, sam2_mask_refine=False, empty_cache=False,
        pxl_thre=50.0, depth_traj_start_iter=start_b)
    loss = scene.compute_global_alignment(init="group", niter=niter, schedule="linear", lr=0.03)
    ref = {"depth": torch.stack(scene.get_depthmaps()).detach(), "poses": scene.get_im_poses().detach(),
           "focal": float(scene.get_focals()[0]), "pw_poses": scene.get_pw_poses().detach(),
           "s_depth": scene.s_depth.detach().clone(), "t_depth": scene.t_depth.detach().clone(),
           "valid_traj": list(scene.valid_traj_group_list), "invalid_depth": list(scene.invalid_depth_group),
           "loss": float(loss)}
    al = oa.GroupAligner(groups, preds, depth_traj_start_iter=start_b, lad_max_iters=5000)
    al.compute_global_alignment(niter=niter, lr=0.03, schedule="linear")
    r = al.results()
    errs = {"depth_absrel": float(((r["depth"] - ref["depth"]).abs() / ref["depth"]).mean()),
            "pose_t_max": float((r["poses"][:, :3, 3] - ref["poses"][:, :3, 3]).norm(dim=-1).max()),
            "pose_R_max": float((r["poses"][:, :3, :3] - ref["poses"][:, :3, :3]).abs().max()),
            "focal_rel": abs(r["focal"] - ref["focal"]) / ref["focal"],
            "s_depth": float((r["s_depth"] - ref["s_depth"]).abs().max()),
            "t_depth": float((r["t_depth"] - ref["t_depth"]).abs().max())}
    report["align_oracle_vs_reference"] = errs
    assert al.valid_traj_groups == ref["valid_traj"] and al.invalid_depth_group == ref["invalid_depth"]
    assert errs["depth_absrel"] < 1e-3 and errs["pose_t_max"] < 1e-3 and errs["pose_R_max"] < 1e-3, errs
    assert errs["focal_rel"] < 1e-3 and errs["s_depth"] < 1e-2 and errs["t_depth"] < 1e-2, errs
    ref["scene"] = dict(T=T, H=H, W=W, noise=0.003, niter=niter, start_b=start_b)
    torch.save(ref, os.path.join(GOLD, "align_ref.pt"))


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    os.makedirs(GOLD, exist_ok=True)
    install_shims()
    sys.path.insert(0, REF)
    torch.set_num_threads(os.cpu_count() or 8)
    report = {}

    # ------------------------------------------------------------------ U-Net
    from lvdm.modules.networks.openaimodel3d import UNetModel
    from oracle import unet as ou

    def ref_unet(cfg: ou.UNetConfig, device="cpu"):
        with torch.device(device):
            return UNetModel(
                in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                model_channels=cfg.model_channels,
                attention_resolutions=list(cfg.attention_resolutions),
                num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult),
                dropout=0.1, num_head_channels=cfg.num_head_channels,
                transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                use_linear=cfg.use_linear, use_checkpoint=False,
                temporal_conv=cfg.temporal_conv, temporal_attention=cfg.temporal_attention,
                temporal_selfatt_only=True, use_relative_position=False,
                use_causal_attention=False, temporal_length=cfg.temporal_length,
                addition_attention=cfg.addition_attention,
                image_cross_attention=cfg.image_cross_attention,
                default_fs=cfg.default_fs, fs_condition=cfg.fs_condition).eval()

    # full config: key/shape inventory on the meta device (1516 tensors, 1438.9 M params)
    full = ou.UNetConfig()
    ref_full = ref_unet(full, "meta")
    ref_shapes = {k: tuple(v.shape) for k, v in ref_full.state_dict().items()}
    my_shapes = dict(ou.param_shapes(full))
    assert list(ref_shapes.keys()) == list(my_shapes.keys()), "U-Net key order/name mismatch"
    assert ref_shapes == my_shapes, "U-Net shape mismatch"
    nparam = sum(int(np.prod(s)) for s in my_shapes.values())
    report["unet_full"] = {"tensors": len(my_shapes), "params": nparam}
    with open(os.path.join(GOLD, "unet_full_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in ref_shapes.items()}, f)
    del ref_full

    tiny = ou.UNetConfig.tiny()
    sd = ou.init_params(ou.param_shapes(tiny), seed=0)
    net = ref_unet(tiny)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    b, t, hh, ww = 1, tiny.temporal_length, 8, 16
    x = torch.randn(b, tiny.in_channels, t, hh, ww, generator=g)
    ctx = torch.randn(b, 77 + 16 * t, tiny.context_dim, generator=g)
    ts = torch.tensor([499], dtype=torch.long)
    fs = torch.tensor([24], dtype=torch.long)
    with torch.no_grad():
        y_ref = net(x, ts, context=ctx, fs=fs)
    y_or = ou.forward(tiny, sd, x, ts, ctx, fs)
    e = rel_l2(y_or, y_ref)
    report["unet_tiny_rel_l2"] = e
    assert e < 1e-5, e
    torch.save({"x": x, "context": ctx, "timesteps": ts, "fs": fs, "y": y_ref,
                "cfg": dict(model_channels=64, context_dim=64, temporal_length=4), "seed": 0,
                "w_checksum": float(sum(v.double().sum() for v in sd.values()))},
               os.path.join(GOLD, "unet_tiny.pt"))
    # batch 2 + default fs path + second timestep
    x2 = torch.randn(2, tiny.in_channels, t, hh, ww, generator=g)
    ctx2 = torch.randn(2, 77 + 16 * t, tiny.context_dim, generator=g)
    ts2 = torch.tensor([999, 19], dtype=torch.long)
    with torch.no_grad():
        y2_ref = net(x2, ts2, context=ctx2, fs=None)
    e2 = rel_l2(ou.forward(tiny, sd, x2, ts2, ctx2, None), y2_ref)
    report["unet_tiny_b2_rel_l2"] = e2
    assert e2 < 1e-5, e2
    del net

    # ------------------------------------------------------------------ VAE
    from lvdm.modules.networks.ae_modules import Encoder, Decoder
    from lvdm.models.autoencoder_adaptor import VAEDecoderadaptor
    from oracle import vae as ov

    def ref_vae(cfg: ov.VAEConfig, device="cpu"):
        dd = dict(double_z=True, z_channels=cfg.z_channels, resolution=256,
                  in_channels=cfg.in_channels, out_ch=cfg.out_ch, ch=cfg.ch,
                  ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=[], dropout=0.0)
        ad = dict(double_z=True, z_channels=cfg.z_channels, resolution=256, in_channels=3,
                  out_ch=cfg.adaptor_out_ch, ch=cfg.adaptor_ch, ch_mult=[1],
                  num_res_blocks=cfg.adaptor_res_blocks, attn_resolutions=[], dropout=0.0)
        with torch.device(device):
            m = torch.nn.Module()
            m.encoder = Encoder(**dd)
            m.decoder = Decoder(**dd)
            m.quant_conv = torch.nn.Conv2d(2 * cfg.z_channels, 2 * cfg.embed_dim, 1)
            m.post_quant_conv = torch.nn.Conv2d(cfg.embed_dim, cfg.z_channels, 1)
            m.decoder_adaptor = VAEDecoderadaptor(**ad)
        return m.eval()

    vfull = ov.VAEConfig()
    rv = ref_vae(vfull, "meta")
    ref_vs = {k: tuple(v.shape) for k, v in rv.state_dict().items()}
    my_vs = dict(ov.param_shapes(vfull))
    assert set(ref_vs.keys()) == set(my_vs.keys()), (set(ref_vs) ^ set(my_vs))
    assert ref_vs == my_vs
    report["vae_full"] = {"tensors": len(my_vs),
                          "params": sum(int(np.prod(s)) for s in my_vs.values())}
    with open(os.path.join(GOLD, "vae_full_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in ref_vs.items()}, f)

    vt = ov.VAEConfig.tiny()
    vsd = ou.init_params(ov.param_shapes(vt), seed=3)
    rv = ref_vae(vt)
    rv.load_state_dict(vsd, strict=True)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 64, 64, generator=g)
    z = torch.randn(2, 4, 8, 8, generator=g)
    with torch.no_grad():
        mom_ref = rv.quant_conv(rv.encoder(img))
        dec_ref = rv.decoder(rv.post_quant_conv(z))
        rv.decoder.give_pre_and_end = True
        rgb_ref, pre_ref = rv.decoder(rv.post_quant_conv(z))
        rv.decoder.give_pre_and_end = False
        conf_ref = rv.decoder_adaptor(pre_ref)
    e_enc = rel_l2(ov.encode_moments(vt, vsd, img), mom_ref)
    e_dec = rel_l2(ov.decode(vt, vsd, z), dec_ref)
    e_conf = rel_l2(ov.decode_with_conf_adaptor(vt, vsd, z), torch.cat([rgb_ref, conf_ref], 1))
    report["vae_tiny_rel_l2"] = {"encode": e_enc, "decode": e_dec, "decode_conf": e_conf}
    assert max(e_enc, e_dec, e_conf) < 1e-5, report["vae_tiny_rel_l2"]
    torch.save({"img": img, "z": z, "moments": mom_ref, "dec": dec_ref,
                "dec_conf": torch.cat([rgb_ref, conf_ref], 1), "seed": 3,
                "cfg": dict(ch=64, adaptor_ch=64)}, os.path.join(GOLD, "vae_tiny.pt"))

    # ------------------------------------------------------------------ schedule + DDIM
    from lvdm.models.utils_diffusion import (make_beta_schedule, rescale_zero_terminal_snr,
                                             make_ddim_timesteps, timestep_embedding)
    from lvdm.models.samplers.ddim import DDIMSampler
    from lvdm.common import extract_into_tensor
    from oracle import ddim as od

    betas = rescale_zero_terminal_snr(make_beta_schedule("linear", 1000, 0.00085, 0.012))
    sch = od.Schedule.geo4d()
    assert np.allclose(sch.betas, betas.astype(np.float32), rtol=0, atol=0)
    for m, S in (("uniform_trailing", 50), ("uniform_trailing", 5), ("uniform_trailing", 2),
                 ("uniform", 50)):
        assert (make_ddim_timesteps(m, S, 1000, verbose=False) == od.make_ddim_timesteps(m, S)).all()
    te = timestep_embedding(torch.tensor([999, 24]), 320)
    assert torch.equal(te, ou.timestep_embedding(torch.tensor([999, 24]), 320))

    class StubModel:
        """Just the attributes DDIMSampler reads (ddim.py:14,27-37,231,262)."""
        num_timesteps = 1000
        parameterization = "v"
        use_dynamic_rescale = True
        device = torch.device("cpu")

        def __init__(self, fn):
            self.betas = torch.tensor(sch.betas)
            self.alphas_cumprod = torch.tensor(sch.alphas_cumprod)
            self.alphas_cumprod_prev = torch.tensor(sch.alphas_cumprod_prev)
            self.sqrt_alphas_cumprod = torch.tensor(sch.sqrt_alphas_cumprod)
            self.sqrt_one_minus_alphas_cumprod = torch.tensor(sch.sqrt_one_minus_alphas_cumprod)
            self.scale_arr = torch.tensor(sch.scale_arr)
            self.fn = fn

        def apply_model(self, x, t, c, **kw):
            return self.fn(x, t)

        # ddpm3d.py:278-290 (restated here because ddpm3d needs the full lightning stack)
        def predict_start_from_z_and_v(self, x_t, t, v):
            return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t -
                    extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

        def predict_eps_from_z_and_v(self, x_t, t, v):
            return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * v +
                    extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * x_t)

    class CPUSampler(DDIMSampler):
        def register_buffer(self, name, attr):  # ddim.py:18-22 forces CUDA
            setattr(self, name, attr)

    g = torch.Generator().manual_seed(11)
    W = torch.randn(16, 16, generator=g) * 0.3

    def toy_model(x, t):
        # a deterministic, t-dependent nonlinear "denoiser"
        s = (t.float() / 1000.0).reshape(-1, 1, 1, 1, 1)
        return torch.tanh(torch.einsum("oc,bcthw->bothw", W, x)) * (0.5 + s) - 0.1 * x

    x_T = torch.randn(1, 16, 4, 4, 8, generator=g)
    gold_ddim = {"W": W, "x_T": x_T}
    for S in (2, 5, 50):
        smp = CPUSampler(StubModel(toy_model))
        out_ref, _ = smp.sample(S=S, batch_size=1, shape=(16, 4, 4, 8), conditioning=None,
                                eta=0.0, verbose=False, x_T=x_T,
                                timestep_spacing="uniform_trailing")
        out_or, _ = od.ddim_sample(toy_model, x_T, sch, S)
        e = rel_l2(out_or, out_ref)
        report[f"ddim_S{S}_rel_l2"] = e
        assert e < 1e-5, (S, e)
        gold_ddim[f"out_S{S}"] = out_ref
        tab = od.make_ddim_tables(sch, S)
        assert np.allclose(tab.alphas, np.asarray(smp.ddim_alphas), atol=0)
        assert np.allclose(tab.alphas_prev, np.asarray(smp.ddim_alphas_prev), atol=0)
        assert np.allclose(tab.scale, smp.ddim_scale_arr.numpy(), atol=0)
        assert np.allclose(tab.scale_prev, smp.ddim_scale_arr_prev.numpy(), atol=0)
    torch.save(gold_ddim, os.path.join(GOLD, "ddim_toy.pt"))
    kat = {
        "alphas_cumprod": {str(i): float(sch.alphas_cumprod[i]) for i in (0, 1, 499, 998, 999)},
        "ddim_timesteps_50_first3_last2": od.make_ddim_timesteps("uniform_trailing", 50)[[0, 1, 2, -2, -1]].tolist(),
        "scale_arr_19_39_59": [float(sch.scale_arr[i]) for i in (19, 39, 59)],
        "timestep_embedding_sum_999_24": te.sum(1).tolist(),
    }
    with open(os.path.join(GOLD, "schedule_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    report["schedule_kat"] = kat

    pin_alignment(report)

    with open(os.path.join(GOLD, "gen_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
