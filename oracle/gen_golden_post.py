#!/usr/bin/env python
"""TEST INFRASTRUCTURE (run in the build container, where /root/reference is mounted; never on the GPU box).

Pins the per-window post-processing and the ray-map -> camera conversion (a14) of oracle/align.py
against the reference's OWN code and writes tests/golden/post_ref.pt:

* `raymap_to_camera_matrix` (scripts/evaluation/infer_geo4d.py:657-674) is executed from the reference source; it
  calls the reference's `utils.rays.cameras_from_plucker` -> `rays_to_cameras` -> `utils.normalize.
  intersect_skew_lines_high_dim`, imported unmodified with two shims for packages that are not installed here
  (`ipdb`: never reached; `pytorch3d.renderer.PerspectiveCameras`: an R / T / focal container, only `.R`, `.T`,
  `clone()` and `len()` are used on this path, rays.py:330-364).
* `get_sky_mask`, `get_far_away_mask`, `denormalize_pc_bbox2` (infer_geo4d.py:83-88, 275-286) are executed from the
  reference source (the script itself cannot be imported: decord / av / open_clip are missing); the dozen glue lines
  between them (infer_geo4d.py:461-487: channel split, Softplus, masks -> confidence 999 -> 1/conf -> 0) are
  restated here around those calls, each with its line number.
"""
import ast
import os
import sys
import types

import torch
from einops import rearrange

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GEO4D_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def install_shims():
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))

    class PerspectiveCameras:
        def __init__(self, focal_length=(1.0,), device="cpu", R=None, T=None):
            n = len(focal_length)
            self.focal_length = torch.as_tensor(focal_length, dtype=torch.float32, device=device).reshape(n, -1)
            self.R = torch.eye(3, device=device).repeat(n, 1, 1) if R is None else R
            self.T = torch.zeros(n, 3, device=device) if T is None else T
            self.device = device

        def __len__(self):
            return self.R.shape[0]

        def clone(self):
            return PerspectiveCameras(tuple(self.focal_length.reshape(-1).tolist()), self.device, self.R.clone(),
                                      self.T.clone())

    p3 = types.ModuleType("pytorch3d")
    rend = types.ModuleType("pytorch3d.renderer")
    rend.PerspectiveCameras = PerspectiveCameras
    rend.RayBundle = object
    tr = types.ModuleType("pytorch3d.transforms")     # utils/normalize.py:7 imports two names it never uses here
    tr.Rotate = tr.Translate = object
    p3.__path__ = []
    p3.renderer, p3.transforms = rend, tr
    sys.modules["pytorch3d"] = p3
    sys.modules["pytorch3d.renderer"] = rend
    sys.modules["pytorch3d.transforms"] = tr


def reference_functions(path, names, namespace):
    """exec the named top-level functions of a reference file (source taken verbatim from the file)"""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), namespace)
    missing = [n for n in names if n not in namespace]
    assert not missing, missing
    return namespace


def synthetic_maps(T=5, H=24, W=40, seed=0):
    """11-channel decoded maps of a moving pinhole camera over a smooth scene, plus sky / far-away / noise pixels,
    so that every branch of the post-processing is exercised.  Layout as produced by the VAE decodes
    (infer_geo4d.py:256-268): xyz (normalised), raw confidence, ray directions, ray moments, inverse depth."""
    g = torch.Generator().manual_seed(seed)
    f = 0.9 * max(H, W)
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dirs_cam = torch.stack([(u - W / 2 + 0.5) / f, (v - H / 2 + 0.5) / f, torch.ones_like(u)], -1)
    maps = torch.zeros(1, 11, T, H, W)
    for t in range(T):
        ang = 0.05 * t
        R = torch.tensor([[torch.cos(torch.tensor(ang)), 0, torch.sin(torch.tensor(ang))], [0, 1, 0],
                          [-torch.sin(torch.tensor(ang)), 0, torch.cos(torch.tensor(ang))]])
        c = torch.tensor([0.08 * t, -0.02 * t, 0.05 * t])
        d = dirs_cam @ R.T                                          # world-frame ray directions (c2w rotation R)
        d = d / d.norm(dim=-1, keepdim=True)
        m = torch.cross(c.expand_as(d), d, dim=-1)                  # Pluecker moment  c x d
        depth = 2.0 + 0.5 * torch.sin(u / 7) * torch.cos(v / 5) + 0.05 * torch.randn(H, W, generator=g)
        pts = dirs_cam * depth[..., None]
        xyz = torch.stack([pts[..., 0] * 2.0 / 3, pts[..., 1] * 2.0 / 3, pts[..., 2] / 2.5 - 1], -1).clamp(-2.5, 2.5)
        maps[0, 0:3, t] = xyz.permute(2, 0, 1)
        maps[0, 3, t] = torch.randn(H, W, generator=g)              # raw confidence (Softplus input)
        maps[0, 4:7, t] = (d * (1 + 0.01 * torch.randn(H, W, 1, generator=g))).permute(2, 0, 1)
        maps[0, 7:10, t] = (m + 0.002 * torch.randn(H, W, 3, generator=g)).permute(2, 0, 1)
        maps[0, 10, t] = (1.0 / depth).clamp(0, 1) * 2 - 1
    maps[0, 0:3, :, :3, :] = 1.05 + 0.03 * torch.randn(3, T, 3, W, generator=g)     # sky band
    maps[0, 0, :, -2:, :] = 2.2                                                        # far-away rows
    return maps


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    install_shims()
    sys.path.insert(0, REF)
    from utils.rays import cameras_from_plucker                     # the reference's own module
    ns = {"torch": torch, "cameras_from_plucker": cameras_from_plucker}
    reference_functions(os.path.join(REF, "scripts", "evaluation", "infer_geo4d.py"),
                        ["denormalize_pc_bbox2", "get_sky_mask", "get_far_away_mask", "raymap_to_camera_matrix"], ns)
    out = {}
    # (square inputs are not a case: the reference leaves num_patches_x unset when H == W, rays.py:399-414)
    for name, (T, H, W, seed) in {"wide": (5, 24, 40, 0), "tall": (3, 40, 24, 1), "wide16": (16, 20, 32, 2)}.items():
        batch_samples = synthetic_maps(T, H, W, seed)
        bs = batch_samples.clone()
        # ---- infer_geo4d.py:447-487 (modality pc_ray_cross_depth, use_raymap / crossmap / inverse depth / traj)
        raymap, crossmap = bs[:, 4:7], bs[:, 7:10]                                  # :447-448
        traj = ns["raymap_to_camera_matrix"](raymap, crossmap)                      # :449  (reference code)
        inverse_depthmap = rearrange(bs[:, 10:11], "b c t h w -> (b t) c h w")      # :454-455
        inverse_depthmap = rearrange(inverse_depthmap, "t c h w -> t h w c")        # :456
        inverse_depthmap = (inverse_depthmap + 1.0) / 2.0                           # :457
        x_recon = rearrange(bs[:, :4], "b c t h w -> (b t) c h w")                  # :460, :464
        confidence = torch.nn.Softplus()(x_recon[:, [-1], :, :])                    # :465-467
        confidence = rearrange(confidence, "t c h w -> t h w c")                    # :468
        x_recon = x_recon[:, :-1, :, :]                                             # :472
        x_recon_reshape = rearrange(x_recon, "t c h w -> t h w c")                  # :474
        invalid_pts = ns["get_sky_mask"](x_recon_reshape, sky_value=1.05, eps=0.1)  # :477  (reference code)
        far_away_mask = ns["get_far_away_mask"](x_recon_reshape, far_away_value=1.99)   # :478  (reference code)
        invalid_pts = invalid_pts | far_away_mask                                   # :479
        confidence[invalid_pts] = 999.0                                             # :480
        inverse_confidence = 1 / confidence                                         # :483
        inverse_confidence[invalid_pts] = 0.0                                       # :484
        x_recon = rearrange(x_recon, "t c h w -> t h w c")                          # :485
        x_recon = ns["denormalize_pc_bbox2"](x_recon, alpha=2.0, beta=2.0)          # :486  (reference code)
        out[name] = {"maps": batch_samples, "pts3d": x_recon, "conf": inverse_confidence,
                     "inverse_depthmap": inverse_depthmap, "traj": traj, "valid": ~invalid_pts}
    # ---- compare the oracle restatement
    from oracle import align as oa
    report = {}
    for name, r in out.items():
        o = oa.postprocess_window(r["maps"])
        report[name] = {
            "pts3d_max_abs": float((o["pts3d"] - r["pts3d"]).abs().max()),
            "conf_max_abs": float((o["conf"] - r["conf"]).abs().max()),
            "invdepth_max_abs": float((o["inverse_depthmap"] - r["inverse_depthmap"]).abs().max()),
            "valid_mismatch": int((o["valid"] != r["valid"]).sum()),
            "traj_max_abs": float((o["traj"] - r["traj"]).abs().max()),
            "invalid_fraction": float((~r["valid"]).float().mean()),
        }
        print(name, report[name])
    torch.save(out, os.path.join(GOLD, "post_ref.pt"))
    import json
    json.dump(report, open(os.path.join(GOLD, "gen_report_post.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
