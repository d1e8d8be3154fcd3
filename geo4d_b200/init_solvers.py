"""Host-side initialisation solvers of the alignment, called exactly like the reference calls them
(GPU ports are the next row, the reference runs them on the CPU too):

* `focal_per_group`: per-window focal from the reference frame's point map -- z-shift by scipy
  Levenberg-Marquardt (utils/geometry.py point_map_to_depth:162-215, solve_optimal_shift_focal:232-270,
  image_plane_uv:217-230) and the outlier clamp of align_group_prefix (init_im_poses.py:244-271);
* `fast_pnp`: RANSAC-PnP per frame over three tentative focals (init_im_poses.py:824-865,
  cv2.solvePnPRansac with SOLVEPNP_SQPNP).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


def _image_plane_uv(width: int, height: int) -> np.ndarray:
    ar = width / height
    sx = ar / (1 + ar ** 2) ** 0.5
    sy = 1 / (1 + ar ** 2) ** 0.5
    u = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, width)
    v = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, height)
    u, v = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([u, v], dim=-1).numpy()


def _solve_shift_focal(uv: np.ndarray, xyz: np.ndarray):
    from scipy.optimize import least_squares
    uv, xy, z = uv.reshape(-1, 2), xyz[..., :2].reshape(-1, 2), xyz[..., 2].reshape(-1)

    def residual(shift):
        proj = xy / (z + shift)[:, None]
        f = (proj * uv).sum() / np.square(proj).sum()
        return (f * proj - uv).ravel()

    shift = least_squares(residual, x0=0, ftol=1e-3, method="lm")["x"].squeeze().astype(np.float32)
    proj = xy / (z + shift)[:, None]
    return shift, (proj * uv).sum() / (proj * proj).sum()


def focal_per_group(ref_pointmap: torch.Tensor, ref_conf: torch.Tensor) -> List[float]:
    """ref_pointmap [G, H, W, 3], ref_conf [G, H, W] (CPU) -> focal in pixels per window."""
    G, H, W, _ = ref_pointmap.shape
    mask = (ref_conf > 0.5).numpy()
    pm = ref_pointmap.clone()
    pm[..., 2] = pm[..., 2] - pm[..., 2].min() + 1
    pm = pm.numpy()
    uv = _image_plane_uv(W, H)
    diag = (H ** 2 + W ** 2) ** 0.5
    foc = torch.tensor([float(_solve_shift_focal(uv[mask[i]], pm[i][mask[i]])[1]) for i in range(G)],
                       dtype=torch.float32)
    fx = 0.5 / torch.tan(torch.atan(W / diag / foc))
    fy = 0.5 / torch.tan(torch.atan(H / diag / foc))
    focal_group = ((fx * W) + (fy * H)) / 2
    mean_f = focal_group[focal_group > 30].mean()
    rel = torch.abs(focal_group - mean_f) / mean_f
    focal_group[rel > 0.6] = mean_f
    return focal_group.numpy().tolist()


def fast_pnp(pts3d: np.ndarray, focal: Optional[float], msk: np.ndarray, niter_PnP: int = 10):
    """pts3d [H, W, 3], msk [H, W] bool -> (best focal, cam-to-world 4x4) or None."""
    import cv2
    if msk.sum() < 4:
        return None
    H, W, _ = pts3d.shape
    pixels = np.mgrid[:W, :H].T.astype(np.float32)
    S = max(W, H)
    if focal is None:
        tentative = np.geomspace(S / 2, S * 3, 63)
    else:
        tentative = [focal] + list(np.geomspace(-0.03 * S + focal, 0.03 * S + focal, 2))
    best = (0,)
    for f in tentative:
        K = np.float32([(f, 0, W / 2), (0, f, H / 2), (0, 0, 1)])
        ok, R, T, inl = cv2.solvePnPRansac(pts3d[msk], pixels[msk], K, None, iterationsCount=niter_PnP,
                                           reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP)
        if ok and len(inl) > best[0]:
            best = (len(inl), R, T, f)
    if not best[0]:
        return None
    _, R, T, bf = best
    w2c = np.eye(4)
    w2c[:3, :3] = cv2.Rodrigues(R)[0]
    w2c[:3, 3] = np.asarray(T).ravel()
    return bf, np.linalg.inv(w2c)
