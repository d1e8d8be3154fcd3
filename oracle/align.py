"""CPU restatement of the sliding-window alignment (test oracle, torch autograd, fp32).

Follows dust3r/cloud_opt/optimizer_group.py (LightPointCloudGroupOptimizer.__init__:37-107,
forward:440-525, _set_st_depth:333-372, _set_traj:242-267, relative_pose_loss:529-542,
_fast_depthmap_to_pts3d:559-566), base_opt_group.py (_init_from_views:112-200, _get_poses:260-265,
_set_pose:267-288, get_pw_scale/get_pw_poses:303-320, global_alignment_loop/iter:553-626),
init_im_poses.py (init_from_group:60-80, align_group_prefix:226-405, init_from_pts3d_group:569-633,
fast_pnp:824-865), utils/geometry.py (point_map_to_depth:162-215, solve_optimal_shift_focal:232-270,
image_plane_uv:217-230), dust3r/depth_eval.py (absolute_value_scaling2:112-145, depth_evaluation:147-359),
utils/rays.py + utils/normalize.py (cameras_from_plucker:387-433, rays_to_cameras:301-367,
compute_optimal_rotation_alignment:579-595, intersect_skew_lines_high_dim:25-51) and
scripts/evaluation/infer_geo4d.py (per-window post-processing :447-500, raymap_to_camera_matrix:657-674)
of jzr99/Geo4D.

Third-party pieces that are NOT vendored in the reference are restated from their published
algorithms and are "parity unpinned" beyond the analytic known-answer tests in
tests/test_align_oracle.py:
  * roma (requirements.txt:27, unpinned): rigid_points_registration = weighted Umeyama (Umeyama 1991),
    special_procrustes, quaternion <-> rotation in xyzw order;
  * evo (requirements.txt:47, unpinned): PoseTrajectory3D.align_origin (P = T_ref0 T_est0^-1) and the
    delta=1 all-pairs RPE rotation angle RMSE in degrees.
Everything else is pinned by oracle/gen_golden.py, which runs the reference's own
LightPointCloudGroupOptimizer on CPU (roma/evo calls routed to the functions below) on a seeded
synthetic scene and stores its outputs in tests/golden/align_ref.pt.
cv2.solvePnPRansac and scipy.optimize.least_squares are called exactly as the reference calls them.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ============================================================================= roma restatements
def special_procrustes(M: torch.Tensor, return_singular_values: bool = False):
    """Closest rotation to M (det +1): U diag(1,1,det(U)det(V)) V^T."""
    U, D, Vh = torch.linalg.svd(M)
    det = torch.det(U) * torch.det(Vh)
    Dm = torch.ones(3, dtype=M.dtype)
    Dm[2] = det
    R = U @ torch.diag(Dm) @ Vh
    if return_singular_values:
        return R, D * Dm
    return R


def rigid_points_registration(x, y, weights=None, compute_scaling=False):
    """y ~ s R x + t (weighted).  Returns (R, t, s) like roma.rigid_points_registration."""
    if weights is None:
        w = torch.ones(x.shape[0], dtype=x.dtype)
    else:
        w = weights
    sw = w.sum()
    xm = (w[:, None] * x).sum(0) / sw
    ym = (w[:, None] * y).sum(0) / sw
    xh, yh = x - xm, y - ym
    M = (w[:, None] * yh).T @ xh
    if compute_scaling:
        R, DS = special_procrustes(M, return_singular_values=True)
        s = DS.sum() / (w * (xh * xh).sum(-1)).sum()
        t = ym - s * (R @ xm)
        return R, t, s
    R = special_procrustes(M)
    return R, ym - R @ xm


def unitquat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """xyzw unit quaternion(s) [..., 4] -> rotation matrices [..., 3, 3]."""
    x, y, z, w = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)
    return R.reshape(*q.shape[:-1], 3, 3)


def rotmat_to_unitquat(R: torch.Tensor) -> torch.Tensor:
    """Rotation matrix [3,3] -> xyzw unit quaternion (largest-component branch, as scipy/roma do)."""
    m = R.detach().double().numpy()
    d = np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 0] + m[1, 1] + m[2, 2]])
    k = int(np.argmax(d))
    q = np.empty(4)
    if k == 3:
        q[0] = m[2, 1] - m[1, 2]
        q[1] = m[0, 2] - m[2, 0]
        q[2] = m[1, 0] - m[0, 1]
        q[3] = 1 + d[3]
    else:
        i, j, l = k, (k + 1) % 3, (k + 2) % 3
        q[i] = 1 - d[3] + 2 * m[i, i]
        q[j] = m[j, i] + m[i, j]
        q[l] = m[l, i] + m[i, l]
        q[3] = m[l, j] - m[j, l]
    q /= np.linalg.norm(q)
    return torch.tensor(q, dtype=R.dtype)


def signed_log1p(x):
    return torch.sign(x) * torch.log1p(torch.abs(x))


def signed_expm1(x):
    return torch.sign(x) * torch.expm1(torch.abs(x))


def poses_from_params(p: torch.Tensor) -> torch.Tensor:
    """_get_poses base_opt_group.py:260-265: rows [q_xyzw(4), signed-log T(3), ...] -> [n,4,4]."""
    q = p[:, :4]
    q = q / q.norm(dim=-1, keepdim=True)
    R = unitquat_to_rotmat(q)
    T = signed_expm1(p[:, 4:7])
    out = torch.zeros(p.shape[0], 4, 4, dtype=p.dtype)
    out[:, :3, :3] = R
    out[:, :3, 3] = T
    out[:, 3, 3] = 1
    return out


# ============================================================================= evo restatements
def se3_inv(T: np.ndarray) -> np.ndarray:
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def align_origin_and_rpe_rot(traj_est: np.ndarray, traj_ref: np.ndarray):
    """evo PoseTrajectory3D.align_origin + main_rpe.rpe(rotation_angle_deg, delta=1 frame, all_pairs).
    traj_* [n,4,4] c2w.  Returns (P, rpe_rot_rmse_deg); P = ref_0 est_0^-1."""
    P = traj_ref[0] @ se3_inv(traj_est[0])
    ang = []
    for i in range(len(traj_est) - 1):
        q_rel = se3_inv(traj_ref[i]) @ traj_ref[i + 1]
        p_rel = se3_inv(traj_est[i]) @ traj_est[i + 1]
        E = se3_inv(q_rel) @ p_rel
        c = np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
        ang.append(np.degrees(np.arccos(c)))
    ang = np.asarray(ang)
    return P, float(np.sqrt(np.mean(ang ** 2))) if len(ang) else 0.0


# ============================================================================= per-window post-processing
def softplus(x):
    return F.softplus(x)


def postprocess_window(batch_images: torch.Tensor, sky_eps=0.1, far_value=1.99, has_conf=True):
    """infer_geo4d.py:447-500 for modality pc_ray_cross_depth.  batch_images [1, 11, t, h, w].
    Returns dict(pts3d [t,h,w,3], conf [t,h,w,1], inverse_depthmap [t,h,w,1], traj [t,4,4], valid [t,h,w,1])."""
    bs = batch_images[0]
    raymap, crossmap = bs[4:7], bs[7:10]
    traj = raymap_to_camera_matrix(raymap[None], crossmap[None])
    invd = (bs[10:11].permute(1, 2, 3, 0) + 1.0) / 2.0
    x = bs[0:3].permute(1, 2, 3, 0)
    conf = softplus(bs[3:4]).permute(1, 2, 3, 0) if has_conf else torch.ones_like(invd)
    lo, hi = 1.05 - sky_eps, 1.05 + sky_eps
    sky = ((x > lo) & (x < hi)).all(-1, keepdim=True)
    far = (x.abs() > far_value).any(-1, keepdim=True)
    invalid = sky | far
    inv_conf = 1.0 / conf
    inv_conf[invalid] = 0.0
    pts = x.clone()
    pts[..., 0] = pts[..., 0] / 2.0
    pts[..., 1] = pts[..., 1] / 2.0
    pts[..., 2] = (pts[..., 2] + 1) / 2
    return {"pts3d": pts, "conf": inv_conf, "inverse_depthmap": invd, "traj": traj, "valid": ~invalid}


def raymap_to_camera_matrix(raydir: torch.Tensor, raymoment: torch.Tensor) -> torch.Tensor:
    """cameras_from_plucker + rays_to_cameras + the R/T juggling of infer_geo4d.py:657-674 -> c2w [t,4,4]
    = [[R, c],[0,1]].  raydir/raymoment [1, 3, t, h, w]."""
    d = raydir[0].permute(1, 2, 3, 0)
    m = raymoment[0].permute(1, 2, 3, 0)
    T, H, W, _ = d.shape
    if H > W:
        c = (H - W) // 2
        d, m = d[:, c:-c], m[:, c:-c]
    elif W > H:
        c = (W - H) // 2
        d, m = d[:, :, c:-c], m[:, :, c:-c]
    d = d / torch.norm(d, dim=-1, keepdim=True)
    d = d.reshape(T, -1, 3)
    m = m.reshape(T, -1, 3)
    dn = F.normalize(d, dim=-1)
    p = torch.cross(dn, m, dim=-1)  # Rays.to_point_direction (rays.py:141-147)
    r = F.normalize(dn, dim=-1)
    eye = torch.eye(3)[None, None]
    I_min_cov = eye - r[..., None] * r[..., None, :]
    sum_proj = I_min_cov.matmul(p[..., None]).sum(dim=-3)
    centers = torch.linalg.lstsq(I_min_cov.sum(dim=-3), sum_proj).solution[..., 0]
    out = torch.eye(4).repeat(T, 1, 1)
    A = d[0]
    for t in range(T):
        Hm = d[t].T @ A
        U, _, Vh = torch.linalg.svd(Hm, full_matrices=True)
        s = torch.linalg.det(U @ Vh)
        R = U @ torch.diag(torch.tensor([1.0, 1.0, float(torch.sign(s))])) @ Vh
        out[t, :3, :3] = R
        out[t, :3, 3] = centers[t]
    return out


# ============================================================================= init helpers
def image_plane_uv(width, height):
    ar = width / height
    sx = ar / (1 + ar ** 2) ** 0.5
    sy = 1 / (1 + ar ** 2) ** 0.5
    u = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, width)
    v = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, height)
    u, v = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([u, v], dim=-1)


def solve_optimal_shift_focal(uv: np.ndarray, xyz: np.ndarray):
    """utils/geometry.py:232-270 with ransac_iters=None (scipy LM on the z-shift)."""
    from scipy.optimize import least_squares
    uv, xy, z = uv.reshape(-1, 2), xyz[..., :2].reshape(-1, 2), xyz[..., 2].reshape(-1)

    def fn(shift):
        xy_proj = xy / (z + shift)[:, None]
        f = (xy_proj * uv).sum() / np.square(xy_proj).sum()
        return (f * xy_proj - uv).ravel()

    sol = least_squares(fn, x0=0, ftol=1e-3, method="lm")
    shift = sol["x"].squeeze().astype(np.float32)
    xy_proj = xy / (z + shift)[:, None]
    focal = (xy_proj * uv).sum() / (xy_proj * xy_proj).sum()
    return shift, focal


def focal_per_group(ref_pointmap: torch.Tensor, ref_conf: torch.Tensor) -> List[float]:
    """align_group_prefix fast_focal block init_im_poses.py:244-271 (try-branch)."""
    B, H, W, _ = ref_pointmap.shape
    mask = ref_conf > 0.5
    pm = ref_pointmap.clone()
    pm[..., 2] = pm[..., 2] - pm[..., 2].min() + 1
    uv = image_plane_uv(W, H).numpy()
    diag = (H ** 2 + W ** 2) ** 0.5
    foc = []
    for i in range(B):
        mk = mask[i].numpy()
        _, f = solve_optimal_shift_focal(uv[mk], pm[i].numpy()[mk])
        foc.append(float(f))
    foc = torch.tensor(foc, dtype=torch.float32)
    fov_x = 2 * torch.atan(W / diag / foc)
    fov_y = 2 * torch.atan(H / diag / foc)
    fx = 0.5 / torch.tan(fov_x / 2)
    fy = 0.5 / torch.tan(fov_y / 2)
    focal_group = ((fx * W) + (fy * H)) / 2
    mean_f = focal_group[focal_group > 30].mean()
    rel = torch.abs(focal_group - mean_f) / mean_f
    focal_group[rel > 0.6] = mean_f
    return focal_group.numpy().tolist()


def fast_pnp(pts3d: torch.Tensor, focal, msk: torch.Tensor, niter_PnP=10):
    """init_im_poses.py:824-865 (cv2.solvePnPRansac, SQPNP, 3 tentative focals)."""
    import cv2
    if msk.sum() < 4:
        return None
    pts, mk = pts3d.numpy(), msk.numpy()
    H, W, _ = pts.shape
    pixels = np.mgrid[:W, :H].T.astype(np.float32)
    S = max(W, H)
    if focal is None:
        tentative = np.geomspace(S / 2, S * 3, 63)
    else:
        tentative = [focal] + list(np.geomspace(-0.03 * S + focal, 0.03 * S + focal, 2))
    pp = (W / 2, H / 2)
    best = (0,)
    for f in tentative:
        K = np.float32([(f, 0, pp[0]), (0, f, pp[1]), (0, 0, 1)])
        ok, R, T, inl = cv2.solvePnPRansac(pts[mk], pixels[mk], K, None, iterationsCount=niter_PnP,
                                           reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP)
        if not ok:
            continue
        if len(inl) > best[0]:
            best = (len(inl), R, T, f)
    if not best[0]:
        return None
    _, R, T, bf = best
    R = cv2.Rodrigues(R)[0]
    w2c = torch.eye(4)
    w2c[:3, :3] = torch.from_numpy(R)
    w2c[:3, 3] = torch.from_numpy(T).ravel()
    return bf, torch.linalg.inv(w2c)


def _umeyama_pts(pred, pts, conf):
    R, T, s = rigid_points_registration(pred.reshape(-1, 3), pts.reshape(-1, 3), weights=conf.ravel(),
                                        compute_scaling=True)
    return s, R, T


def _srt(s, R, T):
    trf = torch.eye(4)
    trf[:3, :3] = R * s
    trf[:3, 3] = T.ravel()
    return trf


def _geotrf(trf, pts):
    return pts @ trf[:3, :3].T + trf[:3, 3]


# ============================================================================= LAD fit + metric
def lad_adam(x: torch.Tensor, y: torch.Tensor, s_init: float, lr: float, max_iters: int, tol=1e-6):
    """absolute_value_scaling2 depth_eval.py:112-145."""
    s = torch.tensor([s_init], requires_grad=True, dtype=x.dtype)
    t = torch.tensor([0.0], requires_grad=True, dtype=x.dtype)
    opt = torch.optim.Adam([s, t], lr=lr)
    prev = None
    with torch.enable_grad():
        for _ in range(max_iters):
            opt.zero_grad()
            loss = torch.sum(torch.abs(s * x + t - y))
            loss.backward()
            opt.step()
            if prev is not None and abs(prev - loss.item()) < tol:
                break
            prev = loss.item()
    return s.detach().item(), t.detach().item()


def depth_eval_lad2(pred, gt, custom_mask, lr, max_iters):
    """depth_evaluation(..., max_depth=None, align_with_lad2=True, custom_mask=..., return_st=True)
    reduced to what _set_st_depth consumes: {'s','t','d1'}."""
    mask = gt > 0
    p, g = pred[mask], gt[mask]
    s_init = (torch.median(g) / torch.median(p)).item()
    s, t = lad_adam(p, g, s_init, lr, max_iters)
    pa = (s * p + t)[custom_mask[mask]]
    ga = g[custom_mask[mask]]
    if pa.numel() == 0:
        return {"s": s, "t": t, "d1": 0.0}
    pa = torch.clamp(pa, min=1e-5)
    ratio = torch.maximum(pa / ga, ga / pa)
    return {"s": s, "t": t, "d1": torch.mean((ratio < 1.25).float()).item()}


# ============================================================================= the optimiser
class GroupAligner:
    """LightPointCloudGroupOptimizer, group path, shared focal, conf='id', conf_optimize=True."""

    def __init__(self, groups: Sequence[Sequence[int]], pred_list: List[Dict[str, torch.Tensor]], *,
                 temporal_smoothing_weight=0.015, translation_weight=1.0, depth_traj_start_iter=150,
                 base_scale=0.5, focal_break=20.0, lad_max_iters=5000, verbose=False):
        self.groups = [list(g) for g in groups]
        self.G = len(groups)
        self.gs = len(groups[0])
        self.N = max(max(g) for g in groups) + 1
        self.pred = [p["pts3d"].float() for p in pred_list]
        self.conf = [p["conf"].squeeze(-1).float() for p in pred_list]
        self.invd = [p["inverse_depthmap"].float() for p in pred_list]
        self.traj = [p["traj"].float() for p in pred_list]
        self.H, self.W = self.pred[0].shape[1:3]
        self.HW = self.H * self.W
        self.tsw, self.tw = temporal_smoothing_weight, translation_weight
        self.start_b = depth_traj_start_iter
        self.base_scale, self.focal_break = base_scale, focal_break
        self.lad_max_iters = lad_max_iters
        self.verbose = verbose
        # parameters (base_opt_group.py:176-184, optimizer_group.py:58-67); random values are all
        # overwritten before use, zeros keep this deterministic
        z = lambda *s: torch.zeros(*s, requires_grad=True)
        self.s_depth = torch.ones(self.G, 1, requires_grad=True)
        self.t_depth = z(self.G, 1)
        self.pw_poses = z(self.G, 8)
        self.traj_align_poses = z(self.G, 8)
        self.im_depthmaps = z(self.N, self.HW)
        self.im_poses = z(self.N, 7)
        self.im_focals = torch.full((1, 1), focal_break * math.log(max(self.H, self.W)), requires_grad=True)
        self.pp = torch.tensor([self.W / 2, self.H / 2])
        ys, xs = torch.meshgrid(torch.arange(self.H), torch.arange(self.W), indexing="ij")
        self.grid = torch.stack([xs, ys], -1).reshape(self.HW, 2).float()
        self.e_all = torch.tensor([j for g in self.groups for j in g])
        self.weight_all = torch.stack([self.conf[g][i].reshape(-1) for g in range(self.G) for i in range(self.gs)])
        self.pred_all = torch.stack([self.pred[g][i].reshape(-1, 3) for g in range(self.G) for i in range(self.gs)])
        self.invd_all = torch.stack([self.invd[g][i].reshape(-1, 1) for g in range(self.G) for i in range(self.gs)])
        self.traj_all = torch.stack([self.traj[g][i] for g in range(self.G) for i in range(self.gs)])
        self.total_area = self.G * self.gs * self.HW
        self.invalid_depth_group: List[int] = []
        self.valid_traj_groups: List[int] = []
        self.valid_group_idx: List[int] = []

    # ------------------------------------------------------------------ parametrisations
    def get_focal(self):
        return (self.im_focals / self.focal_break).exp()  # [1,1]

    def get_im_poses(self):
        return poses_from_params(self.im_poses)

    def pw_scale(self):
        f = (math.log(self.base_scale) - self.pw_poses[:, -1].mean()).exp()
        return self.pw_poses[:, -1].exp() * f

    def get_pw_poses(self):
        RT = poses_from_params(self.pw_poses).clone()
        sc = self.pw_scale().view(-1, 1, 1)
        RT = torch.cat([RT[:, :3] * sc, RT[:, 3:]], 1)
        return RT

    def get_depthmaps(self):
        return self.im_depthmaps.exp()

    def depth_to_pts3d(self):
        depth = self.get_depthmaps().unsqueeze(-1)
        f = self.get_focal().expand(self.N, 1).unsqueeze(1)
        rel = torch.cat((depth * (self.grid[None] - self.pp[None, None]) / f, depth), dim=-1)
        P = self.get_im_poses()
        return torch.einsum("nij,npj->npi", P[:, :3, :3], rel) + P[:, None, :3, 3]

    def relative_pose_loss(self, RT1, RT2):
        rel = torch.matmul(torch.inverse(RT1), RT2)
        rot = torch.norm(rel[:, :3, :3] - torch.eye(3), dim=(1, 2))
        tr = torch.norm(rel[:, :3, 3], dim=1)
        return rot + tr * self.tw

    # ------------------------------------------------------------------ init (init_from_group)
    @torch.no_grad()
    def _set_pose(self, poses, idx, R, T, scale=None, scale_T=True):
        poses.data[idx, 0:4] = rotmat_to_unitquat(R)
        if scale_T:
            poses.data[idx, 4:7] = signed_log1p(T / (scale if scale is not None else 1))
        else:
            poses.data[idx, 4:7] = signed_log1p(T)
        if scale is not None:
            poses.data[idx, -1] = math.log(float(scale))

    @torch.no_grad()
    def init_from_group(self, niter_PnP=10):
        G, gs, N = self.G, self.gs, self.N
        focal_group = focal_per_group(torch.stack([self.pred[i][0] for i in range(G)]),
                                      torch.stack([self.conf[i][0] for i in range(G)]))
        pts3d, conf_list = [None] * N, [None] * N
        im_poses, im_focals = [None] * N, [None] * N
        done = set()
        for gi, img in enumerate(self.groups[0]):
            if gi == 0:
                im_focals[img] = focal_group[0]
            pts3d[img] = self.pred[0][gi].clone()
            conf_list[img] = self.conf[0][gi].clone()
            msk = self.conf[0][gi] > 0.5
            temp_focal = im_focals[img - 1] if img != 0 else im_focals[img]
            res = fast_pnp(pts3d[img], temp_focal, msk, niter_PnP)
            if res:
                im_focals[img], im_poses[img] = res
            if im_poses[img] is None:
                im_poses[img] = torch.eye(4)
            done.add(img)
        for i, group in enumerate(self.groups):
            if i == 0:
                continue
            assert group[0] in done
            seen = [(gi, img) for gi, img in enumerate(group) if img in done]
            s, R, T = _umeyama_pts(torch.stack([self.pred[i][gi] for gi, _ in seen]),
                                   torch.stack([pts3d[img] for _, img in seen]),
                                   torch.stack([self.conf[i][gi] * conf_list[img] for gi, img in seen]))
            trf = _srt(s, R, T)
            for gi, img in enumerate(group):
                pts3d[img] = _geotrf(trf, self.pred[i][gi])
                conf_list[img] = self.conf[i][gi]
                done.add(img)
                if gi == 0 and im_poses[img] is None:
                    im_poses[img] = _srt(1, R, T)
                msk = self.conf[i][gi] > 0.5
                temp_focal = focal_group[i] if gi == 0 else im_focals[img - 1]
                res = fast_pnp(pts3d[img], temp_focal, msk, niter_PnP)
                if res:
                    im_focals[img], im_poses[img] = res
                if im_poses[img] is None:
                    im_poses[img] = torch.eye(4)
        im_poses = torch.stack(im_poses)
        # init_from_pts3d_group
        for e, group in enumerate(self.groups):
            s, R, T = _umeyama_pts(torch.stack([self.pred[e][i] for i in range(gs)]),
                                   torch.stack([pts3d[g] for g in group]),
                                   torch.stack([self.conf[e][i] * conf_list[g] for i, g in enumerate(group)]))
            self._set_pose(self.pw_poses, e, R, T, scale=s)
        s_factor = (math.log(self.base_scale) - self.pw_poses[:, -1].mean()).exp()
        im_poses[:, :3, 3] *= s_factor
        pts3d = [p * s_factor for p in pts3d]
        sky_distance = 0
        for i in range(N):
            c2w = im_poses[i]
            depth = _geotrf(torch.linalg.inv(c2w), pts3d[i])[..., 2]
            sky = conf_list[i] < 1e-4
            if i == 0:
                depth[sky] = depth.max()
                sky_distance = depth.max()
            else:
                depth[sky] = sky_distance
            self.im_depthmaps.data[i] = depth.reshape(-1).log().nan_to_num(neginf=0)
            self._set_pose(self.im_poses, i, c2w[:3, :3], c2w[:3, 3])
        self.im_focals.data[:] = self.focal_break * math.log(sum(im_focals) / N)
        self.init_im_focals = im_focals
        self.init_im_poses = im_poses

    # ------------------------------------------------------------------ iter-150 sub-alignments
    @torch.no_grad()
    def _set_st_depth(self):
        depth = self.get_depthmaps()
        invdepth = 1.0 / (depth + 1e-6)
        inv_g = invdepth[self.e_all].reshape(self.G, -1).clone()
        rho = self.invd_all.reshape(self.G, -1).clone()
        w = self.weight_all.reshape(self.G, -1).clone()
        cmask = (w > 0.5) & (rho > 0.05)
        invalid = []
        for i in range(self.G):
            best = depth_eval_lad2(rho[i], inv_g[i], cmask[i], 1e-2, self.lad_max_iters)
            self.s_depth.data[i], self.t_depth.data[i] = best["s"], best["t"]
            if best["d1"] < 0.8:
                for lr in (1e-4, 1e-3):
                    r = depth_eval_lad2(rho[i], inv_g[i], cmask[i], lr, min(3000, self.lad_max_iters))
                    if r["d1"] > best["d1"]:
                        best = r
                        self.s_depth.data[i], self.t_depth.data[i] = r["s"], r["t"]
            if best["d1"] < 0.3:
                invalid.append(i)
        return invalid

    @torch.no_grad()
    def _set_traj(self):
        im_pose = self.get_im_poses()
        pw_scale = self.pw_scale()
        valid, valid_idx = [], []
        for i in range(self.G):
            group = self.groups[i]
            traj = self.traj[i].clone()
            traj[:, :3, 3] = traj[:, :3, 3] * pw_scale[i]
            P, rpe_rot = align_origin_and_rpe_rot(traj.double().numpy(), im_pose[group].double().numpy())
            P = torch.from_numpy(P).float()
            self._set_pose(self.traj_align_poses, i, P[:3, :3], P[:3, 3], scale=float(pw_scale[i]), scale_T=False)
            if rpe_rot < 4:
                valid.append(i)
                valid_idx += group
        return valid, valid_idx

    # ------------------------------------------------------------------ loss
    def forward(self, epoch=9999):
        pw_poses = self.get_pw_poses()
        proj = self.depth_to_pts3d()
        new_pw = pw_poses.unsqueeze(1).repeat(1, self.gs, 1, 1).reshape(-1, 4, 4)
        aligned = torch.einsum("eij,epj->epi", new_pw[:, :3, :3], self.pred_all) + new_pw[:, None, :3, 3]
        self.weight_all[self.weight_all > 10] = 10
        li = ((proj[self.e_all] - aligned).norm(dim=-1) * self.weight_all).sum() / self.total_area
        depth_loss = 0
        loss_traj = 0
        if epoch >= self.start_b:
            if epoch == self.start_b:
                self.invalid_depth_group = self._set_st_depth()
            depth = self.get_depthmaps()
            inv_pred = (1 / (depth + 1e-6)).unsqueeze(-1)
            s = self.s_depth.unsqueeze(1).repeat(1, self.gs, 1).reshape(-1, 1, 1)
            t = self.t_depth.unsqueeze(1).repeat(1, self.gs, 1).reshape(-1, 1, 1)
            weight = torch.ones_like(inv_pred[self.e_all])
            weight[~(self.invd_all > 0.05)] = 0
            if len(self.invalid_depth_group) > 0:
                weight = weight.reshape(self.G, self.gs, -1, 1)
                weight[self.invalid_depth_group] = 0
                weight = weight.reshape(self.G * self.gs, -1, 1)
            scaled = self.invd_all * s + t
            depth_loss = ((inv_pred[self.e_all] - scaled).norm(dim=-1) * weight[..., 0]).sum() / self.total_area
            depth_loss = depth_loss * 2
            if epoch == self.start_b:
                self.valid_traj_groups, self.valid_group_idx = self._set_traj()
            if len(self.valid_traj_groups) > 0:
                scale = self.traj_align_poses[:, -1].exp()[self.valid_traj_groups]
                RT = poses_from_params(self.traj_align_poses)[self.valid_traj_groups]
                st = self.traj_all.reshape(self.G, self.gs, 4, 4)[self.valid_traj_groups]
                xyz = st[:, :, :3, [3]] * scale.reshape(-1, 1, 1, 1)
                homo = torch.cat([torch.cat([st[:, :, :3, :3], xyz], -1),
                                  torch.tensor([0., 0, 0, 1]).reshape(1, 1, 1, 4).repeat(st.shape[0], self.gs, 1, 1)], -2)
                homo = torch.bmm(RT.reshape(-1, 1, 4, 4).repeat(1, self.gs, 1, 1).reshape(-1, 4, 4),
                                 homo.reshape(-1, 4, 4))
                loss_traj = self.relative_pose_loss(homo, self.get_im_poses()[self.valid_group_idx]).sum()
        if self.tsw > 0:
            P = self.get_im_poses()
            smooth = self.relative_pose_loss(P[:-1], P[1:]).sum()
        else:
            smooth = 0
        return (li + depth_loss) + loss_traj * 0.005 + self.tsw * smooth

    # ------------------------------------------------------------------ optimisation loop
    def compute_global_alignment(self, niter=500, lr=0.03, lr_min=1e-3, schedule="linear", niter_PnP=10):
        self.init_from_group(niter_PnP)
        params = [self.s_depth, self.t_depth, self.pw_poses, self.traj_align_poses, self.im_depthmaps,
                  self.im_poses, self.im_focals]
        opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.9))
        loss = float("inf")
        for it in range(niter):
            t = it / niter
            cur = lr + (lr_min - lr) * t if schedule == "linear" else \
                lr_min + (lr - lr_min) * (1 + np.cos(t * np.pi)) / 2
            for g in opt.param_groups:
                g["lr"] = cur
            opt.zero_grad()
            L = self.forward(epoch=it)
            L.backward()
            opt.step()
            loss = float(L.detach())
        return loss

    # ------------------------------------------------------------------ outputs
    @torch.no_grad()
    def results(self):
        return {"depth": self.get_depthmaps().reshape(self.N, self.H, self.W).clone(),
                "poses": self.get_im_poses().clone(), "focal": float(self.get_focal()),
                "pw_poses": self.get_pw_poses().clone(), "s_depth": self.s_depth.detach().clone(),
                "t_depth": self.t_depth.detach().clone()}


# ============================================================================= synthetic scene for tests
def synthetic_scene(T=24, H=32, W=48, stride=8, window=16, seed=0, noise=0.005):
    """Smooth random depth + smooth camera path -> per-window predictions in the window-local frame
    (first frame of the window = identity), as Geo4D would produce them."""
    g = torch.Generator().manual_seed(seed)
    f = 0.9 * W
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    base = 2.0 + 1.5 * torch.sin(xs / W * 3.1) * torch.cos(ys / H * 2.3)
    poses = []
    for t in range(T):
        a = 0.01 * t
        R = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
        P = torch.eye(4)
        P[:3, :3] = R
        P[:3, 3] = torch.tensor([0.02 * t, 0.005 * math.sin(0.3 * t), 0.01 * t])
        poses.append(P)
    poses = torch.stack(poses)
    depth = torch.stack([base + 0.05 * torch.sin(0.2 * t + xs / 7.0) for t in range(T)])
    cam = torch.stack([(xs - W / 2) / f, (ys - H / 2) / f, torch.ones_like(xs)], -1)[None] * depth[..., None]
    world = torch.einsum("tij,thwj->thwi", poses[:, :3, :3], cam) + poses[:, None, None, :3, 3]
    starts = list(range(0, T - window + 1, stride))
    if T - window not in starts:
        starts.append(T - window)
    groups, preds = [], []
    for k, s0 in enumerate(starts):
        idx = list(range(s0, s0 + window))
        inv0 = torch.linalg.inv(poses[s0])
        sc = 0.6 + 0.1 * k
        local = (torch.einsum("ij,thwj->thwi", inv0[:3, :3], world[idx]) + inv0[:3, 3]) * sc
        local = local + noise * torch.randn(local.shape, generator=g)
        conf = 0.6 + 2.0 * torch.rand(window, H, W, 1, generator=g)
        conf[:, :2, :3] = 0.0  # a few invalid ("sky") pixels
        traj = torch.stack([inv0 @ poses[i] for i in idx])
        traj[:, :3, 3] *= sc
        d_local = depth[idx] * sc
        invd = (1.0 / d_local).unsqueeze(-1)
        invd = invd / invd.max() * 0.9 + 0.05
        groups.append(idx)
        preds.append({"pts3d": local, "conf": conf, "inverse_depthmap": invd, "traj": traj})
    return groups, preds, {"poses": poses, "depth": depth, "focal": f}
