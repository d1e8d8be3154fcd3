"""Sliding-window global alignment behind the reference's `LightPointCloudGroupOptimizer` seam
(dust3r/cloud_opt/optimizer_group.py + base_opt_group.py + init_im_poses.py, group path).

Same constructor arguments, `compute_global_alignment(init='group', niter, schedule, lr)` and getters
(`get_depthmaps`, `get_im_poses`, `get_focals`, `get_intrinsics`, `get_pts3d`, `get_tum_poses`, `get_masks`,
`save_*`) as the reference.  What runs where:

* the optimisation loop -- per pixel: objective, gradient w.r.t. every log-depth value, Adam update of the N x HW
  log-depth maps, reductions of the gradient w.r.t. poses / focal / window sim(3) / depth scale-shift; then the
  O(N + G) small-parameter part (chain rule to the quaternion / signed-log / log-scale parametrisations,
  temporal-smoothing and trajectory-prior terms, Adam) -- runs as ONE persistent cooperative kernel per phase
  (geo4d_align_loop, engine="loop", the default): two grid barriers per iteration, deterministic reductions.  Under
  torch.distributed the images are split over the ranks and the kernel exchanges the reduced gradients by stores
  into peer-mapped memory (sharding.PeerExchange), every rank ending with bit-identical parameters.
  engine="steps" keeps the round-1 form (geo4d_align_iter + geo4d_align_small_step per iteration, CUDA-graph
  replayed); GEO4D_ALIGN_AUTOGRAD=1 switches the small part to torch autograd + torch.optim.Adam (used by the tests
  to cross-check the hand-derived gradients);
* the weighted Umeyama registrations, the LAD scale/shift fit (same Adam iteration as the reference, one cooperative
  launch per window, sync-free) and the delta<1.25 gate are reduction kernels (csrc/align.cu);
* the initialisation solvers the reference runs on the CPU -- the shift/focal least-squares fit (scipy LM) and
  the per-frame RANSAC-PnP (cv2, SQPnP) -- reduce their per-pixel sums on the GPU (geo4d_shift_focal_sums,
  geo4d_pnp_moments) and solve only 1-D / 9x9 problems on the host (init_solvers.py, csrc/sqpnp_host.cu): every window
  in its own frame (windows independent -> batched and shardable), RANSAC hypotheses scored in one launch; set
  GEO4D_INIT_SOLVERS=host to call scipy / cv2 exactly like the reference (1.5-2 s per frame at 320x512).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._cabi import require_device


# --------------------------------------------------------------------------- small helpers (host / torch)
def signed_log1p(x):
    return torch.sign(x) * torch.log1p(torch.abs(x))


def signed_expm1(x):
    return torch.sign(x) * torch.expm1(torch.abs(x))


def unitquat_to_rotmat(q):
    """xyzw (roma convention, base_opt_group.py:264,279)."""
    x, y, z, w = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)
    return R.reshape(*q.shape[:-1], 3, 3)


def rotmat_to_unitquat_np(m: np.ndarray) -> np.ndarray:
    d = np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 0] + m[1, 1] + m[2, 2]])
    k = int(np.argmax(d))
    q = np.empty(4)
    if k == 3:
        q[:] = (m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], 1 + d[3])
    else:
        i, j, l = k, (k + 1) % 3, (k + 2) % 3
        q[i] = 1 - d[3] + 2 * m[i, i]
        q[j] = m[j, i] + m[i, j]
        q[l] = m[l, i] + m[i, l]
        q[3] = m[l, j] - m[j, l]
    return q / np.linalg.norm(q)


def poses_from_params(p):
    """_get_poses (base_opt_group.py:260-265) -> [n, 4, 4] cam-to-world."""
    q = p[:, :4]
    q = q / q.norm(dim=-1, keepdim=True)
    R = unitquat_to_rotmat(q)
    T = signed_expm1(p[:, 4:7])
    top = torch.cat([R, T.unsqueeze(-1)], -1)
    bot = torch.zeros(p.shape[0], 1, 4, device=p.device, dtype=p.dtype)
    bot[:, 0, 3] = 1
    return torch.cat([top, bot], 1)


def umeyama_from_moments(m0: np.ndarray, m1: np.ndarray):
    """(sum w, sum w x, sum w y) and (sum w|x^|^2, sum w y^ x^T) -> (s, R, T) with y ~ s R x + T."""
    sw = m0[0]
    if not (np.isfinite(m0).all() and np.isfinite(m1).all() and sw > 0 and m1[0] > 0):
        return float("nan"), np.eye(3), np.zeros(3)   # no overlap / all-zero weights: the caller applies its guard
    xm, ym = m0[1:4] / sw, m0[4:7] / sw
    M = m1[1:10].reshape(3, 3)
    U, D, Vt = np.linalg.svd(M)
    det = np.linalg.det(U) * np.linalg.det(Vt)
    Dm = np.array([1.0, 1.0, det])
    R = U @ np.diag(Dm) @ Vt
    s = float((D * Dm).sum() / m1[0])
    T = ym - s * (R @ xm)
    return s, R, T


def _se3_inv(T):
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def c2w_to_tumpose(c2w: np.ndarray) -> np.ndarray:
    """base_opt_group.py:28-41: xyz + quaternion wxyz."""
    from scipy.spatial.transform import Rotation
    qx, qy, qz, qw = Rotation.from_matrix(c2w[:3, :3]).as_quat()
    return np.concatenate([c2w[:3, -1], [qw, qx, qy, qz]])


class LightPointCloudGroupOptimizer(nn.Module):
    POSE_DIM = 7

    def __init__(self, view_list, pred_list, dist="l1", conf="log", min_conf_thr=3, thr_for_init_conf=False,
                 base_scale=0.5, allow_pw_adaptors=False, pw_break=20, rand_pose=torch.randn, empty_cache=False,
                 verbose=True, opt_raydir=False, optimize_pp=False, focal_break=20, shared_focal=False,
                 flow_loss_fn="smooth_l1", flow_loss_weight=0.0, depth_regularize_weight=0.0, num_total_iter=300,
                 temporal_smoothing_weight=0, translation_weight=0.1, flow_loss_start_epoch=0.15, flow_loss_thre=50,
                 sintel_ckpt=False, use_self_mask=False, pxl_thre=50, sam2_mask_refine=True, motion_mask_thre=0.35,
                 conf_optimize=False, depth_traj_start_iter=150, use_cuda_graph=True, lad_max_iters=5000,
                 shard_sub_alignments=True, engine=None, shard_alignment=True):
        super().__init__()
        if dist != "l1" or conf not in ("id", "none") or not shared_focal or not conf_optimize or opt_raydir \
                or optimize_pp or allow_pw_adaptors or flow_loss_weight != 0.0 or depth_regularize_weight != 0.0:
            raise NotImplementedError(
                "geo4d_b200 implements the configuration the Geo4D scripts use (infer_geo4d.py:32-40): dist='l1', "
                "conf='id', conf_optimize=True, shared_focal=True, no flow / raydir / pp / adaptor terms")
        self.groups = [[int(v["idx"][-1]) for v in views] for views in view_list]
        self.n_groups = len(self.groups)
        self.group_size = len(self.groups[0])
        self.n_imgs = max(max(g) for g in self.groups) + 1
        p0 = pred_list[0]["pts3d"]
        self.H, self.W = int(p0.shape[1]), int(p0.shape[2])
        self.HW = self.H * self.W
        self.imshapes = [(self.H, self.W)] * self.n_imgs
        self.imshape = (self.H, self.W)
        self.verbose = verbose
        self.base_scale = base_scale
        self.focal_break = focal_break
        self.shared_focal = shared_focal
        self.num_total_iter = num_total_iter
        self.temporal_smoothing_weight = temporal_smoothing_weight
        self.translation_weight = translation_weight
        self.depth_traj_start_iter = depth_traj_start_iter
        self.min_conf_thr = min_conf_thr
        self.thr_for_init_conf = thr_for_init_conf
        self.has_im_poses = True
        self.use_cuda_graph = use_cuda_graph and os.environ.get("GEO4D_ALIGN_EAGER", "0") != "1"
        self.lad_max_iters = lad_max_iters
        self.shard_sub_alignments = shard_sub_alignments  # split per-window LAD fits over torch.distributed ranks
        # "loop": the whole optimisation loop is one persistent cooperative kernel per phase (geo4d_align_loop);
        # "steps": two kernels per iteration (geo4d_align_iter + geo4d_align_small_step), CUDA-graph replayed
        self.engine = engine or os.environ.get("GEO4D_ALIGN_ENGINE", "loop")
        if self.engine not in ("loop", "steps"):
            raise ValueError(f"engine={self.engine!r}: 'loop' or 'steps'")
        # under torch.distributed the images of the loop engine are sharded over the ranks (peer-to-peer exchange
        # of the reduced gradients inside the kernel); False = every rank optimises the whole clip (replicated)
        self.shard_alignment = shard_alignment
        self._shard = None
        self._profile = None
        dev = p0.device
        G, gs, N, HW = self.n_groups, self.group_size, self.n_imgs, self.HW
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
        # stacked per-(window, frame) observations (optimizer_group.py:91-104)
        self.register_buffer("_stacked_pred_all", torch.stack([f32(p["pts3d"]).reshape(gs, HW, 3) for p in pred_list])
                             .reshape(G * gs, HW, 3).contiguous())
        self.register_buffer("_weight_all", torch.stack([f32(p["conf"]).reshape(gs, HW) for p in pred_list])
                             .reshape(G * gs, HW).contiguous())
        self.has_invdepth = pred_list[0].get("inverse_depthmap", None) is not None
        self.has_traj = pred_list[0].get("traj", None) is not None
        if self.has_invdepth:
            self.register_buffer("_stacked_depthmap_all", torch.stack(
                [f32(p["inverse_depthmap"]).reshape(gs, HW) for p in pred_list]).reshape(G * gs, HW).contiguous())
        if self.has_traj:
            self.register_buffer("_stacked_traj_all", torch.stack([f32(p["traj"]) for p in pred_list])
                                 .reshape(G * gs, 4, 4).contiguous())
        e_all = [j for g in self.groups for j in g]
        self.register_buffer("_e_all", torch.tensor(e_all, device=dev))
        # images -> incident edges (CSR) for the fused kernel
        inc: List[List[int]] = [[] for _ in range(N)]
        for e, n in enumerate(e_all):
            inc[n].append(e)
        self.max_edges_per_image = max(len(x) for x in inc)
        ptr = np.zeros(N + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(x) for x in inc])
        self.register_buffer("_edge_ptr", torch.tensor(ptr, device=dev))
        self.register_buffer("_edge_idx", torch.tensor([e for x in inc for e in x], dtype=torch.int32, device=dev))
        self.total_area_all = G * gs * HW
        # single-image confidence = max over windows (base_opt_group.py:229-235)
        im_conf = torch.zeros(N, HW, device=dev)
        for e, n in enumerate(e_all):
            im_conf[n] = torch.maximum(im_conf[n], self._weight_all[e])
        self.im_conf = [c.reshape(self.H, self.W) for c in im_conf]
        self.init_conf_maps = [c.clone() for c in self.im_conf]
        # parameters.  The reference draws torch.randn initial values that init_from_group / _set_traj fully
        # overwrite before first use; zeros keep the run deterministic.
        self.im_depthmaps = nn.Parameter(torch.zeros(N, HW, device=dev), requires_grad=False)  # log-depth
        self.im_poses = nn.Parameter(torch.zeros(N, self.POSE_DIM, device=dev))
        self.im_focals = nn.Parameter(torch.full((1, 1), focal_break * math.log(max(self.H, self.W)), device=dev))
        self.pw_poses = nn.Parameter(torch.zeros(G, 1 + self.POSE_DIM, device=dev))
        self.traj_align_poses = nn.Parameter(torch.zeros(G, 1 + self.POSE_DIM, device=dev))
        self.s_depth = nn.Parameter(torch.ones(G, 1, device=dev))
        self.t_depth = nn.Parameter(torch.zeros(G, 1, device=dev))
        with torch.no_grad():
            self.im_poses[:, 3] = 1
            self.pw_poses[:, 3] = 1
            self.traj_align_poses[:, 3] = 1
        self.register_buffer("_pp", torch.tensor([(self.W / 2, self.H / 2)] * N, device=dev))
        self.invalid_depth_group: List[int] = []
        self.valid_traj_group_list: List[int] = []
        self.valid_group_idx: List[int] = []
        self.imgs = None
        if "img" in view_list[0][0]:
            imgs = [None] * N
            for views in view_list:
                for v in views:
                    im = v["img"][0] if v["img"].dim() == 4 else v["img"]
                    imgs[int(v["idx"][-1])] = (im.detach().float().cpu().permute(1, 2, 0).numpy() * 0.5 + 0.5)
            self.imgs = imgs

    @property
    def device(self):
        return self.im_poses.device

    # ------------------------------------------------------------------ getters (reference names)
    def get_focals(self):
        lf = torch.stack([self.im_focals[0]] * self.n_imgs, dim=0)
        return (lf / self.focal_break).exp()

    def get_principal_points(self):
        return self._pp

    def get_intrinsics(self):
        K = torch.zeros((self.n_imgs, 3, 3), device=self.device)
        f = self.get_focals().flatten()
        K[:, 0, 0] = K[:, 1, 1] = f
        K[:, :2, 2] = self.get_principal_points()
        K[:, 2, 2] = 1
        return K

    def get_im_poses(self):
        return poses_from_params(self.im_poses)

    def get_pw_norm_scale_factor(self):
        return (math.log(self.base_scale) - self.pw_poses[:, -1].mean()).exp()

    def get_pw_scale(self):
        return self.pw_poses[:, -1].exp() * self.get_pw_norm_scale_factor()

    def get_pw_poses(self):
        RT = poses_from_params(self.pw_poses)
        sc = self.get_pw_scale().view(-1, 1, 1)
        return torch.cat([RT[:, :3] * sc, RT[:, 3:]], 1)

    def get_depthmaps(self, raw=False):
        res = self.im_depthmaps.detach().exp()
        if not raw:
            res = [dm.view(self.H, self.W) for dm in res]
        return res

    def get_pts3d(self, raw=False, **kw):
        with torch.no_grad():
            depth = self.im_depthmaps.exp().unsqueeze(-1)
            ys, xs = torch.meshgrid(torch.arange(self.H, device=self.device),
                                    torch.arange(self.W, device=self.device), indexing="ij")
            grid = torch.stack([xs, ys], -1).reshape(1, self.HW, 2).float()
            f = self.get_focals().unsqueeze(1)
            rel = torch.cat((depth * (grid - self._pp.unsqueeze(1)) / f, depth), dim=-1)
            P = self.get_im_poses()
            res = torch.einsum("nij,npj->npi", P[:, :3, :3], rel) + P[:, None, :3, 3]
        if not raw:
            res = [r.view(self.H, self.W, 3) for r in res]
        return res

    def get_masks(self):
        confs = self.init_conf_maps if self.thr_for_init_conf else self.im_conf
        return [(c > self.min_conf_thr) for c in confs]

    def get_conf(self, mode=None):
        return list(self.im_conf)

    def get_tum_poses(self):
        poses = self.get_im_poses().detach().cpu().numpy()
        return [np.stack([c2w_to_tumpose(p) for p in poses], 0), np.arange(len(poses)).astype(float)]

    def save_tum_poses(self, path):
        """base_opt_group.py:390-393 -> vo_eval.save_trajectory_tum_format (:465-473): one line per pose
        `timestamp x y z qw qx qy qz` -- the quaternion stays in the wxyz order get_tum_poses returns (the
        reference's viewer reads column 4 as w)."""
        from .metrics import save_trajectory_tum_format
        traj = self.get_tum_poses()
        save_trajectory_tum_format(traj, path)
        return traj[0]

    def save_focals(self, path):
        focals = self.get_focals()
        np.savetxt(path, focals.detach().cpu().numpy(), fmt="%.6f")
        return focals

    def save_intrinsics(self, path):
        K = self.get_intrinsics().detach().cpu().numpy()
        np.savetxt(path, K.reshape(-1, 9), fmt="%.6f")
        return K

    def save_depth_maps(self, path):
        """base_opt_group.py:433-464: frame_%04d.npy (depth), frame_colordepth_%04d.png (inverse depth, 2-98
        percentile window over the sequence, inferno colour map) and colored_depth_maps.gif.  The reference takes
        the colour table from matplotlib (cm.get_cmap('inferno').colors, truncated to 8 bits); matplotlib is not a
        dependency here, so the same published table is taken from OpenCV (COLORMAP_INFERNO, rounded to 8 bits):
        colours may differ by one code value."""
        import cv2
        dms = self.get_depthmaps()
        for i, d in enumerate(dms):
            np.save(f"{path}/frame_{i:04d}.npy", d.detach().cpu().numpy())
        inv = (1.0 / (self.get_depthmaps(raw=True).reshape(-1, self.H, self.W) + 1e-6)).cpu().numpy()
        v_min, v_max = np.percentile(inv, 2), np.percentile(inv, 98)
        idx = np.clip(((inv - v_min) / (v_max - v_min) * 255).astype(np.int64), 0, 255).astype(np.uint8)
        frames = []
        for i, im in enumerate(idx):
            bgr = cv2.applyColorMap(im, cv2.COLORMAP_INFERNO)
            cv2.imwrite(f"{path}/frame_colordepth_{i:04d}.png", bgr)
            frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        try:
            from PIL import Image
            ims = [Image.fromarray(f) for f in frames]
            ims[0].save(f"{path}/colored_depth_maps.gif", save_all=True, append_images=ims[1:], duration=100, loop=0)
        except ImportError:  # pragma: no cover
            pass
        return torch.from_numpy(inv)

    def save_rgb_imgs(self, path):
        """base_opt_group.py:419-425: frame_%04d.png of the input frames (RGB in [0, 1] -> BGR 8 bit)."""
        import cv2
        if self.imgs is None:
            raise ValueError("save_rgb_imgs: the views carried no 'img' entries")
        for i, img in enumerate(self.imgs):
            cv2.imwrite(f"{path}/frame_{i:04d}.png", np.ascontiguousarray(img[..., ::-1]) * 255)
        return self.imgs

    def save_conf_maps(self, path):
        for i, c in enumerate(self.im_conf):
            np.save(f"{path}/conf_{i}.npy", c.detach().cpu().numpy())
        return self.im_conf

    def save_init_conf_maps(self, path):
        for i, c in enumerate(self.init_conf_maps):
            np.save(f"{path}/init_conf_{i}.npy", c.detach().cpu().numpy())
        return self.init_conf_maps

    def preset_focal(self, known_focals, msk=None, requires_grad=False):
        with torch.no_grad():
            self.im_focals[:] = self.focal_break * math.log(float(sum(float(f) for f in known_focals) /
                                                                  len(known_focals)))
        if len(known_focals) == self.n_imgs:
            self.im_focals.requires_grad_(requires_grad)

    # ------------------------------------------------------------------ init (init_from_group)
    def _umeyama(self, x, y, w1, w2):
        """weighted Umeyama y ~ s R x + T on the GPU reduction kernel; x, y [n,3], w1, w2 [n]."""
        n = x.shape[0]
        m0 = ops.umeyama_moments(x, y, w1, w2, n, 0, None)
        means = (m0[1:7] / m0[0]).contiguous()
        m1 = ops.umeyama_moments(x, y, w1, w2, n, 1, means)
        s, R, T = umeyama_from_moments(m0.cpu().numpy(), m1.cpu().numpy())
        # Degenerate registrations (uncorrelated point sets, e.g. the output of a randomly initialised network:
        # the fitted scale decays geometrically along the window chain until the fp32 point maps underflow to 0)
        # would make the reference write log(0) = -inf into pw_poses and carry NaNs from there on; a floor keeps
        # every later stage finite.  It never binds on registrable windows (s ~ 1).
        if not (s > 1e-12) or not math.isfinite(s):
            s = 1e-12
        if not np.all(np.isfinite(R)):
            R = np.eye(3)
        if not np.all(np.isfinite(T)):
            T = np.zeros(3)
        return s, R, T

    @staticmethod
    def _set_pose(poses, idx, R, T, scale=None, scale_T=True):
        with torch.no_grad():
            poses[idx, 0:4] = torch.as_tensor(rotmat_to_unitquat_np(np.asarray(R, dtype=np.float64)),
                                              dtype=poses.dtype, device=poses.device)
            Tt = torch.as_tensor(np.asarray(T, dtype=np.float64), dtype=poses.dtype, device=poses.device)
            poses[idx, 4:7] = signed_log1p(Tt / ((scale or 1) if scale_T else 1))   # base_opt_group.py:280-283
            if scale is not None:
                with np.errstate(divide="ignore", invalid="ignore"):
                    poses[idx, -1] = float(np.log(float(scale)))                       # :287 (numpy: log 0 = -inf)

    @torch.no_grad()
    def _init_from_group(self, niter_PnP=10):
        from . import init_solvers as isv
        G, gs, N, HW, H, W = self.n_groups, self.group_size, self.n_imgs, self.HW, self.H, self.W
        dev = self.device
        pred = self._stacked_pred_all.view(G, gs, HW, 3)
        conf = self._weight_all.view(G, gs, HW)
        host_solvers = os.environ.get("GEO4D_INIT_SOLVERS", "gpu") == "host"  # cv2 / scipy exactly like the reference
        try:
            if host_solvers:
                focal_group = isv.focal_per_group(pred[:, 0].reshape(G, H, W, 3).cpu(), conf[:, 0].reshape(G, H, W).cpu())
            else:
                focal_group = isv.gpu_focal_per_group(ops, pred[:, 0].contiguous(), conf[:, 0].contiguous(), H, W)
            if not all(math.isfinite(f) for f in focal_group):
                raise ValueError("non-finite focal")
        except Exception:
            # same fallback as align_group_prefix (init_im_poses.py:272-278): focal search by PnP on the first
            # frame, shared by all windows
            if self.verbose:
                print("Error in computing focal length")
            tmp_f, tmp_p = [None], [None]
            isv.gpu_fast_pnp_frames(ops, pred[0, 0:1].contiguous(), conf[0, 0:1].contiguous(), H, W,
                                    lambda k, img: None, tmp_f, tmp_p, [0], niter_PnP)
            focal_group = [tmp_f[0] if tmp_f[0] else float(max(H, W))] * G
        pts3d = torch.zeros(N, HW, 3, device=dev)
        conf_list = torch.zeros(N, HW, device=dev)
        im_poses: List[Optional[np.ndarray]] = [None] * N
        im_focals: List[Optional[float]] = [None] * N
        done = set()
        chain_pnp = host_solvers or os.environ.get("GEO4D_INIT_PNP", "window") == "chain"

        def pnp_frames(frame_ids, pts_gpu, conf_gpu, first_focal_of):
            """per-frame pose + focal (fast_pnp, init_im_poses.py:824-865) on world-frame points, frame by frame"""
            if host_solvers:
                pts_cpu = pts_gpu.reshape(-1, H, W, 3).cpu().numpy()
                msk_cpu = (conf_gpu > 0.5).reshape(-1, H, W).cpu().numpy()
                for k, img in enumerate(frame_ids):
                    res = isv.fast_pnp(pts_cpu[k], first_focal_of(k, img), msk_cpu[k], niter_PnP)
                    if res:
                        im_focals[img], im_poses[img] = res
                    if im_poses[img] is None:
                        im_poses[img] = np.eye(4)
            else:
                isv.gpu_fast_pnp_frames(ops, pts_gpu.contiguous(), conf_gpu.contiguous(), H, W, first_focal_of,
                                        im_focals, im_poses, frame_ids, niter_PnP)

        # Stage 1 (default): PnP of every frame in its window's OWN frame.  Windows are independent there, so under
        # torch.distributed rank r solves the windows g with g % world == r and the (focal, pose) records are
        # all-gathered; the sequential window chain below only composes them with each window's registration.
        local = None
        if not chain_pnp:
            world, rank = self._dist_sub()
            from . import sharding
            mine = sharding.windows_for_rank(G, rank, world) if world > 1 else list(range(G))
            per = -(-G // world)
            rec = torch.zeros(per, gs * 18, dtype=torch.float64, device=dev)
            if mine:
                idx = torch.tensor(mine, device=dev)
                f_l, c_l, ok_l = isv.gpu_fast_pnp_windows(ops, pred[idx].contiguous(), conf[idx].contiguous(), H, W,
                                                          [focal_group[g] for g in mine], niter_PnP)
                r_np = np.concatenate([np.nan_to_num(f_l, nan=-1.0)[..., None], ok_l[..., None].astype(np.float64),
                                       c_l.reshape(len(mine), gs, 16)], -1).reshape(len(mine), gs * 18)
                rec[:len(mine)] = torch.from_numpy(r_np).to(dev)
            full = (sharding.gather_group_records(rec, G) if world > 1 else rec[:G]).cpu().numpy().reshape(G, gs, 18)
            local = (full[..., 0], full[..., 1] > 0.5, full[..., 2:].reshape(G, gs, 4, 4))

        def compose(i, group, s_i, R_i, T_i):
            """window-frame PnP results -> world frame through the window's registration y = s R x + T"""
            f_l, ok_l, c_l = local
            for k, img in enumerate(group):
                if ok_l[i, k]:
                    P = np.eye(4)
                    P[:3, :3] = R_i @ c_l[i, k, :3, :3]
                    P[:3, 3] = s_i * (R_i @ c_l[i, k, :3, 3]) + T_i
                    im_focals[img], im_poses[img] = float(f_l[i, k]), P
                if im_poses[img] is None:
                    im_poses[img] = np.eye(4)

        g0 = self.groups[0]
        im_focals[g0[0]] = focal_group[0]
        pts3d[g0] = pred[0]
        conf_list[g0] = conf[0]
        if chain_pnp:
            pnp_frames(g0, pred[0], conf[0], lambda k, img: im_focals[img - 1] if img != 0 else im_focals[img])
        else:
            compose(0, g0, 1.0, np.eye(3), np.zeros(3))
        done.update(g0)
        for i in range(1, G):
            group = self.groups[i]
            assert group[0] in done, "The first image of the following group should be in the previous group"
            seen = [(k, img) for k, img in enumerate(group) if img in done]
            ks = torch.tensor([k for k, _ in seen], device=dev)
            ims = torch.tensor([img for _, img in seen], device=dev)
            s, R, T = self._umeyama(pred[i][ks].reshape(-1, 3).contiguous(), pts3d[ims].reshape(-1, 3).contiguous(),
                                    conf[i][ks].reshape(-1).contiguous(), conf_list[ims].reshape(-1).contiguous())
            reg = np.concatenate([s * R, np.asarray(T).reshape(3, 1)], 1).reshape(1, 12)
            new_pts = ops.transform_points(pred[i].reshape(1, gs * HW, 3), torch.from_numpy(reg)).reshape(gs, HW, 3)
            pts3d[group] = new_pts
            conf_list[group] = conf[i]
            if im_poses[group[0]] is None:
                P = np.eye(4)
                P[:3, :3], P[:3, 3] = R, T
                im_poses[group[0]] = P
            if chain_pnp:
                pnp_frames(group, new_pts, conf[i],
                           lambda k, img, fg=focal_group[i]: fg if k == 0 else im_focals[img - 1])
            else:
                compose(i, group, s, R, T)
            done.update(group)
        im_poses_np = np.stack(im_poses)
        # init_from_pts3d_group (init_im_poses.py:569-633)
        for e, group in enumerate(self.groups):
            gi = torch.tensor(group, device=dev)
            s, R, T = self._umeyama(pred[e].reshape(-1, 3).contiguous(), pts3d[gi].reshape(-1, 3).contiguous(),
                                    conf[e].reshape(-1).contiguous(), conf_list[gi].reshape(-1).contiguous())
            self._set_pose(self.pw_poses, e, R, T, scale=s)
        s_factor = float(self.get_pw_norm_scale_factor())
        im_poses_np[:, :3, 3] *= s_factor
        pts3d *= s_factor
        w2c = torch.from_numpy(np.linalg.inv(im_poses_np)[:, :3, :].reshape(N, 12))
        depth = ops.transform_points(pts3d, w2c, depth_only=True)
        sky = conf_list < 1e-4
        sky_distance = depth[0].max()
        depth = torch.where(sky, sky_distance.expand_as(depth), depth)
        self.im_depthmaps.data.copy_(depth.log().nan_to_num(neginf=0))
        for i in range(N):
            self._set_pose(self.im_poses, i, im_poses_np[i][:3, :3], im_poses_np[i][:3, 3])
        # (the reference assumes every frame got a focal from PnP; a frame whose PnP failed inherits its neighbour's)
        for i in range(N):
            if im_focals[i] is None or not math.isfinite(im_focals[i]) or im_focals[i] <= 0:
                im_focals[i] = im_focals[i - 1] if i > 0 and im_focals[i - 1] else float(max(H, W))
        self.im_focals.data[:] = self.focal_break * math.log(sum(im_focals) / N)
        self._init_im_focals = im_focals

    # ------------------------------------------------------------------ iteration-150 sub-alignments
    @torch.no_grad()
    def _set_st_depth(self):
        """optimizer_group.py:333-372 with the LAD fit of depth_eval.py:112-145 on the GPU.  The per-window fits are
        independent, so under torch.distributed (one process per GPU, replicated alignment) each rank fits the
        windows g with g % world == rank and the (s, t, delta<1.25) triples are all-gathered: every rank ends up
        with bit-identical values and the stage costs one window's fit instead of G."""
        G, n = self.n_groups, self.group_size * self.HW
        dev = self.device
        y_all = (1.0 / (self.im_depthmaps.exp() + 1e-6))[self._e_all].reshape(G, n)
        x_all = self._stacked_depthmap_all.reshape(G, n)
        w_all = self._weight_all.reshape(G, n)

        def solve(sel):
            k = len(sel)
            idx = torch.tensor(sel, device=dev, dtype=torch.long)
            x, y, w = x_all[idx].contiguous(), y_all[idx].contiguous(), w_all[idx].contiguous()
            s_init = torch.median(y, dim=1).values / torch.median(x, dim=1).values

            def fit(lr, iters):
                state = torch.zeros(k, 9, device=dev)
                state[:, 0] = s_init
                acc = torch.zeros(k * 4, device=dev, dtype=torch.float64)
                if self.use_cuda_graph and os.environ.get("GEO4D_LAD_STEPWISE", "0") != "1":
                    # one cooperative launch for the whole fit (grid barrier per iteration, data held in smem)
                    ops.lad_fit(x, y, n, k, state, acc, lr, iters)
                    d = ops.delta125(x, y, w, n, k, state, 9).cpu().numpy()
                    d1 = np.where(d[:, 1] > 0, d[:, 0] / np.maximum(d[:, 1], 1), 0.0)
                    return state[:, :2].clone(), d1
                chunk = 250
                graph = None
                done_iters = 0
                if self.use_cuda_graph and iters >= 2 * chunk:
                    ops.lad_step(x, y, n, k, state, acc, lr)  # warm-up (real iteration 0)
                    done_iters = 1
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    n0 = ops.raw_launch_count()
                    with torch.cuda.stream(side), ops.capture_graph(graph):
                        for _ in range(chunk):
                            ops.lad_step(x, y, n, k, state, acc, lr)
                    torch.cuda.current_stream().wait_stream(side)
                    nk = ops.raw_launch_count() - n0
                    ops.note_replay(nk, -1)
                    while done_iters + chunk <= iters:
                        graph.replay()
                        ops.note_replay(nk)
                        done_iters += chunk
                        # the kernel freezes a window once |delta loss| < tol (depth_eval.py:139-141 breaks there);
                        # one flag read per 250 iterations stops launching no-op kernels when every window is done
                        if bool((state[:, 8] != 0).all()):
                            done_iters = iters
                            break
                for _ in range(iters - done_iters):
                    ops.lad_step(x, y, n, k, state, acc, lr)
                d = ops.delta125(x, y, w, n, k, state, 9).cpu().numpy()
                d1 = np.where(d[:, 1] > 0, d[:, 0] / np.maximum(d[:, 1], 1), 0.0)
                return state[:, :2].clone(), d1

            best_st, best_d1 = fit(1e-2, self.lad_max_iters)
            retry = best_d1 < 0.8
            if retry.any():
                for lr in (1e-4, 1e-3):
                    st, d1 = fit(lr, min(3000, self.lad_max_iters))
                    better = retry & (d1 > best_d1)
                    bt = torch.tensor(better, device=dev)
                    best_st = torch.where(bt[:, None], st, best_st)
                    best_d1 = np.where(better, d1, best_d1)
            return best_st, best_d1

        def solve_all(sel):
            """the one-launch fit keeps a window's samples in shared memory only while (samples / SMs-per-window)
            fits; several big windows on one GPU are therefore fitted one after the other, each on every SM"""
            if len(sel) <= 1 or n * 8 * len(sel) <= 148 * 190 * 1024 or not self.use_cuda_graph:
                return solve(sel)
            parts = [solve([g]) for g in sel]
            return torch.cat([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts])

        world, rank = self._dist_sub()
        if world > 1 and G > 1:
            from . import sharding
            per = -(-G // world)
            mine = sharding.windows_for_rank(G, rank, world)
            rec = torch.zeros(per, 3, device=dev, dtype=torch.float64)
            if mine:
                st, d1 = solve_all(mine)
                rec[:len(mine), :2] = st.double()
                rec[:len(mine), 2] = torch.as_tensor(d1, device=dev, dtype=torch.float64)
            full = sharding.gather_group_records(rec, G)
            best_st, best_d1 = full[:, :2].float(), full[:, 2].cpu().numpy()
        else:
            best_st, best_d1 = solve_all(list(range(G)))
        self.s_depth.data[:, 0] = best_st[:, 0]
        self.t_depth.data[:, 0] = best_st[:, 1]
        return [int(i) for i in np.nonzero(best_d1 < 0.3)[0]]

    @torch.no_grad()
    def _set_traj(self):
        """optimizer_group.py:242-267; evo align_origin + RPE-rot gate restated."""
        im_pose = self.get_im_poses().double().cpu().numpy()
        pw_scale = self.get_pw_scale().double().cpu().numpy()
        traj_all = self._stacked_traj_all.view(self.n_groups, self.group_size, 4, 4).double().cpu().numpy()
        valid, valid_idx = [], []
        for i in range(self.n_groups):
            group = self.groups[i]
            traj = traj_all[i].copy()
            traj[:, :3, 3] *= pw_scale[i]
            ref = im_pose[group]
            P = ref[0] @ _se3_inv(traj[0])
            ang = []
            for k in range(len(group) - 1):
                q_rel = _se3_inv(ref[k]) @ ref[k + 1]
                p_rel = _se3_inv(traj[k]) @ traj[k + 1]
                E = _se3_inv(q_rel) @ p_rel
                ang.append(np.degrees(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1))))
            rpe_rot = float(np.sqrt(np.mean(np.square(ang)))) if ang else 0.0
            self._set_pose(self.traj_align_poses, i, P[:3, :3], P[:3, 3], scale=float(pw_scale[i]), scale_T=False)
            if rpe_rot < 4:
                valid.append(i)
                valid_idx += group
        return valid, valid_idx

    # ------------------------------------------------------------------ one optimisation step
    @staticmethod
    def _rigid_inverse(RT):
        """inverse of [R T; 0 1] = [R^T, -R^T T; 0 1] (the reference calls torch.inverse on rigid matrices;
        the closed form is CUDA-graph capturable and sync-free)."""
        Rt = RT[:, :3, :3].transpose(1, 2)
        t = -torch.matmul(Rt, RT[:, :3, 3:4])
        return torch.cat([torch.cat([Rt, t], -1), RT[:, 3:4, :]], 1)

    def relative_pose_loss(self, RT1, RT2):
        rel = torch.matmul(self._rigid_inverse(RT1), RT2)
        rot = torch.norm(rel[:, :3, :3] - torch.eye(3, device=RT1.device), dim=(1, 2))
        tr = torch.norm(rel[:, :3, 3], dim=1)
        return rot + tr * self.translation_weight

    def _iteration(self, st: dict, phase_b: bool):
        N, G, HW = self.n_imgs, self.n_groups, self.HW
        for opt in st["opts"] if phase_b else st["opts"][:1]:
            opt.zero_grad(set_to_none=True)
        lr_now = st["lr_table"].index_select(0, st["it"].long())
        st["lr_a"].copy_(lr_now.reshape(()))
        if phase_b:
            st["lr_b"].copy_(lr_now.reshape(()))
        P = self.get_im_poses()
        Pw = self.get_pw_poses()
        invf = (-self.im_focals / self.focal_break).exp()
        with torch.no_grad():
            st["poses"].copy_(P[:, :3, :].reshape(N, 12))
            st["S"].copy_(Pw[:, :3, :].reshape(G, 12))
            st["invf"].copy_(invf.reshape(1))
            if phase_b:
                st["st"][:, 0].copy_(self.s_depth[:, 0])
                st["st"][:, 1].copy_(self.t_depth[:, 0])
        ops.check(ops.lib().geo4d_align_iter(
            ops._vp(self.im_depthmaps), ops._vp(st["m"]), ops._vp(st["v"]), ops._vp(self._stacked_pred_all),
            ops._vp(self._weight_all), ops._vp(self._stacked_depthmap_all if self.has_invdepth else None),
            ops._vp(self._edge_ptr), ops._vp(self._edge_idx), ops._vp(st["poses"]), ops._vp(st["S"]),
            ops._vp(st["scal"]), ops._vp(st["invf"]), ops._vp(st["it"]), ops._vp(st["st"]), ops._vp(st["gpose"]),
            ops._vp(st["gS"]), ops._vp(st["gscal"]), ops._vp(st["gst"]), N, G, HW, self.W, self.group_size,
            self.max_edges_per_image, ops._s()), "geo4d_align_iter")
        # surrogate whose gradient w.r.t. the small parameters equals that of the dense objective
        sur = (st["gpose"].float().view(N, 3, 4) * P[:, :3, :]).sum() + \
              (st["gS"].float().view(G, 3, 4) * Pw[:, :3, :]).sum() + st["gscal"][0].float() * invf.sum()
        if phase_b and self.has_invdepth:
            gst = st["gst"].float().view(G, 2)
            sur = sur + (gst[:, 0] * self.s_depth[:, 0]).sum() + (gst[:, 1] * self.t_depth[:, 0]).sum()
        if phase_b and self.has_traj and len(self.valid_traj_group_list) > 0:
            vg = st["valid_groups"]
            scale = self.traj_align_poses[:, -1].exp()[vg]
            RT = poses_from_params(self.traj_align_poses)[vg]
            tr = self._stacked_traj_all.view(G, self.group_size, 4, 4)[vg]
            xyz = tr[:, :, :3, 3:4] * scale.reshape(-1, 1, 1, 1)
            top = torch.cat([tr[:, :, :3, :3], xyz], -1)
            homo = torch.cat([top, tr[:, :, 3:4, :]], -2)
            homo = torch.matmul(RT.unsqueeze(1), homo).reshape(-1, 4, 4)
            sur = sur + 0.005 * self.relative_pose_loss(homo, P[st["valid_idx"]]).sum()
        if self.temporal_smoothing_weight > 0 and N > 1:
            sur = sur + self.temporal_smoothing_weight * self.relative_pose_loss(P[:-1], P[1:]).sum()
        sur.backward()
        st["opts"][0].step()
        if phase_b:
            st["opts"][1].step()
        ops.advance_counter(st["it"], 1)

    def _iteration_fused(self, st: dict):
        """dense kernel + small-parameter kernel; no torch ops, no host sync"""
        N, G, HW = self.n_imgs, self.n_groups, self.HW
        ops.check(ops.lib().geo4d_align_iter(
            ops._vp(self.im_depthmaps), ops._vp(st["m"]), ops._vp(st["v"]), ops._vp(self._stacked_pred_all),
            ops._vp(self._weight_all), ops._vp(self._stacked_depthmap_all if self.has_invdepth else None),
            ops._vp(self._edge_ptr), ops._vp(self._edge_idx), ops._vp(st["poses"]), ops._vp(st["S"]),
            ops._vp(st["scal"]), ops._vp(st["invf"]), ops._vp(st["it"]), ops._vp(st["st"]), ops._vp(st["gpose"]),
            ops._vp(st["gS"]), ops._vp(st["gscal"]), ops._vp(st["gst"]), N, G, HW, self.W, self.group_size,
            self.max_edges_per_image, ops._s()), "geo4d_align_iter")
        import ctypes as C
        ops.check(ops.lib().geo4d_align_small_step(
            ops._vp(self.im_poses), ops._vp(self.im_focals), ops._vp(self.pw_poses), ops._vp(self.s_depth),
            ops._vp(self.t_depth), ops._vp(self.traj_align_poses), ops._vp(st["adam_small"]), ops._vp(st["gpose"]),
            ops._vp(st["gS"]), ops._vp(st["gscal"]), ops._vp(st["gst"]), ops._vp(st["traj16"]), ops._vp(st["e_img"]),
            ops._vp(self._edge_ptr), ops._vp(self._edge_idx), ops._vp(st["valid_traj"]), ops._vp(st["scal"]),
            ops._vp(st["it"]), ops._vp(st["poses"]), ops._vp(st["S"]), ops._vp(st["invf"]), ops._vp(st["st"]), N, G,
            self.group_size, st["start_b"], C.c_float(self.temporal_smoothing_weight),
            C.c_float(self.translation_weight), C.c_float(self.base_scale), C.c_float(self.focal_break), ops._s()),
            "geo4d_align_small_step")
        ops.advance_counter(st["it"], 1)

    @torch.no_grad()
    def _refresh_matrices(self, st):
        """pose / sim(3) / focal matrices for the dense kernel from the current parameters"""
        N, G = self.n_imgs, self.n_groups
        st["poses"].copy_(self.get_im_poses()[:, :3, :].reshape(N, 12))
        st["S"].copy_(self.get_pw_poses()[:, :3, :].reshape(G, 12))
        st["invf"].copy_((-self.im_focals / self.focal_break).exp().reshape(1))
        st["st"][:, 0].copy_(self.s_depth[:, 0])
        st["st"][:, 1].copy_(self.t_depth[:, 0])

    def _run_phase_fused(self, st, it0, it1):
        if it1 <= it0:
            return
        self._refresh_matrices(st)
        self._iteration_fused(st)  # first iteration eagerly (sets kernel attributes, validates arguments)
        n_left = it1 - it0 - 1
        if n_left <= 0:
            return
        if not self.use_cuda_graph:
            for _ in range(n_left):
                self._iteration_fused(st)
            return
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        n0 = ops.raw_launch_count()
        with torch.cuda.stream(side), ops.capture_graph(graph):
            self._iteration_fused(st)
        torch.cuda.current_stream().wait_stream(side)
        nk = ops.raw_launch_count() - n0
        ops.note_replay(nk, -1)
        for _ in range(n_left):
            graph.replay()
        ops.note_replay(nk, n_left)

    def _run_phase(self, st, it0, it1, phase_b):
        """iterations [it0, it1): a few eager ones (they also serve as the graph warm-up), then replays."""
        it = it0
        n_eager = 3 if self.use_cuda_graph else (it1 - it0)
        while it < it1 and n_eager > 0:
            self._iteration(st, phase_b)
            it += 1
            n_eager -= 1
        if it >= it1:
            return
        self._tick("  eager_iters")
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        n0 = ops.raw_launch_count()
        with torch.cuda.stream(side), ops.capture_graph(graph):
            self._iteration(st, phase_b)
        torch.cuda.current_stream().wait_stream(side)
        nk = ops.raw_launch_count() - n0
        ops.note_replay(nk, -1)
        self._tick("  graph_capture")
        for _ in range(it, it1):
            graph.replay()
        ops.note_replay(nk, it1 - it)

    def _dist_sub(self):
        """(world, rank) for the per-window sub-problems (PnP initialisation, LAD fits) split round-robin over the
        ranks; (1, 0) on one GPU, for a single window, or when sharding is switched off (GEO4D_ALIGN_SHARD=0: a rank
        that runs an alignment on its own must not enter collectives the other ranks do not)."""
        if self.shard_sub_alignments and self.n_groups > 1 and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1 \
                and os.environ.get("GEO4D_ALIGN_SHARD", "1") != "0":
            return torch.distributed.get_world_size(), torch.distributed.get_rank()
        return 1, 0

    # ------------------------------------------------------------------ persistent loop engine
    def _dist(self):
        """(world, rank) of the sharded alignment, (1, 0) when it runs on one GPU / replicated."""
        if self.shard_alignment and self.engine == "loop" and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1 \
                and self.n_imgs >= 2 and os.environ.get("GEO4D_ALIGN_SHARD", "1") != "0":
            return torch.distributed.get_world_size(), torch.distributed.get_rank()
        return 1, 0

    def _setup_loop(self, st):
        """Buffers and the descriptor of geo4d_align_loop; under torch.distributed also the image partition and
        the peer-mapped receive buffers of the in-kernel exchange."""
        from . import sharding
        from ._cabi import AlignLoopDesc
        import ctypes as C
        L = ops.lib()
        L.geo4d_align_loop_part_floats.restype = C.c_size_t
        dev = self.device
        N, G, HW = self.n_imgs, self.n_groups, self.HW
        if HW % 4:
            raise NotImplementedError("engine='loop' needs H*W to be a multiple of 4 (Geo4D sizes are multiples of 16)")
        world, rank = self._dist()
        img_lo = [0, N]
        ex = None
        rec = 0
        if world > 1:
            edges = np.diff(self._edge_ptr.cpu().numpy()).tolist()
            img_lo = sharding.partition_images(edges, world)
            max_loc = max(img_lo[r + 1] - img_lo[r] for r in range(world))
            rec = int(L.geo4d_align_loop_record_doubles(max_loc, G))
            need = 256 + 2 * world * rec * 8
            nbytes = max(1 << 20, 1 << (need - 1).bit_length())
            ex = sharding.peer_exchange(nbytes, dev)
            if ex is None:   # no peer-memory transport on this box: optimise the whole clip on every rank
                world, rank, img_lo, rec = 1, 0, [0, N], 0
        n_lo, n_hi = img_lo[rank], img_lo[rank + 1]
        chunks = int(L.geo4d_align_loop_chunks(max(1, n_hi - n_lo), HW))
        d = AlignLoopDesc()
        st["loop_keep"] = keep = {
            "part": torch.zeros(max(1, int(L.geo4d_align_loop_part_floats(max(1, n_hi - n_lo), chunks))), device=dev),
            "bar": torch.zeros(2, device=dev, dtype=torch.int32),
        }
        P = ops._ptr
        d.logd, d.adam_m, d.adam_v = P(self.im_depthmaps), P(st["m"]), P(st["v"])
        d.pred, d.weight = P(self._stacked_pred_all), P(self._weight_all)
        d.invd = P(self._stacked_depthmap_all) if self.has_invdepth else None
        d.edge_ptr, d.edge_idx, d.scal = P(self._edge_ptr), P(self._edge_idx), P(st["scal"])
        d.poses, d.S, d.invf, d.st = P(st["poses"]), P(st["S"]), P(st["invf"]), P(st["st"])
        d.gpose, d.gS, d.gscal, d.gst = P(st["gpose"]), P(st["gS"]), P(st["gscal"]), P(st["gst"])
        d.part, d.bar = P(keep["part"]), P(keep["bar"])
        d.im_poses, d.im_focal, d.pw_poses = P(self.im_poses), P(self.im_focals), P(self.pw_poses)
        d.s_depth, d.t_depth, d.ta_poses = P(self.s_depth), P(self.t_depth), P(self.traj_align_poses)
        d.adam_small, d.traj, d.e_img, d.valid_traj = P(st["adam_small"]), P(st["traj16"]), P(st["e_img"]), P(st["valid_traj"])
        d.N, d.G, d.HW, d.W, d.group_size = N, G, HW, self.W, self.group_size
        d.max_edges_per_image = self.max_edges_per_image
        d.n_lo, d.n_hi, d.chunks, d.start_b = n_lo, n_hi, chunks, st["start_b"]
        d.temporal_smoothing_weight = self.temporal_smoothing_weight
        d.translation_weight = self.translation_weight
        d.base_scale, d.focal_break = self.base_scale, self.focal_break
        d.world, d.rank, d.rec_doubles = world, rank, rec
        for r in range(world + 1):
            d.img_lo[r] = img_lo[r]
        if ex is not None:
            for r in range(world):
                d.peer_flag[r] = ex.ptrs[r]
                d.peer_rec[r] = ex.ptrs[r] + 256
            d.flag_base = sharding.reserve_flags(self._niter_total)
        if os.environ.get("GEO4D_ALIGN_TRACE", "0") == "1":
            keep["dbg"] = torch.zeros(8, device=dev, dtype=torch.int64)
            d.debug_ns = P(keep["dbg"])
        st["loop_desc"] = d
        self._shard = {"world": world, "rank": rank, "img_lo": img_lo, "transport": ex.how if ex else None}

    def _run_loop(self, st, it0, it1):
        if it1 <= it0:
            return
        if "loop_desc" not in st:
            self._setup_loop(st)
        self._refresh_matrices(st)
        d = st["loop_desc"]
        d.it0, d.it1 = it0, it1
        import ctypes as C
        ops.check(ops.lib().geo4d_align_loop(C.byref(d), ops._s()), "geo4d_align_loop")

    @torch.no_grad()
    def _broadcast_state(self):
        """Sharded alignment: every rank starts from rank 0's initialisation (the GPU reductions of the
        initialisation use atomics, so replicas could differ in the last bit)."""
        for t in (self.im_depthmaps, self.im_poses, self.im_focals, self.pw_poses):
            torch.distributed.broadcast(t.data, src=0)

    @torch.no_grad()
    def _gather_depth(self):
        """Sharded alignment: every rank owns the log-depth maps of its images; make them complete everywhere."""
        sh = self._shard
        if not sh or sh["world"] <= 1:
            return
        lo, hi = sh["img_lo"][sh["rank"]], sh["img_lo"][sh["rank"] + 1]
        d = self.im_depthmaps.data
        d[:lo].zero_()
        d[hi:].zero_()
        torch.distributed.all_reduce(d)

    @torch.no_grad()
    def _current_loss(self, st, phase_b) -> float:
        """Objective value at the current parameters (optimizer_group.py:523), from the kernel's reductions."""
        li = float(st["gscal"][1]) / self.total_area_all
        dl = 2.0 * float(st["gscal"][2]) / self.total_area_all if phase_b else 0.0
        return li + dl

    # ------------------------------------------------------------------ public entry point
    def _tick(self, name):
        if self._profile is not None:
            import time
            torch.cuda.synchronize()
            now = time.time()
            self._profile.append((name, now - self._t_last))
            self._t_last = now

    def compute_global_alignment(self, init=None, save_score_path=None, save_score_only=False, niter_PnP=10,
                                 lr=0.01, niter=300, schedule="cosine", lr_min=1e-3, **kw):
        require_device()
        self._profile = [] if os.environ.get("GEO4D_ALIGN_PROFILE", "0") == "1" else None
        if self._profile is not None:
            import time
            torch.cuda.synchronize()
            self._t_last = time.time()
        if init == "group":
            self._init_from_group(niter_PnP=niter_PnP)
            self._tick("init_from_group")
        elif init is not None:
            raise NotImplementedError(f"init={init!r}: only the 'group' initialisation is used by Geo4D")
        return self._global_alignment_loop(lr=lr, niter=niter, schedule=schedule, lr_min=lr_min)

    def _global_alignment_loop(self, lr, niter, schedule, lr_min):
        dev = self.device
        N, G = self.n_imgs, self.n_groups
        tt = np.arange(niter) / niter
        if schedule == "linear":
            lrs = lr + (lr_min - lr) * tt
        elif schedule == "cosine":
            lrs = lr_min + (lr - lr_min) * (1 + np.cos(tt * np.pi)) / 2
        else:
            raise ValueError(f"bad lr schedule={schedule!r}")
        start_b = self.depth_traj_start_iter if (self.has_invdepth or self.has_traj) else niter
        scal = np.zeros((niter, 8), dtype=np.float32)
        steps = np.arange(1, niter + 1, dtype=np.float64)
        scal[:, 1], scal[:, 2] = self.W / 2, self.H / 2
        scal[:, 3] = lrs
        scal[:, 4] = 1.0 - 0.9 ** steps
        scal[:, 5] = np.sqrt(1.0 - 0.9 ** steps)
        scal[:, 6] = 1.0 / self.total_area_all
        scal[start_b:, 7] = 1.0
        lr_a = torch.tensor(float(lrs[0]), device=dev)
        lr_b = torch.tensor(float(lrs[0]), device=dev)
        opt_a = torch.optim.Adam([self.im_poses, self.im_focals, self.pw_poses], lr=lr_a, betas=(0.9, 0.9),
                                 capturable=True, foreach=True)
        opt_b = torch.optim.Adam([self.s_depth, self.t_depth, self.traj_align_poses], lr=lr_b, betas=(0.9, 0.9),
                                 capturable=True, foreach=True)
        st = {
            "opts": [opt_a, opt_b], "lr_a": lr_a, "lr_b": lr_b,
            "lr_table": torch.tensor(lrs, device=dev, dtype=torch.float32),
            "scal": torch.tensor(scal, device=dev), "it": torch.zeros(1, device=dev, dtype=torch.int32),
            "m": torch.zeros(N, self.HW, device=dev), "v": torch.zeros(N, self.HW, device=dev),
            "poses": torch.zeros(N, 12, device=dev), "S": torch.zeros(G, 12, device=dev),
            "invf": torch.zeros(1, device=dev), "st": torch.ones(G, 3, device=dev),
            "gpose": torch.zeros(N, 12, device=dev, dtype=torch.float64),
            "gS": torch.zeros(G, 12, device=dev, dtype=torch.float64),
            "gscal": torch.zeros(3, device=dev, dtype=torch.float64),
            "gst": torch.zeros(G, 2, device=dev, dtype=torch.float64),
        }
        opt_a.param_groups[0]["lr"] = lr_a
        opt_b.param_groups[0]["lr"] = lr_b
        fused = os.environ.get("GEO4D_ALIGN_AUTOGRAD", "0") != "1"
        if fused:
            f = ops.lib().geo4d_align_small_adam_floats
            import ctypes as C
            f.restype = C.c_size_t
            st["adam_small"] = torch.zeros(int(f(N, G)), device=dev)
            st["traj16"] = (self._stacked_traj_all.reshape(G * self.group_size, 16).contiguous() if self.has_traj
                            else torch.eye(4, device=dev).reshape(1, 16).repeat(G * self.group_size, 1))
            st["e_img"] = self._e_all.to(torch.int32).contiguous()
            st["valid_traj"] = torch.zeros(G, device=dev)
            st["start_b"] = int(start_b)
        with torch.no_grad():
            self._weight_all.clamp_(max=10)  # conf_optimize clip, optimizer_group.py:455-456
        loop = fused and self.engine == "loop"
        self._niter_total = int(niter)
        if loop and self._dist()[0] > 1:
            self._broadcast_state()
        with torch.enable_grad():
            if loop:
                self._run_loop(st, 0, min(start_b, niter))
            elif fused:
                self._run_phase_fused(st, 0, min(start_b, niter))
            else:
                self._run_phase(st, 0, min(start_b, niter), False)
            self._tick("phase_A")
            if niter > start_b:
                if loop:
                    self._gather_depth()
                if self.has_invdepth:
                    self.invalid_depth_group = self._set_st_depth()
                    self._tick("set_st_depth(LAD)")
                    st["st"][:, 2] = 1.0
                    if self.invalid_depth_group:
                        st["st"][self.invalid_depth_group, 2] = 0.0
                if self.has_traj:
                    self.valid_traj_group_list, self.valid_group_idx = self._set_traj()
                    st["valid_groups"] = torch.tensor(self.valid_traj_group_list, device=dev, dtype=torch.long)
                    st["valid_idx"] = torch.tensor(self.valid_group_idx, device=dev, dtype=torch.long)
                if self.verbose:
                    print("invalid_depth_group", self.invalid_depth_group)
                    print("valid_traj_group_list", self.valid_traj_group_list)
                self._tick("set_traj")
                if fused:
                    if self.has_traj and self.valid_traj_group_list:
                        st["valid_traj"][self.valid_traj_group_list] = 1.0
                    if loop:
                        self._run_loop(st, start_b, niter)
                    else:
                        self._run_phase_fused(st, start_b, niter)
                else:
                    self._run_phase(st, start_b, niter, True)
                self._tick("phase_B")
        if loop:
            self._gather_depth()
        torch.cuda.synchronize()
        if loop and "dbg" in st.get("loop_keep", {}):
            t = st["loop_keep"]["dbg"].cpu().numpy().astype(np.float64)
            n = max(t[6], 1.0)
            print("align loop trace (us per iteration, CTA 0): " + ", ".join(
                f"{k}={t[i] / n * 1e-3:.1f}" for i, k in enumerate(("dense", "barrier1", "fold", "exchange", "small", "barrier2"))))
        if self._profile is not None:
            print("align profile:", ", ".join(f"{n}={t * 1e3:.1f}ms" for n, t in self._profile))
        self._state = st
        return self._current_loss(st, niter > start_b)
