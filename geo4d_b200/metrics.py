"""The repo's own depth / pose metrics and trajectory writers.

* `depth_evaluation`  -- dust3r/depth_eval.py:147-359 (same arguments, same four return values).  The LAD
  alignment (`align_with_lad2`, depth_eval.py:112-145) runs as ONE cooperative kernel launch (geo4d_lad_fit:
  same Adam recurrence, same |delta loss| < tol exit) when the tensors live on a CUDA device; on CPU tensors
  it is the reference's torch loop (the reference runs it on the CPU too when use_gpu=False).
* `eval_metrics`      -- dust3r/utils/vo_eval.py:174-257: ATE (Sim(3)-aligned translation RMSE) and RPE
  translation / rotation (delta = 1 frame, all pairs) as the reference obtains them from `evo`.  evo is an
  un-vendored, unpinned dependency (requirements.txt:47): its published algorithm is restated here (Umeyama
  1991 alignment, APE / RPE definitions of evo.core.metrics) -- "parity unpinned", analytic known-answer tests
  only (tests/test_metrics_cpu.py).
* `save_trajectory_tum_format` / `load_tum_trajectory` -- vo_eval.py:465-473: `ts x y z qw qx qy qz`.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------- depth
def depth2disparity(depth: torch.Tensor) -> torch.Tensor:
    """dust3r/depth_eval.py:90-98: 1/depth where depth > 0, else 0."""
    disp = torch.zeros_like(depth)
    nz = depth > 0
    disp[nz] = 1.0 / depth[nz]
    return disp


def absolute_value_scaling2(pred: torch.Tensor, gt: torch.Tensor, s_init: float = 1.0, t_init: float = 0.0,
                            lr: float = 1e-4, max_iters: int = 1000, tol: float = 1e-6) -> Tuple[float, float]:
    """min_{s,t} sum |s x + t - y| by Adam (depth_eval.py:112-145)."""
    if pred.is_cuda:
        from . import ops
        n = pred.numel()
        x = pred.reshape(1, n).float().contiguous()
        y = gt.reshape(1, n).float().contiguous()
        state = torch.zeros(1, 9, device=pred.device)
        state[0, 0] = s_init
        state[0, 1] = t_init
        acc = torch.zeros(4, device=pred.device, dtype=torch.float64)
        ops.lad_fit(x, y, n, 1, state, acc, lr, int(max_iters), tol)
        st = state[0, :2].cpu()
        return float(st[0]), float(st[1])
    s = torch.tensor([s_init], requires_grad=True, dtype=pred.dtype)
    t = torch.tensor([t_init], requires_grad=True, dtype=pred.dtype)
    opt = torch.optim.Adam([s, t], lr=lr)
    prev = None
    with torch.enable_grad():
        for _ in range(max_iters):
            opt.zero_grad()
            loss = torch.sum(torch.abs(s * pred + t - gt))
            loss.backward()
            opt.step()
            cur = float(loss.detach())
            if prev is not None and abs(prev - cur) < tol:
                break
            prev = cur
    return float(s.detach()), float(t.detach())


@torch.no_grad()
def depth_evaluation(predicted_depth_original, ground_truth_depth_original, max_depth=80, custom_mask=None,
                     post_clip_min=None, post_clip_max=None, pre_clip_min=None, pre_clip_max=None,
                     align_with_lstsq=False, align_with_lad=False, align_with_lad2=False, lr=1e-4, max_iters=1000,
                     use_gpu=False, align_with_scale=False, disp_input=False, align_mask=None, return_st=False):
    """dust3r/depth_eval.py:147-359.  Returns (metrics dict, error map, aligned prediction, masked ground truth)."""
    as_t = lambda a: torch.from_numpy(a) if isinstance(a, np.ndarray) else a
    pred_o, gt_o = as_t(predicted_depth_original), as_t(ground_truth_depth_original)
    custom_mask = as_t(custom_mask) if custom_mask is not None else None
    if align_with_lad:
        raise NotImplementedError("align_with_lad (scipy.optimize.minimize) is not used by the Geo4D scripts; "
                                  "use align_with_lad2")
    if pred_o.dim() == 3:
        w = pred_o.shape[-1]
        pred_o, gt_o = pred_o.reshape(-1, w), gt_o.reshape(-1, w)
        if custom_mask is not None:
            custom_mask = custom_mask.reshape(-1, w)
    if use_gpu:
        pred_o, gt_o = pred_o.cuda(), gt_o.cuda()
    mask = (gt_o > 0) & (gt_o < max_depth) if max_depth is not None else (gt_o > 0)
    pred, gt = pred_o[mask], gt_o[mask]
    if align_mask is not None:
        # only used while fitting the alignment; the errors are computed over every valid ground-truth pixel
        align_mask = as_t(align_mask).to(mask.device).bool().reshape(mask.shape)[mask]
    if pre_clip_min is not None:
        pred = torch.clamp(pred, min=pre_clip_min)
    if pre_clip_max is not None:
        pred = torch.clamp(pred, max=pre_clip_max)
    if disp_input:
        real_gt = gt.clone()
        gt = 1 / (gt + 1e-8)
    s = t = scale_factor = None
    sel = (lambda v: v) if align_mask is None else (lambda v: v[align_mask])
    if align_with_lstsq:
        if align_mask is not None:
            raise NotImplementedError
        A = torch.stack([pred.double().cpu(), torch.ones_like(pred, dtype=torch.float64).cpu()], 1).numpy()
        sol = np.linalg.lstsq(A, gt.double().cpu().numpy().reshape(-1, 1), rcond=None)[0]
        s = torch.tensor(sol[0], device=pred_o.device).to(pred.dtype)
        t = torch.tensor(sol[1], device=pred_o.device).to(pred.dtype)
        pred = s * pred + t
    elif align_with_lad2:
        s_init = float(torch.median(sel(gt)) / torch.median(sel(pred)))
        s, t = absolute_value_scaling2(sel(pred), sel(gt), s_init=s_init, lr=lr, max_iters=max_iters)
        pred = s * pred + t
    elif align_with_scale:
        if align_mask is not None:
            raise NotImplementedError
        s = torch.nanmean(gt) / torch.nanmean(pred)
        for _ in range(10):   # Weiszfeld IRLS, depth_eval.py:235-245
            wts = 1.0 / ((s * pred - gt).abs() + 1e-8)
            s = torch.sum(wts * pred * gt) / torch.sum(wts * pred ** 2)
        s = s.clamp(min=1e-3)
        pred = s * pred
    else:
        scale_factor = torch.median(sel(gt)) / torch.median(sel(pred))
        pred = pred * scale_factor
    if disp_input:
        gt = real_gt
        pred = depth2disparity(pred)
    if post_clip_min is not None:
        pred = torch.clamp(pred, min=post_clip_min)
    if post_clip_max is not None:
        pred = torch.clamp(pred, max=post_clip_max)
    mask_within = None
    if custom_mask is not None:
        assert custom_mask.shape == gt_o.shape
        mask_within = custom_mask.to(mask.device)[mask]
        pred, gt = pred[mask_within], gt[mask_within]
    n_valid = int(mask.sum()) if custom_mask is None else int(mask_within.sum())
    if n_valid == 0:
        abs_rel = sq_rel = rmse = log_rmse = d1 = d2 = d3 = 0
    else:
        abs_rel = float(torch.mean(torch.abs(pred - gt) / gt))
        sq_rel = float(torch.mean(((pred - gt) ** 2) / gt))
        rmse = float(torch.sqrt(torch.mean((pred - gt) ** 2)))
        pc = torch.clamp(pred, min=1e-5)
        log_rmse = float(torch.sqrt(torch.mean((torch.log(pc) - torch.log(gt)) ** 2)))
        ratio = torch.maximum(pc / gt, gt / pc)
        d1 = float(torch.mean((ratio < 1.25).float()))
        d2 = float(torch.mean((ratio < 1.25 ** 2).float()))
        d3 = float(torch.mean((ratio < 1.25 ** 3).float()))
    # error map of the aligned full prediction (depth_eval.py:317-341)
    if align_with_lstsq or align_with_lad2:
        full = pred_o * s + t
    elif align_with_scale:
        full = pred_o * s
    else:
        full = pred_o * scale_factor
    if disp_input:
        full = depth2disparity(full)
    err = torch.abs(full - gt_o) / gt_o
    err_full = torch.where(mask, err, torch.zeros_like(gt_o))
    gt_full = torch.where(mask, gt_o, torch.zeros_like(gt_o))
    results = {"Abs Rel": abs_rel, "Sq Rel": sq_rel, "RMSE": rmse, "Log RMSE": log_rmse, "δ < 1.25": d1,
               "δ < 1.25^2": d2, "δ < 1.25^3": d3, "valid_pixels": n_valid}
    if return_st:
        results["s"], results["t"] = s, t
    return results, err_full, full, gt_full


def average_depth_metrics(gathered: Sequence[dict]) -> dict:
    """valid-pixel weighted mean over sequences (infer_geo4d.py:612-619)."""
    wts = [m["valid_pixels"] for m in gathered]
    return {k: float(np.average([m[k] for m in gathered], weights=wts)) for k in gathered[0] if k != "valid_pixels"}


# ----------------------------------------------------------------------------------------------- poses
def tum_to_matrices(traj: np.ndarray) -> np.ndarray:
    """[n, 7] xyz + quaternion wxyz (the reference's `make_traj`, vo_eval.py:161-171) -> [n, 4, 4]."""
    traj = np.asarray(traj, dtype=np.float64)
    q = traj[:, 3:7] / np.linalg.norm(traj[:, 3:7], axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    P = np.tile(np.eye(4), (len(traj), 1, 1))
    P[:, :3, :3] = R
    P[:, :3, 3] = traj[:, :3]
    return P


def umeyama_alignment(x: np.ndarray, y: np.ndarray, with_scale: bool = True):
    """Least-squares Sim(3) y ~ c R x + t for point sets [3, n] (Umeyama 1991; evo.core.geometry.umeyama_alignment
    as called at vo_eval.py:341).  Returns (R, t, c)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    n = x.shape[1]
    mx, my = x.mean(1), y.mean(1)
    sx = (np.linalg.norm(x - mx[:, None], axis=0) ** 2).sum() / n
    cov = (y - my[:, None]) @ (x - mx[:, None]).T / n
    U, D, Vt = np.linalg.svd(cov)
    if np.count_nonzero(D > np.finfo(D.dtype).eps) < 2:
        raise ValueError("Degenerate covariance rank, Umeyama alignment is not possible")
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    c = float(np.trace(np.diag(D) @ S) / sx) if with_scale else 1.0
    t = my - c * (R @ mx)
    return R, t, c


def _se3_inv(P: np.ndarray) -> np.ndarray:
    out = np.tile(np.eye(4), P.shape[:-2] + (1, 1))
    Rt = np.swapaxes(P[..., :3, :3], -1, -2)
    out[..., :3, :3] = Rt
    out[..., :3, 3] = -(Rt @ P[..., :3, 3:4])[..., 0]
    return out


def align_trajectory(est: np.ndarray, ref: np.ndarray, correct_scale: bool = True) -> np.ndarray:
    """evo `PoseTrajectory3D.align`: Umeyama on the positions, scale applied to the positions, then the rigid
    transform applied to the poses."""
    R, t, c = umeyama_alignment(est[:, :3, 3].T, ref[:, :3, 3].T, correct_scale)
    out = est.copy()
    out[:, :3, 3] *= c
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T[None] @ out


def ape_translation(ref: np.ndarray, est: np.ndarray) -> np.ndarray:
    """evo APE, PoseRelation.translation_part: |t_ref - t_est| per pose."""
    return np.linalg.norm(ref[:, :3, 3] - est[:, :3, 3], axis=1)


def rpe(ref: np.ndarray, est: np.ndarray, delta: int = 1):
    """evo RPE with delta frames and all_pairs=True: E_i = (Q_i^-1 Q_{i+d})^-1 (P_i^-1 P_{i+d}); returns
    (translation norms, rotation angles in degrees)."""
    n = len(ref)
    if n <= delta:
        return np.zeros(0), np.zeros(0)
    i0, i1 = np.arange(0, n - delta), np.arange(delta, n)
    q_rel = _se3_inv(ref[i0]) @ ref[i1]
    p_rel = _se3_inv(est[i0]) @ est[i1]
    E = _se3_inv(q_rel) @ p_rel
    tr = np.linalg.norm(E[:, :3, 3], axis=1)
    cosang = np.clip((np.trace(E[:, :3, :3], axis1=1, axis2=2) - 1) / 2, -1, 1)
    return tr, np.degrees(np.arccos(cosang))


def _rmse(v: np.ndarray) -> float:
    return float(np.sqrt(np.mean(np.square(v)))) if len(v) else 0.0


def eval_metrics(pred_traj, gt_traj, seq: str = "", filename: Optional[str] = None, sample_stride: int = 1):
    """vo_eval.py:174-257: (ATE rmse, RPE-trans rmse, RPE-rot rmse in degrees), every metric after a Sim(3)
    alignment of the estimate to the ground truth.  Trajectories are `[poses [n,7] xyz+wxyz, timestamps [n]]`
    (what `get_tum_poses` returns) or [n,4,4] cam-to-world matrices."""
    def mats(tr):
        if isinstance(tr, (list, tuple)):
            p = np.asarray(tr[0])[::sample_stride]
            return tum_to_matrices(p)
        return np.asarray(tr, np.float64)[::sample_stride]
    est, ref = mats(pred_traj), mats(gt_traj)
    n = min(len(est), len(ref))     # sync.associate_trajectories on identical timestamps
    est, ref = est[:n], ref[:n]
    est_al = align_trajectory(est, ref, correct_scale=True)
    ate = _rmse(ape_translation(ref, est_al))
    tr, rot = rpe(ref, est_al, delta=1)
    rpe_trans, rpe_rot = _rmse(tr), _rmse(rot)
    if filename:
        with open(filename, "w+") as f:
            f.write(f"Seq: {seq} \n\n")
            f.write(f"APE w.r.t. translation part (m) (with Sim(3) Umeyama alignment)\n  rmse\t{ate}\n")
            f.write(f"RPE w.r.t. rotation angle in degrees (deg) for delta = 1 (frames) using all pairs\n  rmse\t{rpe_rot}\n")
            f.write(f"RPE w.r.t. translation part (m) for delta = 1 (frames) using all pairs\n  rmse\t{rpe_trans}\n")
    return ate, rpe_trans, rpe_rot


def save_trajectory_tum_format(traj, filename: str) -> None:
    """vo_eval.py:465-473: one line per pose `timestamp x y z qw qx qy qz` (quaternion kept in wxyz order)."""
    poses, ts = np.asarray(traj[0]), np.asarray(traj[1])
    tostr = lambda a: " ".join(map(str, a))
    with open(filename, "w") as f:
        for i in range(len(poses)):
            f.write(f"{ts[i]} {tostr(poses[i, :3])} {tostr(poses[i, 3:7])}\n")


def load_tum_trajectory(filename: str) -> List[np.ndarray]:
    rows = np.loadtxt(filename, ndmin=2)
    return [rows[:, 1:8], rows[:, 0]]
