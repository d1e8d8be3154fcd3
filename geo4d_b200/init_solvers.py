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

import os

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


# =============================================================================== GPU-reduced variants
# The per-pixel work (41 SQPnP moments per frame, shift/focal sums per window) is reduced on the GPU
# (geo4d_pnp_moments / geo4d_shift_focal_sums); only 9x9 / 1-D problems are solved here, in fp64.

def moments_numpy(pts: np.ndarray, mask: np.ndarray, cx: float, cy: float) -> np.ndarray:
    """Reference (host) computation of the 41 moments of geo4d_pnp_moments, for tests."""
    H, W, _ = pts.shape
    v, u = np.mgrid[:H, :W]
    m = pts[mask].astype(np.float64)
    du, dv = (u[mask] - cx).astype(np.float64), (v[mask] - cy).astype(np.float64)
    r2 = du * du + dv * dv
    out = np.zeros(41)
    out[0:4] = (len(m), du.sum(), dv.sum(), r2.sum())
    out[4:7] = m.sum(0)
    out[7:10] = (du[:, None] * m).sum(0)
    out[10:13] = (dv[:, None] * m).sum(0)
    out[13:16] = (r2[:, None] * m).sum(0)
    mm = np.stack([m[:, 0] * m[:, 0], m[:, 0] * m[:, 1], m[:, 0] * m[:, 2], m[:, 1] * m[:, 1], m[:, 1] * m[:, 2],
                   m[:, 2] * m[:, 2]], 1)
    out[16:22] = mm.sum(0)
    out[22:28] = (du[:, None] * mm).sum(0)
    out[28:34] = (dv[:, None] * mm).sum(0)
    out[34:40] = (r2[:, None] * mm).sum(0)
    out[40] = len(m)
    return out


def _sym6(v):
    return np.array([[v[0], v[1], v[2]], [v[1], v[3], v[4]], [v[2], v[4], v[5]]])


def _nearest_rotation(e: np.ndarray) -> np.ndarray:
    U, _, Vt = np.linalg.svd(e.reshape(3, 3))
    R = U @ Vt
    if np.linalg.det(R) < 0:
        R = U @ np.diag([1.0, 1.0, -1.0]) @ Vt
    return R.reshape(9)


def _sqp_refine(r: np.ndarray, Omega: np.ndarray, max_iter: int = 15, tol: float = 1e-10) -> np.ndarray:
    """Sequential quadratic programming on min r^T Omega r s.t. r in SO(3) (rows of R in r)."""
    for _ in range(max_iter):
        r1, r2, r3 = r[0:3], r[3:6], r[6:9]
        g = np.array([r1 @ r1 - 1, r2 @ r2 - 1, r3 @ r3 - 1, r1 @ r2, r1 @ r3, r2 @ r3])
        J = np.zeros((6, 9))
        J[0, 0:3] = 2 * r1
        J[1, 3:6] = 2 * r2
        J[2, 6:9] = 2 * r3
        J[3, 0:3], J[3, 3:6] = r2, r1
        J[4, 0:3], J[4, 6:9] = r3, r1
        J[5, 3:6], J[5, 6:9] = r3, r2
        # one SQP step = the equality-constrained QP  min (r+d)^T Omega (r+d)  s.t.  J d = -g, solved through its
        # 15x15 KKT system (same step as the null-space form: min-norm solution of J d = -g plus the minimiser
        # over null(J), but one small LU instead of a 6x9 SVD)
        K = np.zeros((15, 15))
        K[:9, :9] = Omega
        K[:9, 9:] = J.T
        K[9:, :9] = J
        rhs = np.concatenate([-(Omega @ r), -g])
        try:
            d = np.linalg.solve(K, rhs)[:9]
        except np.linalg.LinAlgError:
            U, S, Vt = np.linalg.svd(J)
            d_rs = Vt[:6].T @ ((U.T @ (-g)) / S)
            N = Vt[6:].T
            y = -np.linalg.solve(N.T @ Omega @ N, N.T @ Omega @ (r + d_rs))
            d = d_rs + N @ y
        r = r + d
        if d @ d < tol:
            break
    return r


def sqpnp_from_moments(mom: np.ndarray, f: float):
    """SQPnP (cv2.SOLVEPNP_SQPNP) from the reduced moments for focal f.  Returns (R, t) world-to-camera or None."""
    n = mom[0]
    if n < 4 or not np.isfinite(f) or f <= 0:
        return None
    s1, s2 = 1.0 / f, 1.0 / (f * f)
    SQ = np.array([[n, 0, -s1 * mom[1]], [0, n, -s1 * mom[2]], [-s1 * mom[1], -s1 * mom[2], s2 * mom[3]]])
    Sm, Sxm, Sym, Srm = mom[4:7], s1 * mom[7:10], s1 * mom[10:13], s2 * mom[13:16]
    z3 = np.zeros(3)
    QA = np.stack([np.concatenate([Sm, z3, -Sxm]), np.concatenate([z3, Sm, -Sym]), np.concatenate([-Sxm, -Sym, Srm])])
    Mm, Mx, My, Mr = _sym6(mom[16:22]), s1 * _sym6(mom[22:28]), s1 * _sym6(mom[28:34]), s2 * _sym6(mom[34:40])
    Z = np.zeros((3, 3))
    AQA = np.block([[Mm, Z, -Mx], [Z, Mm, -My], [-Mx, -My, Mr]])
    try:
        P = -np.linalg.solve(SQ, QA)
    except np.linalg.LinAlgError:
        return None
    Omega = AQA + QA.T @ P
    Omega = 0.5 * (Omega + Omega.T)
    w, V = np.linalg.eigh(Omega)  # ascending
    mean_pt = Sm / n
    best = None

    def consider(e):
        nonlocal best
        for sgn in (1.0, -1.0):
            r = _sqp_refine(_nearest_rotation(sgn * np.sqrt(3.0) * e), Omega)
            r = _nearest_rotation(r)
            R = r.reshape(3, 3)
            t = P @ r
            if R[2] @ mean_pt + t[2] <= 0:  # cheirality
                continue
            err = float(r @ Omega @ r)
            if best is None or err < best[0]:
                best = (err, R, t)

    consider(V[:, 0])
    k = 1
    while k < 9 and (best is None or best[0] > 3 * w[k]):
        consider(V[:, k])
        k += 1
    if best is None:
        return None
    return best[1], best[2]


def _nearest_rotation_batch(E: np.ndarray) -> np.ndarray:
    """E [B, 9] -> nearest rotations [B, 9] (batched _nearest_rotation)."""
    U, _, Vt = np.linalg.svd(E.reshape(-1, 3, 3))
    R = U @ Vt
    neg = np.linalg.det(R) < 0
    if neg.any():
        D = np.diag([1.0, 1.0, -1.0])
        R[neg] = U[neg] @ D @ Vt[neg]
    return R.reshape(-1, 9)


def _sqp_refine_batch(r: np.ndarray, Omega: np.ndarray, max_iter: int = 15, tol: float = 1e-10) -> np.ndarray:
    """_sqp_refine for B problems at once: r [B, 9], Omega [B, 9, 9] (one batched 15x15 KKT solve per iteration)."""
    r = r.copy()
    B = r.shape[0]
    active = np.ones(B, dtype=bool)
    for _ in range(max_iter):
        r1, r2, r3 = r[:, 0:3], r[:, 3:6], r[:, 6:9]
        g = np.stack([(r1 * r1).sum(1) - 1, (r2 * r2).sum(1) - 1, (r3 * r3).sum(1) - 1, (r1 * r2).sum(1),
                      (r1 * r3).sum(1), (r2 * r3).sum(1)], 1)
        J = np.zeros((B, 6, 9))
        J[:, 0, 0:3] = 2 * r1
        J[:, 1, 3:6] = 2 * r2
        J[:, 2, 6:9] = 2 * r3
        J[:, 3, 0:3], J[:, 3, 3:6] = r2, r1
        J[:, 4, 0:3], J[:, 4, 6:9] = r3, r1
        J[:, 5, 3:6], J[:, 5, 6:9] = r3, r2
        K = np.zeros((B, 15, 15))
        K[:, :9, :9] = Omega
        K[:, :9, 9:] = J.transpose(0, 2, 1)
        K[:, 9:, :9] = J
        rhs = np.concatenate([-np.einsum("bij,bj->bi", Omega, r), -g], 1)
        d = np.linalg.solve(K, rhs[..., None])[:, :9, 0]
        d[~active] = 0.0
        r = r + d
        active &= (d * d).sum(1) >= tol
        if not active.any():
            break
    return r


_SQPNP_FN = None


def sqpnp_from_moments_native(mom: np.ndarray, f: float):
    """sqpnp_from_moments through the library's host solver (geo4d_sqpnp_from_moments, csrc/sqpnp_host.cu):
    same algorithm in C++, ~40x less host time per solve."""
    global _SQPNP_FN
    if _SQPNP_FN is None:
        import ctypes as C
        from ._cabi import lib
        fn = lib().geo4d_sqpnp_from_moments
        fn.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        _SQPNP_FN = fn
    m = np.ascontiguousarray(mom, dtype=np.float64)
    out = np.empty(12, dtype=np.float64)
    ok = _SQPNP_FN(m.ctypes.data, float(f), out.ctypes.data, out.ctypes.data + 72)
    return (out[:9].reshape(3, 3).copy(), out[9:].copy()) if ok else None


_SQPNP_BATCH_FN = None


def sqpnp_from_moments_batch(moms: np.ndarray, focals) -> list:
    """sqpnp_from_moments for B (moments, focal) pairs.  Default: the native host solver on a few host threads
    (geo4d_sqpnp_from_moments_batch); GEO4D_SQPNP=numpy keeps everything in NumPy with the linear algebra batched
    (cases that need more than the smallest eigenvector, or that hit a singular system, fall back to the scalar
    routine)."""
    global _SQPNP_BATCH_FN
    B = len(focals)
    if B == 0:
        return []
    if os.environ.get("GEO4D_SQPNP", "native") != "numpy":
        import ctypes as C
        if _SQPNP_BATCH_FN is None:
            from ._cabi import lib
            fn = lib().geo4d_sqpnp_from_moments_batch
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            fn.restype = C.c_int
            _SQPNP_BATCH_FN = fn
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.float64).reshape(B, -1)[:, :41])
        f = np.ascontiguousarray(np.asarray(focals, dtype=np.float64).reshape(B))
        R, t, ok = np.empty((B, 9)), np.empty((B, 3)), np.zeros(B, dtype=np.int32)
        _SQPNP_BATCH_FN(m.ctypes.data, f.ctypes.data, B, R.ctypes.data, t.ctypes.data, ok.ctypes.data,
                        min(B, 8, os.cpu_count() or 1))
        return [(R[i].reshape(3, 3).copy(), t[i].copy()) if ok[i] else None for i in range(B)]
    out = [None] * B
    moms = np.asarray(moms, dtype=np.float64).reshape(B, -1)
    f = np.asarray(focals, dtype=np.float64)
    good = (moms[:, 0] >= 4) & np.isfinite(f) & (f > 0)
    idx = np.nonzero(good)[0]
    if len(idx) == 0:
        return out
    try:
        m = moms[idx]
        n = m[:, 0]
        s1, s2 = 1.0 / f[idx], 1.0 / (f[idx] * f[idx])
        k = len(idx)
        SQ = np.zeros((k, 3, 3))
        SQ[:, 0, 0] = n; SQ[:, 1, 1] = n
        SQ[:, 0, 2] = SQ[:, 2, 0] = -s1 * m[:, 1]
        SQ[:, 1, 2] = SQ[:, 2, 1] = -s1 * m[:, 2]
        SQ[:, 2, 2] = s2 * m[:, 3]
        Sm, Sxm, Sym, Srm = m[:, 4:7], s1[:, None] * m[:, 7:10], s1[:, None] * m[:, 10:13], s2[:, None] * m[:, 13:16]
        QA = np.zeros((k, 3, 9))
        QA[:, 0, 0:3], QA[:, 0, 6:9] = Sm, -Sxm
        QA[:, 1, 3:6], QA[:, 1, 6:9] = Sm, -Sym
        QA[:, 2, 0:3], QA[:, 2, 3:6], QA[:, 2, 6:9] = -Sxm, -Sym, Srm

        def sym(v):
            M = np.empty((k, 3, 3))
            M[:, 0, 0], M[:, 0, 1], M[:, 0, 2] = v[:, 0], v[:, 1], v[:, 2]
            M[:, 1, 0], M[:, 1, 1], M[:, 1, 2] = v[:, 1], v[:, 3], v[:, 4]
            M[:, 2, 0], M[:, 2, 1], M[:, 2, 2] = v[:, 2], v[:, 4], v[:, 5]
            return M
        Mm, Mx, My, Mr = sym(m[:, 16:22]), s1[:, None, None] * sym(m[:, 22:28]), s1[:, None, None] * sym(m[:, 28:34]), \
            s2[:, None, None] * sym(m[:, 34:40])
        AQA = np.zeros((k, 9, 9))
        AQA[:, 0:3, 0:3] = Mm; AQA[:, 3:6, 3:6] = Mm; AQA[:, 6:9, 6:9] = Mr
        AQA[:, 0:3, 6:9] = -Mx; AQA[:, 6:9, 0:3] = -Mx
        AQA[:, 3:6, 6:9] = -My; AQA[:, 6:9, 3:6] = -My
        P = -np.linalg.solve(SQ, QA)
        Omega = AQA + QA.transpose(0, 2, 1) @ P
        Omega = 0.5 * (Omega + Omega.transpose(0, 2, 1))
        w, V = np.linalg.eigh(Omega)
        e0 = V[:, :, 0]
        mean_pt = Sm / n[:, None]
        start = np.concatenate([np.sqrt(3.0) * e0, -np.sqrt(3.0) * e0], 0)            # both signs
        Om2 = np.concatenate([Omega, Omega], 0)
        r = _nearest_rotation_batch(_sqp_refine_batch(_nearest_rotation_batch(start), Om2))
        t = np.einsum("bij,bj->bi", np.concatenate([P, P], 0), r)
        cheir = (r[:, 6:9] * np.concatenate([mean_pt, mean_pt], 0)).sum(1) + t[:, 2] > 0
        err = np.einsum("bi,bij,bj->b", r, Om2, r)
        for j, i in enumerate(idx):
            best = None
            for c in (j, j + k):
                if cheir[c] and (best is None or err[c] < best[0]):
                    best = (float(err[c]), r[c].reshape(3, 3), t[c])
            if best is None or best[0] > 3 * w[j, 1]:
                out[i] = sqpnp_from_moments(moms[i], float(f[i]))   # rare: look at more eigenvectors
            else:
                out[i] = (best[1], best[2])
    except np.linalg.LinAlgError:
        for i in idx:
            out[i] = sqpnp_from_moments(moms[i], float(f[i]))
    return out


def gpu_fast_pnp_frames(ops, pts: "torch.Tensor", conf: "torch.Tensor", H: int, W: int, first_focal_of, im_focals,
                        im_poses, frame_ids, niter_PnP: int = 10, thr_px: float = 5.0):
    """fast_pnp (init_im_poses.py:824-865) for the frames of one window with the reductions on the GPU:
    for each frame and each tentative focal, SQPnP on all masked points -> consensus set (reprojection error
    < 5 px, the RANSAC criterion) -> SQPnP re-fit on the consensus set; the focal with the most inliers wins.
    Frames are processed in order because frame k's tentative focals derive from frame k-1's result."""
    F = len(frame_ids)
    HW = H * W
    cx, cy = W / 2, H / 2
    S = max(W, H)
    mom_all = ops.pnp_moments(pts, conf, F, HW, W, cx, cy).cpu().numpy()[:, 0]  # focal-independent
    for k, img in enumerate(frame_ids):
        focal = first_focal_of(k, img)
        if focal is None:
            tentative = list(np.geomspace(S / 2, S * 3, 63))
        else:
            lo, hi = -0.03 * S + focal, 0.03 * S + focal
            tentative = [focal] + ([float(x) for x in np.geomspace(lo, hi, 2)] if lo > 0 else [])
        tentative = [float(f) for f in tentative if np.isfinite(f) and f > 0]
        sols = sqpnp_from_moments_batch(np.repeat(mom_all[k][None], len(tentative), 0), tentative)
        ok = [i for i, s in enumerate(sols) if s is not None]
        if ok:
            gate = np.zeros((1, len(ok), 13), dtype=np.float32)
            for j, i in enumerate(ok):
                R, t = sols[i]
                gate[0, j, :12] = np.concatenate([R, t[:, None]], 1).reshape(12)
                gate[0, j, 12] = tentative[i]
            import torch
            g = torch.from_numpy(gate).to(pts.device)
            mom_in = ops.pnp_moments(pts[k:k + 1], conf[k:k + 1], 1, HW, W, cx, cy, gate=g, ncand=len(ok),
                                     thr_px=thr_px).cpu().numpy()[0]
            refit = sqpnp_from_moments_batch(mom_in[:len(ok)], [tentative[i] for i in ok])
            best = (0, None, None)
            for j, i in enumerate(ok):
                ninl = int(round(mom_in[j, 40]))
                if ninl < 4 or ninl <= best[0]:
                    continue
                if refit[j] is not None:
                    best = (ninl, refit[j], tentative[i])
            if best[0]:
                R, t = best[1]
                w2c = np.eye(4)
                w2c[:3, :3], w2c[:3, 3] = R, t
                im_focals[img], im_poses[img] = best[2], np.linalg.inv(w2c)
        if im_poses[img] is None:
            im_poses[img] = np.eye(4)


def _minimal_sample_moments(pts, conf, cand, W, cx, cy):
    """41 SQPnP moments of 4-point samples.  pts [..., D, 3], conf [..., D] hold D >= 4 candidate pixels per sample
    (pixel indices `cand`); the first 4 with conf > 0.5 form the sample.  Returns (mom [..., 41], usable [...])."""
    valid = conf > 0.5
    order = np.argsort(~valid, axis=-1, kind="stable")[..., :4]          # first four valid draws
    usable = np.take_along_axis(valid, order, -1).all(-1)
    m = np.take_along_axis(pts, order[..., None], -2).astype(np.float64)  # [..., 4, 3]
    pix = np.take_along_axis(cand, order, -1)
    du, dv = (pix % W) - cx, (pix // W) - cy
    r2 = du * du + dv * dv
    mom = np.zeros(m.shape[:-2] + (41,))
    mom[..., 0], mom[..., 1], mom[..., 2], mom[..., 3] = 4.0, du.sum(-1), dv.sum(-1), r2.sum(-1)
    mom[..., 4:7] = m.sum(-2)
    mom[..., 7:10] = (du[..., None] * m).sum(-2)
    mom[..., 10:13] = (dv[..., None] * m).sum(-2)
    mom[..., 13:16] = (r2[..., None] * m).sum(-2)
    mm = np.stack([m[..., 0] * m[..., 0], m[..., 0] * m[..., 1], m[..., 0] * m[..., 2], m[..., 1] * m[..., 1],
                   m[..., 1] * m[..., 2], m[..., 2] * m[..., 2]], -1)          # [..., 4, 6]
    mom[..., 16:22] = mm.sum(-2)
    mom[..., 22:28] = (du[..., None] * mm).sum(-2)
    mom[..., 28:34] = (dv[..., None] * mm).sum(-2)
    mom[..., 34:40] = (r2[..., None] * mm).sum(-2)
    mom[..., 40] = 4.0
    return mom, usable


def gpu_fast_pnp_windows(ops, pred: "torch.Tensor", conf: "torch.Tensor", H: int, W: int, focal_group,
                         niter_PnP: int = 10, thr_px: float = 5.0, seed: int = 0):
    """fast_pnp (init_im_poses.py:824-865) for every frame of several windows, each in its window's OWN frame.

    pred [Gm, gs, HW, 3], conf [Gm, gs, HW]: the windows this rank initialises.  PnP is equivariant under the
    window's sim(3) registration (a point map s R X + T seen through the same pixels gives the camera
    [R_c2w' | t'] = [R R_c2w | s R t_c2w + T], identical consensus sets), so it does not have to wait for the
    sequential window chain of align_group_prefix: windows are independent here, which is what lets the ranks
    split them, and frame k of all Gm windows shares ONE moments launch and ONE device -> host read.

    RANSAC as the reference runs it (cv2.solvePnPRansac, iterationsCount = niter_PnP, 5 px, SQPNP): per frame the
    hypotheses are the SQPnP solutions of `niter_PnP` seeded minimal 4-point samples plus the least-squares solution
    over all masked points (one per tentative focal); every hypothesis is scored by its consensus set in a single
    launch (geo4d_pnp_moments with one candidate per hypothesis), the best one is re-fitted on its consensus set --
    what OpenCV does after its RANSAC loop.  Ties go to the all-points solution, so clean point maps give the same
    result as a plain fit -> gate -> refit.  Inside a window the reference's focal chain is kept: frame 0 starts
    from the window's focal (the LM fit), frame k from frame k-1's result (after a failed PnP: the last successful
    one in this window).

    Returns (focals [Gm, gs] (nan = PnP failed), c2w [Gm, gs, 4, 4] (window frame), ok [Gm, gs] bool)."""
    import torch
    Gm, gs = int(pred.shape[0]), int(pred.shape[1])
    HW = H * W
    cx, cy = W / 2, H / 2
    S = max(W, H)
    dev = pred.device
    n_hyp = max(0, int(niter_PnP)) if os.environ.get("GEO4D_PNP_RANSAC", "1") != "0" else 0
    # frame-major copies: frame k of every window contiguous
    pts_t = pred.transpose(0, 1).contiguous()      # [gs, Gm, HW, 3]
    conf_t = conf.transpose(0, 1).contiguous()     # [gs, Gm, HW]
    mom_all = ops.pnp_moments(pts_t.view(gs * Gm, HW, 3), conf_t.view(gs * Gm, HW), gs * Gm, HW, W, cx, cy) \
        .cpu().numpy()[:, 0].reshape(gs, Gm, -1)   # focal-independent moments of every frame, one read
    mom_min = usable = None
    if n_hyp:
        D = 16
        cand = np.random.default_rng(seed).integers(0, HW, size=(gs, Gm, n_hyp * D))
        idx = torch.from_numpy(cand).to(dev)
        sp = torch.gather(pts_t, 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3)).cpu().numpy()
        sc = torch.gather(conf_t, 2, idx).cpu().numpy()
        mom_min, usable = _minimal_sample_moments(sp.reshape(gs, Gm, n_hyp, D, 3), sc.reshape(gs, Gm, n_hyp, D),
                                                  cand.reshape(gs, Gm, n_hyp, D), W, cx, cy)
    focals = np.full((Gm, gs), np.nan)
    c2w = np.tile(np.eye(4), (Gm, gs, 1, 1))
    ok = np.zeros((Gm, gs), dtype=bool)
    prev = [float(f) if f is not None and np.isfinite(f) and f > 0 else None for f in focal_group]
    for k in range(gs):
        # ---- hypotheses of frame k of every window: (moments, focal) problems, solved in one batch
        probs_m, probs_f, owner = [], [], []
        for g in range(Gm):
            focal = prev[g]
            if focal is None:
                t = list(np.geomspace(S / 2, S * 3, 63))
            else:
                lo, hi = -0.03 * S + focal, 0.03 * S + focal
                t = [focal] + ([float(x) for x in np.geomspace(lo, hi, 2)] if lo > 0 else [])
            t = [float(f) for f in t if np.isfinite(f) and f > 0]
            for f in t:                                   # least-squares hypothesis per tentative focal
                probs_m.append(mom_all[k, g]); probs_f.append(f); owner.append(g)
            if n_hyp and t:
                for h in range(n_hyp):                     # minimal-sample hypotheses at the leading focal
                    if usable[k, g, h]:
                        probs_m.append(mom_min[k, g, h]); probs_f.append(t[0]); owner.append(g)
        if not probs_f:
            continue
        sols = sqpnp_from_moments_batch(np.stack(probs_m), probs_f)
        hyp = [[] for _ in range(Gm)]                      # per window: (R, t, focal)
        for sol, f, g in zip(sols, probs_f, owner):
            if sol is not None:
                hyp[g].append((sol[0], sol[1], f))
        C = max(len(x) for x in hyp)
        if C == 0:
            continue
        gate = np.zeros((Gm, C, 13), dtype=np.float32)
        gate[:, :, 11] = -1.0                      # padding candidates: R = 0, t_z = -1 puts every point behind the camera
        gate[:, :, 12] = 1.0
        for g in range(Gm):
            for j, (R, t, f) in enumerate(hyp[g]):
                gate[g, j, :12] = np.concatenate([R, t[:, None]], 1).reshape(12)
                gate[g, j, 12] = f
        mom_in = ops.pnp_moments(pts_t[k], conf_t[k], Gm, HW, W, cx, cy, gate=torch.from_numpy(gate).to(dev), ncand=C,
                                 thr_px=thr_px).cpu().numpy()
        # ---- best consensus set per window (ties: earliest hypothesis = the all-points fit), re-fitted on it
        order = [sorted(range(len(hyp[g])), key=lambda j: (-int(round(mom_in[g, j, 40])), j)) for g in range(Gm)]
        for attempt in range(3):
            todo = [(g, order[g][attempt]) for g in range(Gm)
                    if not ok[g, k] and attempt < len(order[g]) and int(round(mom_in[g, order[g][attempt], 40])) >= 4]
            if not todo:
                break
            refit = sqpnp_from_moments_batch(np.stack([mom_in[g, j] for g, j in todo]), [hyp[g][j][2] for g, j in todo])
            for (g, j), sol in zip(todo, refit):
                if sol is None:
                    continue
                w2c = np.eye(4)
                w2c[:3, :3], w2c[:3, 3] = sol
                focals[g, k], c2w[g, k], ok[g, k] = hyp[g][j][2], np.linalg.inv(w2c), True
                prev[g] = float(hyp[g][j][2])
    return focals, c2w, ok


def gpu_focal_per_group(ops, ref_pts: "torch.Tensor", ref_conf: "torch.Tensor", H: int, W: int):
    """focal_per_group with the sums on the GPU: 1-D damped Newton on the z-shift from s = 0 (the reference runs
    scipy's LM from the same start; both stop at the local minimiser of sum |f p - uv|^2)."""
    import torch
    G = ref_pts.shape[0]
    HW = H * W
    zoff = float(1.0 - ref_pts[..., 2].min())  # z - min(z over all reference frames) + 1
    shift = torch.zeros(G, device=ref_pts.device)

    def evaluate(s):
        o = ops.shift_focal_sums(ref_pts, ref_conf, G, HW, W, H, s, zoff).cpu().numpy()
        S1, S2, d1, d2, uv2 = o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4]
        with np.errstate(divide="ignore", invalid="ignore"):
            cost = uv2 - S1 * S1 / S2
            grad = -(2 * S1 * d1 * S2 - S1 * S1 * d2) / (S2 * S2)
        return cost, grad, S1 / S2, o[:, 5]

    s = np.zeros(G)
    cost, grad, foc, npts = evaluate(shift)
    h = 1e-3
    for _ in range(30):
        # secant Newton on the 1-D gradient
        shift.copy_(torch.from_numpy((s + h).astype(np.float32)))
        _, grad_h, _, _ = evaluate(shift)
        curv = (grad_h - grad) / h
        step = np.where(curv > 0, -grad / np.where(curv > 0, curv, 1.0), -np.sign(grad) * 0.1)
        step = np.clip(step, -0.5, 0.5)
        s_new = s + step
        shift.copy_(torch.from_numpy(s_new.astype(np.float32)))
        cost_new, grad_new, foc_new, _ = evaluate(shift)
        better = cost_new <= cost
        s = np.where(better, s_new, s)
        done = np.abs(np.where(better, cost - cost_new, 0.0)) <= 1e-3 * 1e-3 * np.maximum(cost, 1e-30)
        cost, grad, foc = np.where(better, cost_new, cost), np.where(better, grad_new, grad), np.where(better, foc_new, foc)
        if np.all(done | ~better):
            break
    if not np.all(np.isfinite(foc)) or np.any(npts < 1):
        raise ValueError("shift/focal fit failed")
    foc = torch.tensor(foc, dtype=torch.float32)
    diag = (H ** 2 + W ** 2) ** 0.5
    fx = 0.5 / torch.tan(torch.atan(W / diag / foc))
    fy = 0.5 / torch.tan(torch.atan(H / diag / foc))
    focal_group = ((fx * W) + (fy * H)) / 2
    mean_f = focal_group[focal_group > 30].mean()
    rel = torch.abs(focal_group - mean_f) / mean_f
    focal_group[rel > 0.6] = mean_f
    return focal_group.numpy().tolist()
