#!/usr/bin/env python
"""Compare the aligner initialisation with host (cv2/scipy) vs GPU-reduced solvers on the synthetic scene."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import align as oa
from geo4d_b200.cloud_opt import LightPointCloudGroupOptimizer
from geo4d_b200 import init_solvers as isv, ops

def build(mode):
    os.environ["GEO4D_INIT_SOLVERS"] = mode
    groups, preds, gt = oa.synthetic_scene(T=24, H=32, W=48, noise=0.003)
    views = [[{"idx": (i,)} for i in g] for g in groups]
    pd = [{k: v.cuda() for k, v in p.items()} for p in preds]
    sc = LightPointCloudGroupOptimizer(views, pd, conf="id", conf_optimize=True, verbose=True, shared_focal=True,
                                       num_total_iter=40, temporal_smoothing_weight=0.015, translation_weight=1.0,
                                       depth_traj_start_iter=15, lad_max_iters=300)
    sc._init_from_group()
    return sc, preds

for mode in ("host", "gpu"):
    sc, preds = build(mode)
    print(mode, "focal", float(sc.get_focals()[0]), "init focals", [round(f, 3) for f in sc._init_im_focals[:6]], "...")
    P = sc.get_im_poses().detach().cpu().numpy()
    print(mode, "pose[5] t", P[5, :3, 3].round(4), "pose[20] t", P[20, :3, 3].round(4))
    print(mode, "pw_poses", sc.pw_poses.detach().cpu().numpy().round(4)[:, 4:])
    print(mode, "depth mean", float(sc.get_depthmaps(raw=True).mean()))
G, H, W = 2, 32, 48
groups, preds, gt = oa.synthetic_scene(T=24, H=32, W=48, noise=0.003)
ref_pts = torch.stack([p["pts3d"][0] for p in preds]); ref_conf = torch.stack([p["conf"][0, ..., 0] for p in preds])
print("host focal", isv.focal_per_group(ref_pts, ref_conf))
print("gpu  focal", isv.gpu_focal_per_group(ops, ref_pts.reshape(G, H * W, 3).cuda().contiguous(), ref_conf.reshape(G, H * W).cuda().contiguous(), H, W))
