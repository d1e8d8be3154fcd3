#!/usr/bin/env python
"""Fused small-parameter kernel vs autograd path: parameter differences after n iterations (debug aid)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from oracle import align as oa
from geo4d_b200.cloud_opt import LightPointCloudGroupOptimizer

dev = torch.device("cuda")
groups, preds, _ = oa.synthetic_scene(T=24, H=32, W=48, noise=0.003)
views = [[{"idx": (i,)} for i in g] for g in groups]
names = ["im_poses", "im_focals", "pw_poses", "s_depth", "t_depth", "traj_align_poses", "im_depthmaps"]
for niter in [1, 2, 3, 10, 20, 21, 22, 25, 40, 60]:
    outs = {}
    for mode in ("1", "0"):
        os.environ["GEO4D_ALIGN_AUTOGRAD"] = mode
        preds_d = [{k: v.to(dev) for k, v in p.items()} for p in preds]
        sc = LightPointCloudGroupOptimizer(views, preds_d, conf="id", conf_optimize=True, verbose=False,
                                           shared_focal=True, num_total_iter=niter, temporal_smoothing_weight=0.015,
                                           translation_weight=1.0, depth_traj_start_iter=20, lad_max_iters=300,
                                           use_cuda_graph=False)
        with torch.enable_grad():
            # fixed lr so that runs of different length follow the same trajectory
            sc.compute_global_alignment(init="group", niter=niter, schedule="linear", lr=0.03, lr_min=0.03)
        outs[mode] = {n: getattr(sc, n).detach().double().cpu().clone() for n in names}
        outs[mode]["valid"] = list(sc.valid_traj_group_list)
        outs[mode]["P"] = sc.get_im_poses().detach().double().cpu()
    a, f = outs["1"], outs["0"]
    line = [f"niter {niter:3d} valid {a['valid']}/{f['valid']}"]
    for n in names + ["P"]:
        d = (a[n] - f[n]).abs()
        line.append(f"{n} {float(d.max()):.2e}")
    print(" | ".join(line), flush=True)
    if niter in (1, 21):
        for n in ["im_poses", "pw_poses", "traj_align_poses", "s_depth", "t_depth", "im_focals"]:
            d = (a[n] - f[n]).abs()
            if float(d.max()) > 1e-5:
                idx = int(d.reshape(-1).argmax())
                print(f"   {n}: worst flat idx {idx} of shape {tuple(d.shape)}; autograd {a[n].reshape(-1)[idx]:.6f} fused {f[n].reshape(-1)[idx]:.6f}")
                print("   per-column max diff:", [f"{float(x):.1e}" for x in d.reshape(-1, d.shape[-1]).max(0).values])
