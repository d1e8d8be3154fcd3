#!/usr/bin/env python
"""Pin geo4d_b200.metrics.depth_evaluation against the reference's own function (test infrastructure).

Run in the BUILD container only (needs /root/reference, read-only):

    python oracle/gen_golden_metrics.py

Imports dust3r/depth_eval.py UNMODIFIED (evo / matplotlib / ... shimmed exactly as oracle/gen_golden.py does for
the alignment pin; none of the shimmed packages is touched by depth_evaluation), runs it on seeded synthetic
depth pairs in every alignment mode the Geo4D scripts use (median, lstsq, LAD-Adam with and without align mask,
IRLS scale, disparity input) and stores the REFERENCE's metric dictionaries + error-map checksums in
tests/golden/metrics_ref.json.  tests/test_metrics_cpu.py replays the same cases through geo4d_b200.metrics.

The pose metrics (ATE / RPE) come from `evo`, which is not vendored in the reference and not installed here:
they stay "parity unpinned" (analytic known-answer tests only).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GEO4D_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)


def cases():
    """name -> (pred [T,H,W], gt [T,H,W], kwargs, align_mask or None)"""
    g = torch.Generator().manual_seed(11)
    T, H, W = 3, 24, 40
    gt = 1.0 + 20.0 * torch.rand(T, H, W, generator=g)
    gt[:, :2] = 0.0                          # invalid ground truth rows
    gt[0, 5, 5] = 95.0                       # beyond max_depth
    pred = (gt / 2.7 - 0.05 + 0.02 * torch.randn(T, H, W, generator=g)).clamp(min=1e-3)
    pred[1, 10:12] *= 3.0                    # outliers
    am = torch.rand(T, H, W, generator=g) > 0.2
    out = {
        "median": (pred, gt, dict(max_depth=70), None),
        "lstsq": (pred, gt, dict(max_depth=70, align_with_lstsq=True), None),
        "lad2": (pred, gt, dict(max_depth=70, align_with_lad2=True, lr=1e-2, max_iters=400, post_clip_max=70), None),
        "lad2_mask": (pred, gt, dict(max_depth=70, align_with_lad2=True, lr=1e-2, max_iters=400, post_clip_max=70), am),
        "scale": (pred, gt, dict(max_depth=None, align_with_scale=True), None),
        "disp": (1.0 / pred, gt, dict(max_depth=70, align_with_lad2=True, lr=1e-2, max_iters=300, disp_input=True), None),
        "kitti": (pred, gt, dict(max_depth=None, align_with_lad2=True), None),   # infer_geo4d.py:536 defaults
    }
    return out


def main():
    sys.path.insert(0, REF)
    from oracle.gen_golden import install_shims, install_alignment_shims
    install_shims()
    install_alignment_shims()
    from dust3r.depth_eval import depth_evaluation as ref_eval
    from geo4d_b200 import metrics
    report = {}
    for name, (pred, gt, kw, am) in cases().items():
        kw_ref = dict(kw)
        if am is not None:   # the reference indexes the align mask with the 2-D validity mask (depth_eval.py:192)
            kw_ref["align_mask"] = am.view(-1, am.shape[-1])
        res, err, full, gtf = ref_eval(pred.clone(), gt.clone(), **kw_ref)
        mine, err2, full2, gtf2 = metrics.depth_evaluation(pred.clone(), gt.clone(), align_mask=am, **kw)
        for k, v in res.items():
            assert abs(float(v) - float(mine[k])) <= 2e-4 * max(1.0, abs(float(v))), (name, k, v, mine[k])
        assert torch.allclose(err, err2, rtol=2e-3, atol=2e-4), name
        report[name] = {"metrics": {k: float(v) for k, v in res.items()}, "err_sum": float(err.double().sum()),
                        "pred_sum": float(full.double().sum()), "gt_sum": float(gtf.double().sum())}
        print(name, report[name]["metrics"])
    path = os.path.join(REPO, "tests", "golden", "metrics_ref.json")
    with open(path, "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
