"""CPU: the repo's own depth / pose metrics (geo4d_b200/metrics.py).

* depth_evaluation is replayed on the seeded cases of oracle/gen_golden_metrics.py and compared with the
  outputs of the REFERENCE's own dust3r.depth_eval.depth_evaluation stored in tests/golden/metrics_ref.json
  (tolerance 2e-4 relative: same fp32 arithmetic, different reduction order).
* ATE / RPE restate evo (un-vendored, parity unpinned): analytic known-answer tests -- a Sim(3)-transformed
  copy of a trajectory has zero error; a known per-pose offset / per-step rotation gives the closed-form RMSE.
"""
import json
import os

import numpy as np
import pytest
import torch


def test_depth_evaluation_vs_reference_golden(golden_dir):
    from geo4d_b200 import metrics
    from oracle.gen_golden_metrics import cases
    ref = json.load(open(os.path.join(golden_dir, "metrics_ref.json")))
    for name, (pred, gt, kw, am) in cases().items():
        res, err, full, gtf = metrics.depth_evaluation(pred.clone(), gt.clone(), align_mask=am, **kw)
        for k, v in ref[name]["metrics"].items():
            assert abs(res[k] - v) <= 2e-4 * max(1.0, abs(v)), (name, k, res[k], v)
        assert abs(float(err.double().sum()) - ref[name]["err_sum"]) <= 2e-3 * abs(ref[name]["err_sum"]), name
        assert abs(float(full.double().sum()) - ref[name]["pred_sum"]) <= 2e-3 * abs(ref[name]["pred_sum"]), name
        assert abs(float(gtf.double().sum()) - ref[name]["gt_sum"]) <= 1e-6 * abs(ref[name]["gt_sum"]), name


def test_depth_evaluation_edge_cases():
    from geo4d_b200 import metrics
    gt = torch.zeros(2, 4, 5)
    pred = torch.ones(2, 4, 5)
    res, err, full, gtf = metrics.depth_evaluation(pred, gt, max_depth=70)   # no valid pixel at all
    assert res["valid_pixels"] == 0 and res["Abs Rel"] == 0 and float(err.sum()) == 0
    gt = torch.full((1, 4, 5), 2.0)
    res, *_ = metrics.depth_evaluation(torch.full((1, 4, 5), 0.5), gt, max_depth=70)   # median scaling is exact
    assert res["Abs Rel"] < 1e-7 and res["δ < 1.25"] == 1.0
    avg = metrics.average_depth_metrics([{"Abs Rel": 0.1, "valid_pixels": 10}, {"Abs Rel": 0.3, "valid_pixels": 30}])
    assert abs(avg["Abs Rel"] - 0.25) < 1e-12


def _random_traj(n, seed):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    P = np.tile(np.eye(4), (n, 1, 1))
    P[:, :3, :3] = Rotation.from_rotvec(0.2 * rng.standard_normal((n, 3))).as_matrix()
    P[:, :3, 3] = np.cumsum(0.1 * rng.standard_normal((n, 3)), 0)
    return P


def _to_tum(P):
    from geo4d_b200.cloud_opt import c2w_to_tumpose
    return [np.stack([c2w_to_tumpose(p) for p in P]), np.arange(len(P)).astype(float)]


def test_ate_rpe_known_answers(tmp_path):
    from scipy.spatial.transform import Rotation
    from geo4d_b200 import metrics
    ref = _random_traj(20, 0)
    # (1) a Sim(3)-transformed copy: zero ATE / RPE after alignment
    Rg = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix()
    est = ref.copy()
    est[:, :3, 3] = 2.5 * (ref[:, :3, 3] @ Rg.T) + np.array([1.0, -2.0, 3.0])
    est[:, :3, :3] = Rg[None] @ ref[:, :3, :3]
    ate, rt, rr = metrics.eval_metrics(_to_tum(est), _to_tum(ref), seq="kat", filename=str(tmp_path / "m.txt"))
    assert ate < 1e-9 and rt < 1e-9 and rr < 1e-5
    assert "rmse" in open(tmp_path / "m.txt").read()
    # (2) Umeyama recovers a known Sim(3)
    R, t, c = metrics.umeyama_alignment(ref[:, :3, 3].T, est[:, :3, 3].T, True)
    assert np.allclose(R, Rg, atol=1e-9) and np.allclose(t, [1.0, -2.0, 3.0], atol=1e-9) and abs(c - 2.5) < 1e-9
    # (3) RPE on a rotation-free reference: an extra rotation of a degrees per step gives a degrees per pair;
    #     an extra drift of d per step along x gives |d| per pair (no alignment in metrics.rpe itself)
    a, d = 3.0, 0.02
    ref0 = ref.copy()
    ref0[:, :3, :3] = np.eye(3)
    est2 = ref0.copy()
    for i in range(len(ref)):
        est2[i, :3, :3] = Rotation.from_euler("z", a * i, degrees=True).as_matrix()
    _, rot = metrics.rpe(ref0, est2, delta=1)
    assert np.allclose(rot, a, atol=1e-6)
    est2 = ref0.copy()
    est2[:, 0, 3] += d * np.arange(len(ref))
    tr, rot = metrics.rpe(ref0, est2, delta=1)
    assert np.allclose(tr, d, atol=1e-12) and np.allclose(rot, 0, atol=1e-6)
    # (4) APE of a pure per-pose offset pattern: +-d alternating along x keeps the Umeyama fit at identity scale ~1
    d = 0.01
    est3 = ref.copy()
    est3[:, 0, 3] += d * np.where(np.arange(len(ref)) % 2 == 0, 1.0, -1.0)
    ape = metrics.ape_translation(ref, est3)
    assert np.allclose(ape, d)


def test_tum_file_round_trip(tmp_path):
    """the TUM file keeps the quaternion in wxyz order exactly as get_tum_poses returns it
    (vo_eval.py:465-473 via base_opt_group.py:390-393)."""
    from geo4d_b200 import metrics
    P = _random_traj(7, 3)
    traj = _to_tum(P)
    path = str(tmp_path / "pred_traj.txt")
    metrics.save_trajectory_tum_format(traj, path)
    back = metrics.load_tum_trajectory(path)
    assert np.allclose(back[0], traj[0]) and np.allclose(back[1], traj[1])
    assert np.allclose(metrics.tum_to_matrices(back[0]), P, atol=1e-12)
    first = open(path).read().split("\n")[0].split(" ")
    assert len(first) == 8 and abs(float(first[4]) - traj[0][0, 3]) < 1e-12   # column 4 = qw
