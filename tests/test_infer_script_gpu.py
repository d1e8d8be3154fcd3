"""GPU: scripts/evaluation/infer_geo4d.py end to end (the reference's evaluation entry, infer_geo4d.py:314-647) on a
seeded synthetic sequence with seeded synthetic weights: two sliding windows, 2 DDIM steps, alignment, the repo's depth
and pose metrics, and every result file the reference writes (names as in run_evaluation)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_infer_geo4d_script_writes_metrics_and_files(cuda_device, tmp_path):
    out = str(tmp_path / "res")
    cmd = [sys.executable, os.path.join(REPO, "scripts", "evaluation", "infer_geo4d.py"), "--synthetic_weights",
           "--dataset", "synthetic:1:24", "--height", "128", "--width", "192", "--ddim_steps", "2", "--ddim_eta", "0.0",
           "--stride", "8", "--timestep_spacing", "uniform_trailing", "--savedir", out, "--seed", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    run_dirs = glob.glob(os.path.join(out, "synthetic*"))
    assert len(run_dirs) == 1
    d = run_dirs[0]
    seq = os.path.join(d, "synthetic_00")
    for f in ("pred_traj.txt", "pred_focal.txt", "pred_intrinsics.txt", "frame_0000.npy", "frame_0023.npy",
              "frame_colordepth_0000.png", "colored_depth_maps.gif", "conf_0.npy", "init_conf_0.npy", "frame_0000.png",
              "synthetic_00_error_0.png", "_error_log_depth.txt", "_error_log.txt"):
        assert os.path.exists(os.path.join(seq, f)), f
    assert os.path.exists(os.path.join(d, "synthetic_00_eval_metric.txt"))
    assert os.path.exists(os.path.join(d, "_error_log_all.txt")) and os.path.exists(os.path.join(d, "time_cost.txt"))
    traj = np.loadtxt(os.path.join(seq, "pred_traj.txt"))
    assert traj.shape == (24, 8) and np.isfinite(traj).all()
    assert abs(np.linalg.norm(traj[:, 4:8], axis=1) - 1).max() < 1e-5           # unit quaternions (wxyz)
    assert np.load(os.path.join(seq, "frame_0005.npy")).shape == (128, 192)
    assert "time_for_each_frames" in open(os.path.join(d, "time_cost.txt")).read()
    assert "Abs Rel" in open(os.path.join(seq, "_error_log_depth.txt")).read()
