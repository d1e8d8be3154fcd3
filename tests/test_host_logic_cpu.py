"""CPU: host-side logic of the product that needs no GPU -- schedule tables (bit-identical to the oracle, which is
pinned to the reference), window slicing, config aliasing, GEGLU interleave, M-tile picker."""
import json
import os

import numpy as np
import torch

from geo4d_b200 import schedule as ps
from oracle import ddim as od


def test_schedule_tables_bit_identical_to_oracle():
    bufs = ps.register_schedule_buffers(1000, 0.00085, 0.012, "linear", True)
    sch = od.Schedule.geo4d()
    assert np.array_equal(bufs["betas"], sch.betas)
    assert np.array_equal(bufs["alphas_cumprod"], sch.alphas_cumprod)
    assert np.array_equal(bufs["sqrt_alphas_cumprod"], sch.sqrt_alphas_cumprod)
    assert np.array_equal(bufs["sqrt_one_minus_alphas_cumprod"], sch.sqrt_one_minus_alphas_cumprod)
    assert np.array_equal(ps.make_scale_arr(), sch.scale_arr)
    for S in (2, 5, 50):
        mine = ps.DDIMTables(bufs["alphas_cumprod"], ps.make_scale_arr(), S, "uniform_trailing", 0.0)
        ref = od.make_ddim_tables(sch, S)
        assert np.array_equal(mine.timesteps, ref.timesteps)
        assert np.array_equal(mine.alphas.astype(np.float32), ref.alphas)
        assert np.array_equal(np.asarray(mine.alphas_prev, np.float32), ref.alphas_prev)
        assert np.array_equal(mine.scale, ref.scale) and np.array_equal(mine.scale_prev, ref.scale_prev)


def test_step_coefficients_reproduce_oracle_update():
    """geo4d_ddim_step's {sa, s1, rescale, sqrt(a_prev), dir, sigma} rows give the oracle's p_sample_ddim."""
    bufs = ps.register_schedule_buffers(1000, 0.00085, 0.012, "linear", True)
    sch = od.Schedule.geo4d()
    S = 5
    tab = ps.DDIMTables(bufs["alphas_cumprod"], ps.make_scale_arr(), S, "uniform_trailing", 0.0)
    coef = tab.step_coefficients(bufs["sqrt_alphas_cumprod"], bufs["sqrt_one_minus_alphas_cumprod"])
    otab = od.make_ddim_tables(sch, S)
    g = torch.Generator().manual_seed(0)
    x, v = torch.randn(64, generator=g), torch.randn(64, generator=g)
    for i in range(S):
        index = S - 1 - i
        ref, ref_x0 = od.ddim_step_v(x, v, sch, otab, index)
        sa, s1, rs, sap, dr, sg = [torch.tensor(float(c)) for c in coef[i]]
        e_t = sa * v + s1 * x
        x0 = (sa * x - s1 * v) * rs
        assert torch.allclose(x0, ref_x0, rtol=0, atol=1e-6)
        assert torch.allclose(sap * x0 + dr * e_t, ref, rtol=0, atol=1e-6)
    # first step: abar_999 = 0  =>  pred_x0 = -v * rescale, e_t = x
    assert coef[0][0] == 0.0 and coef[0][1] == 1.0


def test_sliding_windows_match_reference_rule():
    from geo4d_b200.pipeline import sliding_windows
    assert [(s.start, s.stop) for s in sliding_windows(50, 8)] == [(0, 16), (8, 24), (16, 32), (24, 40), (32, 48), (34, 50)]
    assert [(s.start, s.stop) for s in sliding_windows(16, 8)] == [(0, 16)]
    assert len(sliding_windows(264, 8)) == 32


def test_config_aliases_resolve_reference_targets():
    from geo4d_b200.config import get_obj_from_str, load_yaml
    from geo4d_b200.unet import UNetModel
    from geo4d_b200.vae import AutoencoderKL
    assert get_obj_from_str("lvdm.modules.networks.openaimodel3d.UNetModel") is UNetModel
    assert get_obj_from_str("lvdm.models.autoencoder.AutoencoderKL") is AutoencoderKL
    cfg = load_yaml(os.path.join(os.path.dirname(os.path.dirname(__file__)), "configs", "inference_geo4d.yaml"))
    assert cfg["model"]["params"]["unet_config"]["params"]["model_channels"] == 320
    assert cfg["postprocess"]["n_iter"] == 500


def test_config_equals_the_references_parsed_yaml_key_by_key(golden_dir):
    """configs/inference_geo4d.yaml vs tests/golden/config_ref.json (= yaml.safe_load of the reference's
    configs/inference_geo4d.yaml, written by the build container): every section, every key, every value."""
    import json
    from geo4d_b200.config import load_yaml
    mine = load_yaml(os.path.join(os.path.dirname(os.path.dirname(__file__)), "configs", "inference_geo4d.yaml"))
    ref = json.load(open(os.path.join(golden_dir, "config_ref.json")))

    def walk(a, b, path):
        assert type(a) is type(b) or (isinstance(a, (int, float)) and isinstance(b, (int, float))), path
        if isinstance(a, dict):
            assert sorted(a.keys()) == sorted(b.keys()), (path, sorted(a.keys()), sorted(b.keys()))
            for k in a:
                walk(a[k], b[k], path + "." + str(k))
        elif isinstance(a, list):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                walk(x, y, f"{path}[{i}]")
        else:
            assert a == b, (path, a, b)
    walk(mine, ref, "cfg")


def test_geglu_interleave_and_tile_picker():
    from geo4d_b200.unet import _interleave32
    from geo4d_b200.ops import pick_box
    a = torch.arange(64).float()[:, None]
    b = 100 + torch.arange(64).float()[:, None]
    w = _interleave32(a, b)[:, 0]
    assert w[:32].tolist() == list(range(32)) and w[32:64].tolist() == [100 + i for i in range(32)]
    assert w[64:96].tolist() == list(range(32, 64))
    for (W, H, N) in [(64, 40, 16), (32, 20, 16), (16, 10, 16), (8, 5, 16), (512, 320, 4), (2560, 16, 1)]:
        bw, bh, bn = pick_box(W, H, N)
        assert bw * bh * bn <= 128 and bw <= W and bh <= H and bn <= N
    assert pick_box(64, 40, 16) == (64, 2, 1) and pick_box(8, 5, 16) == (8, 1, 16)


def test_full_model_state_dict_layout(golden_dir):
    """LatentVisualDiffusion exposes model.diffusion_model.* / first_stage_model.* + the schedule buffers."""
    from geo4d_b200.config import instantiate_from_config, load_yaml
    cfg = load_yaml(os.path.join(os.path.dirname(os.path.dirname(__file__)), "configs", "inference_geo4d.yaml"))
    with torch.device("meta"):
        m = instantiate_from_config(cfg["model"])
    keys = set(m.state_dict().keys())
    unet = json.load(open(os.path.join(golden_dir, "unet_full_keys.json")))
    assert all(("model.diffusion_model." + k) in keys for k in unet)
    for b in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "scale_arr",
              "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_mean_coef1"):
        assert b in keys
    assert m.scale_arr.shape[0] == 1400 and m.num_timesteps == 1000 and m.parameterization == "v"


def test_native_sqpnp_matches_numpy():
    """geo4d_sqpnp_from_moments (C++ host solver in the library) vs the NumPy statement of the same algorithm on
    random poses / focals / noise levels / mask densities, including wrong tentative focals: same solution."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    from geo4d_b200 import init_solvers as isv
    H, W = 48, 64
    v, u = np.mgrid[:H, :W]
    rng = np.random.default_rng(0)
    compared = 0
    for trial in range(60):
        f = rng.uniform(40, 160)
        z = rng.uniform(1.5, 8, (H, W))
        pts = np.stack([(u - W / 2) * z / f, (v - H / 2) * z / f, z], -1) + rng.normal(0, rng.choice([0.0, 0.01, 0.1]), (H, W, 3))
        Rw = Rotation.from_euler("xyz", rng.uniform(-1, 1, 3)).as_matrix()
        tw = rng.uniform(-1, 1, 3)
        world = (pts - tw) @ Rw                      # camera = Rw world + tw
        mask = rng.random((H, W)) < rng.choice([1.0, 0.5, 0.05])
        mom = isv.moments_numpy(world, mask, W / 2, H / 2)
        ft = f * rng.choice([1.0, 0.97, 1.03, 1.3, 0.7])
        a = isv.sqpnp_from_moments(mom, ft)
        b = isv.sqpnp_from_moments_native(mom, ft)
        assert (a is None) == (b is None)
        if a is None:
            continue
        assert np.abs(a[0] - b[0]).max() < 1e-9 and np.abs(a[1] - b[1]).max() < 1e-9 * max(1.0, np.abs(a[1]).max())
        compared += 1
    assert compared > 40
    assert isv.sqpnp_from_moments_native(np.zeros(41), 100.0) is None      # fewer than 4 points


def test_native_sqpnp_matches_cv2_sqpnp():
    """Pin against the library the reference actually calls: cv2.solvePnP(flags=SOLVEPNP_SQPNP) on all points (the
    model-fitting step inside solvePnPRansac, init_im_poses.py:845-848) vs geo4d_sqpnp_from_moments fed with the 41
    moments of the same correspondences -- including slightly wrong tentative focals, as fast_pnp tries them."""
    import cv2
    import numpy as np
    from scipy.spatial.transform import Rotation
    from geo4d_b200 import init_solvers as isv
    H, W = 48, 64
    v, u = np.mgrid[:H, :W]
    rng = np.random.default_rng(1)
    for trial in range(12):
        f = rng.uniform(40, 160)
        z = rng.uniform(1.5, 8, (H, W))
        pts = np.stack([(u - W / 2) * z / f, (v - H / 2) * z / f, z], -1) + rng.normal(0, 0.02, (H, W, 3))
        Rw = Rotation.from_euler("xyz", rng.uniform(-1, 1, 3)).as_matrix()
        tw = rng.uniform(-1, 1, 3)
        world = (pts - tw) @ Rw
        mask = np.ones((H, W), bool)
        ft = f * rng.choice([1.0, 0.97, 1.03])
        got = isv.sqpnp_from_moments_native(isv.moments_numpy(world, mask, W / 2, H / 2), ft)
        K = np.float64([(ft, 0, W / 2), (0, ft, H / 2), (0, 0, 1)])
        pix = np.stack([u, v], -1).astype(np.float64)
        ok, rv, tv = cv2.solvePnP(world[mask].astype(np.float64), pix[mask], K, None, flags=cv2.SOLVEPNP_SQPNP)
        assert ok and got is not None
        assert np.abs(cv2.Rodrigues(rv)[0] - got[0]).max() < 1e-7 and np.abs(tv.ravel() - got[1]).max() < 1e-7


def test_infer_script_has_the_references_flag_surface(golden_dir):
    """scripts/evaluation/infer_geo4d.py: every option of the reference's get_parser (infer_geo4d.py:688-718, extracted
    by AST into tests/golden/infer_args_ref.json) exists with the same type / action / default."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(__file__))
    spec = importlib.util.spec_from_file_location("infer_geo4d_b200", os.path.join(root, "scripts", "evaluation", "infer_geo4d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mine = {a.option_strings[0]: a for a in mod.get_parser()._actions if a.option_strings}
    ref = json.load(open(os.path.join(golden_dir, "infer_args_ref.json")))
    assert len(ref) >= 25
    for name, kw in ref.items():
        assert name in mine, name
        a = mine[name]
        if kw.get("action") == "store_true":
            assert a.nargs == 0 and a.const is True and a.default is False, name
        else:
            assert (a.type.__name__ if a.type else "str") == kw.get("type", "str"), name
            if name != "--config":        # the reference has no default config; this repo ships one
                assert a.default == kw.get("default"), (name, a.default, kw.get("default"))


def test_minimal_sample_moments_match_the_dense_moment_routine():
    """init_solvers._minimal_sample_moments (4-point RANSAC hypotheses of the windowed PnP) vs moments_numpy on the
    same four pixels; a sample with fewer than four valid draws is flagged unusable."""
    import numpy as np
    from geo4d_b200 import init_solvers as isv
    rng = np.random.default_rng(0)
    H, W = 24, 32
    pts = rng.standard_normal((H, W, 3)).astype(np.float32)
    cand = rng.integers(0, H * W, size=(3, 16))
    conf = (rng.random((3, 16)) > 0.3).astype(np.float32)
    conf[2, :] = 0.0
    conf[2, :3] = 1.0                                   # only three valid draws
    mom, usable = isv._minimal_sample_moments(pts.reshape(-1, 3)[cand], conf, cand, W, W / 2, H / 2)
    assert list(usable) == [True, True, False]
    for i in range(2):
        sel = [int(c) for c, v in zip(cand[i], conf[i] > 0.5) if v][:4]
        ref = np.zeros(41)
        for p in sel:
            mk = np.zeros(H * W, dtype=bool)
            mk[p] = True
            ref += isv.moments_numpy(pts, mk.reshape(H, W), W / 2, H / 2)
        assert np.allclose(mom[i], ref, atol=1e-9)
    # a minimal sample of an exact pin-hole scene reproduces the camera (SQPnP on 4 points)
    f = 0.9 * W
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = 2 + 0.5 * np.sin(xs / 5.0) + 0.3 * np.cos(ys / 3.0)
    cam = np.stack([(xs - W / 2) / f * depth, (ys - H / 2) / f * depth, depth], -1)
    pix = np.array([[3 * W + 4, 20 * W + 29, 11 * W + 15, 7 * W + 25] + [0] * 12])
    m4, ok = isv._minimal_sample_moments(cam.reshape(-1, 3)[pix], np.ones((1, 16), dtype=np.float32), pix, W, W / 2, H / 2)
    sol = isv.sqpnp_from_moments(m4[0], f)
    assert ok[0] and sol is not None
    assert np.allclose(sol[0], np.eye(3), atol=1e-6) and np.allclose(sol[1], 0, atol=1e-6)
