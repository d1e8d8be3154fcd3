"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/geo4d_b200.h declares; the product refuses to run without a device (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "geo4d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(geo4d_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from geo4d_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/geo4d_b200.h but not exported"


def test_abi_version_and_error_string(lib):
    assert lib.geo4d_abi_version() == 3
    lib.geo4d_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.geo4d_last_error(), bytes)


def test_bad_arguments_fail_loudly(lib):
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    assert lib.geo4d_tap_gemm(None, None) < 0
    lib.geo4d_last_error.restype = ctypes.c_char_p
    assert b"null" in lib.geo4d_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_refuses_without_gpu():
    from geo4d_b200 import _cabi
    with pytest.raises(_cabi.Geo4DError):
        _cabi.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "geo4d_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"


def test_unet_host_keys_match_reference(golden_dir):
    import json
    from geo4d_b200.unet import UNetModel
    ref = json.load(open(os.path.join(golden_dir, "unet_full_keys.json")))
    with torch.device("meta"):
        net = UNetModel(in_channels=20, out_channels=16, model_channels=320, attention_resolutions=[4, 2, 1],
                        num_res_blocks=2, channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64,
                        transformer_depth=1, context_dim=1024, use_linear=True, use_checkpoint=True,
                        temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
                        use_relative_position=False, use_causal_attention=False, temporal_length=16,
                        addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True)
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert set(mine.keys()) == set(ref.keys())
    assert mine == ref


def test_round2_entry_points_validate_arguments_before_touching_cuda(lib):
    """geo4d_align_loop / geo4d_cross_attention2 / geo4d_transform_points / geo4d_sqpnp_from_moments_batch: bad
    arguments come back as negative status codes with a message (no CUDA call is made, so this runs without a GPU)."""
    import numpy as np
    lib.geo4d_last_error.restype = ctypes.c_char_p
    assert lib.geo4d_align_loop(None, None) < 0 and b"null" in lib.geo4d_last_error()
    from geo4d_b200._cabi import AlignLoopDesc
    d = AlignLoopDesc()                      # all pointers null
    assert lib.geo4d_align_loop(ctypes.byref(d), None) < 0
    assert lib.geo4d_transform_points(None, 1, ctypes.c_int64(8), None, None, 0, None) < 0
    assert b"transform_points" in lib.geo4d_last_error()
    assert lib.geo4d_cross_attention2(None, ctypes.c_int64(0), None, None, ctypes.c_int64(0), 77, 16, None, None,
                                      ctypes.c_int64(0), 16, 1, None, ctypes.c_int64(0), 1, 1, 128, ctypes.c_float(0.125), None) < 0
    lib.geo4d_align_loop_record_doubles.restype = ctypes.c_int
    assert lib.geo4d_align_loop_record_doubles(11, 8) == 11 * 12 + 8 * 14 + 3
    lib.geo4d_align_loop_part_floats.restype = ctypes.c_size_t
    assert lib.geo4d_align_loop_part_floats(16, 18) == 16 * 18 * 128
    lib.geo4d_lad_fit_workspace_doubles.restype = ctypes.c_size_t
    assert lib.geo4d_lad_fit_workspace_doubles(3) >= 3 * 4
    # the host SQPnP batch solver needs no GPU at all: an exact 4-point pin-hole problem per entry
    f = 100.0
    pts = np.array([[0.3, -0.2, 3.0], [-0.5, 0.4, 2.5], [0.1, 0.6, 4.0], [0.7, 0.2, 3.5], [-0.2, -0.6, 2.8]])
    u, v = f * pts[:, 0] / pts[:, 2], f * pts[:, 1] / pts[:, 2]
    r2 = u * u + v * v
    mom = np.zeros(41)
    mom[0:4] = (len(pts), u.sum(), v.sum(), r2.sum())
    mom[4:7] = pts.sum(0); mom[7:10] = (u[:, None] * pts).sum(0); mom[10:13] = (v[:, None] * pts).sum(0)
    mom[13:16] = (r2[:, None] * pts).sum(0)
    mm = np.stack([pts[:, 0] ** 2, pts[:, 0] * pts[:, 1], pts[:, 0] * pts[:, 2], pts[:, 1] ** 2, pts[:, 1] * pts[:, 2], pts[:, 2] ** 2], 1)
    mom[16:22] = mm.sum(0); mom[22:28] = (u[:, None] * mm).sum(0); mom[28:34] = (v[:, None] * mm).sum(0)
    mom[34:40] = (r2[:, None] * mm).sum(0); mom[40] = len(pts)
    B = 5
    moms = np.ascontiguousarray(np.tile(mom, (B, 1)))
    fs = np.full(B, f)
    R, t, ok = np.empty((B, 9)), np.empty((B, 3)), np.zeros(B, dtype=np.int32)
    lib.geo4d_sqpnp_from_moments_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert lib.geo4d_sqpnp_from_moments_batch(moms.ctypes.data, fs.ctypes.data, B, R.ctypes.data, t.ctypes.data,
                                              ok.ctypes.data, 3) == 1
    assert ok.all() and np.allclose(R, np.eye(3).reshape(9), atol=1e-7) and np.allclose(t, 0, atol=1e-7)
