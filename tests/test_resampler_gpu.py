"""GPU: geo4d_b200.resampler.Resampler (image-token Resampler of the conditioning front-end) through the
C ABI against the outputs of the REFERENCE's own lvdm.modules.encoders.resampler.Resampler stored in
tests/golden/resampler_ref.pt (oracle/gen_golden_resampler.py), same seeded weights; strict state-dict load pins the
checkpoint keys.  Tolerance: 4 transformer layers of bf16 GEMMs / attention vs fp32: rel-L2 <= 2e-2."""
import os
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_resampler_vs_reference_golden(cuda_device, golden_dir):
    from oracle.gen_golden_resampler import seeded_state
    from geo4d_b200.resampler import Resampler
    g = torch.load(os.path.join(golden_dir, "resampler_ref.pt"))
    net = Resampler(**g["kw"])
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s in g["shapes"]]
    sd = seeded_state(g["shapes"], seed=g["seed"])
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).prepare()
    gen = torch.Generator().manual_seed(g["input_seed"])
    x = torch.randn(1, 257, 1280, generator=gen)
    xf = torch.randn(1, 2, 257, 1280, generator=gen)
    assert abs(float(x.double().sum()) - g["x_sum"]) < 1e-6 and abs(float(xf.double().sum()) - g["xf_sum"]) < 1e-6
    y = net(x.to(cuda_device)).cpu()
    ref = g["y"].float()
    e = float((y - ref).norm() / ref.norm())
    assert y.shape == ref.shape and e < 2e-2, e
    # per-frame image tokens (cross_attention=True, infer_geo4d.py:140-149): frame t uses queries [16 t, 16 t + 16)
    nq = g["kw"]["num_queries"]
    net.latents.data = net.latents.data[:, :2 * nq].contiguous()
    net.prepare()
    yf = net(xf.to(cuda_device)).cpu()
    ef = float((yf - g["yf"].float()).norm() / g["yf"].float().norm())
    assert yf.shape == g["yf"].shape and ef < 2e-2, ef
