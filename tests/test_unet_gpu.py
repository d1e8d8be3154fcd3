"""GPU: the CUDA U-Net (geo4d_b200.UNetModel, through the C ABI) against the reference output stored in
tests/golden/unet_tiny.pt and against the CPU oracle on fresh seeded inputs.
Tolerance: one U-Net forward, bf16 tensor-core path vs fp32 reference: rel-L2 <= 2e-2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_UNET_FWD = 2e-2


def rel_l2(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12))


def make_unet(cfg_kw, sd, device):
    from geo4d_b200.unet import UNetModel
    net = UNetModel(in_channels=20, out_channels=16, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                    channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64, transformer_depth=1,
                    use_linear=True, use_checkpoint=True, temporal_conv=True, temporal_attention=True,
                    temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False,
                    addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True,
                    **cfg_kw)
    net.load_state_dict(sd, strict=True)
    return net.to(device).prepare()


def test_unet_tiny_vs_reference_golden(cuda_device, golden_dir):
    from oracle import unet as ou
    g = torch.load(os.path.join(golden_dir, "unet_tiny.pt"))
    cfg = ou.UNetConfig.tiny(**g["cfg"])
    sd = ou.init_params(ou.param_shapes(cfg), seed=g["seed"])
    net = make_unet(g["cfg"], sd, cuda_device)
    y = net(g["x"].to(cuda_device), g["timesteps"].to(cuda_device), context=g["context"].to(cuda_device),
            fs=g["fs"].to(cuda_device))
    torch.cuda.synchronize()
    assert y.shape == g["y"].shape
    assert rel_l2(y, g["y"]) < TOL_UNET_FWD


def test_unet_tiny_layerwise_vs_oracle(cuda_device):
    """Block-by-block comparison (localises a numerics regression to one block)."""
    from oracle import unet as ou
    from geo4d_b200 import ops
    cfg = ou.UNetConfig.tiny()
    sd = ou.init_params(ou.param_shapes(cfg), seed=7)
    net = make_unet(dict(model_channels=64, context_dim=64, temporal_length=4), sd, cuda_device)
    g = torch.Generator().manual_seed(9)
    b, t, hh, ww = 2, 4, 8, 16
    x = torch.randn(b, 20, t, hh, ww, generator=g)
    ctx = torch.randn(b, 77 + 16 * t, 64, generator=g)
    ts = torch.tensor([999, 19])
    taps_o = {}
    y_o = ou.forward(cfg, sd, x, ts, ctx, None, taps=taps_o)
    xd, ctxd = x.to(cuda_device), ctx.to(cuda_device)
    net.set_context(ctxd, t)
    emb_all = net.embed(ts.to(cuda_device), None, b)
    taps_g = {}
    rows = ops.bcthw_to_rows(xd.contiguous(), None, 64)
    y_rows = net.forward_rows(rows, emb_all, (b, t, hh, ww), taps=taps_g)
    y_g = ops.rows_to_bcthw(y_rows, 16, b, t, hh, ww)
    torch.cuda.synchronize()
    worst = 0.0
    for k, (hr, geom) in taps_g.items():
        bb, tt, h2, w2 = geom
        ref = taps_o[k].permute(0, 2, 3, 1).reshape(bb * tt * h2 * w2, -1)
        e = rel_l2(hr, ref)
        worst = max(worst, e)
        assert e < TOL_UNET_FWD, f"{k}: rel-L2 {e}"
    assert rel_l2(y_g, y_o) < TOL_UNET_FWD
