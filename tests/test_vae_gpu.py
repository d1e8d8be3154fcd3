"""GPU: CUDA AutoencoderKL (decode, decode+confidence head, encode) vs the reference outputs stored in
tests/golden/vae_tiny.pt.  Tolerance: bf16 tensor-core path (about 70 bf16-rounded layers deep) vs the fp32 reference:
rel-L2 <= 3e-2 on the decoded maps / moments, and the SURVEY.md 8(c) bar on decoded maps,
per-pixel mean-L1 <= 2e-2 (absolute; the maps live in ~[-2, 2])."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 3e-2
TOL_L1_ABS = 2e-2


def rel_l2(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12))


def make_vae(cfg_kw, sd, device):
    from geo4d_b200.vae import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=cfg_kw["ch"],
              ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    ad = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=1, ch=cfg_kw["adaptor_ch"],
              ch_mult=[1], num_res_blocks=1, attn_resolutions=[], dropout=0.0)
    vae = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4, adaptorconfig=ad)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder_adaptor.") for k in missing)
    return vae.to(device).prepare()


def test_vae_tiny_vs_reference_golden(cuda_device, golden_dir):
    from oracle import unet as ou
    from oracle import vae as ov
    g = torch.load(os.path.join(golden_dir, "vae_tiny.pt"))
    cfg = ov.VAEConfig.tiny(**g["cfg"])
    sd = ou.init_params(ov.param_shapes(cfg), seed=g["seed"])
    vae = make_vae(g["cfg"], sd, cuda_device)
    z = g["z"].to(cuda_device)
    dec = vae.decode(z)
    assert dec.shape == g["dec"].shape
    assert rel_l2(dec, g["dec"]) < TOL
    dc = vae.decode_with_conf_adaptor(z)
    assert dc.shape == g["dec_conf"].shape
    assert rel_l2(dc, g["dec_conf"]) < TOL
    l1 = float((dc.cpu() - g["dec_conf"]).abs().mean())
    assert l1 < TOL_L1_ABS
    mom = vae.encode_moments(g["img"].to(cuda_device))
    assert mom.shape == g["moments"].shape
    assert rel_l2(mom, g["moments"]) < TOL
    post = vae.encode(g["img"].to(cuda_device))
    noise = torch.randn(post.mean.shape, generator=torch.Generator().manual_seed(0))
    zs = post.sample(noise)
    assert rel_l2(zs, ov.posterior_sample(g["moments"], noise)) < TOL
