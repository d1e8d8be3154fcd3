"""fp32 CPU restatement of the reference AutoencoderKL (test oracle).

Follows lvdm/modules/networks/ae_modules.py (AttnBlock:53-78, Downsample:100-109,
Upsample:123-127, ResnetBlock:228-248, Encoder.forward:537-580,
Decoder.forward:661-702), lvdm/models/autoencoder.py (encode:129-134,
decode:136-139, decode_with_conf_adaptor:120-127),
lvdm/models/autoencoder_adaptor.py (VAEDecoderadaptor.forward:277-317) and
lvdm/distributions.py:24-40 of jzr99/Geo4D.  State-dict keys are the ones
`AutoencoderKL.state_dict()` produces (encoder.*, decoder.*, quant_conv.*,
post_quant_conv.*, decoder_adaptor.*).  The encoder adaptor is unused at
inference and is not restated.

Pinned by oracle/gen_golden.py against the imported reference Encoder/Decoder/
VAEDecoderadaptor modules.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    # configs/inference_geo4d.yaml:91-118 (ddconfig) and :119-130 (adaptorconfig)
    ch: int = 128
    ch_mult: Sequence[int] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    in_channels: int = 3
    out_ch: int = 3
    z_channels: int = 4
    embed_dim: int = 4
    adaptor_ch: int = 128
    adaptor_res_blocks: int = 1
    adaptor_out_ch: int = 1

    @staticmethod
    def tiny(**kw) -> "VAEConfig":
        base = dict(ch=64, adaptor_ch=64)
        base.update(kw)
        return VAEConfig(**base)


def _res_shapes(p, cin, cout):
    s = OrderedDict()
    s[f"{p}.norm1.weight"] = (cin,)
    s[f"{p}.norm1.bias"] = (cin,)
    s[f"{p}.conv1.weight"] = (cout, cin, 3, 3)
    s[f"{p}.conv1.bias"] = (cout,)
    s[f"{p}.norm2.weight"] = (cout,)
    s[f"{p}.norm2.bias"] = (cout,)
    s[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
    s[f"{p}.conv2.bias"] = (cout,)
    if cin != cout:
        s[f"{p}.nin_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{p}.nin_shortcut.bias"] = (cout,)
    return s


def _attn_shapes(p, c):
    s = OrderedDict()
    s[f"{p}.norm.weight"] = (c,)
    s[f"{p}.norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        s[f"{p}.{n}.weight"] = (c, c, 1, 1)
        s[f"{p}.{n}.bias"] = (c,)
    return s


def param_shapes(cfg: VAEConfig, with_encoder: bool = True, with_adaptor: bool = True):
    s = OrderedDict()
    nres = len(cfg.ch_mult)
    if with_encoder:
        s["encoder.conv_in.weight"] = (cfg.ch, cfg.in_channels, 3, 3)
        s["encoder.conv_in.bias"] = (cfg.ch,)
        in_mult = (1,) + tuple(cfg.ch_mult)
        bin_ = cfg.ch
        for lvl in range(nres):
            bin_ = cfg.ch * in_mult[lvl]
            bout = cfg.ch * cfg.ch_mult[lvl]
            for ib in range(cfg.num_res_blocks):
                s.update(_res_shapes(f"encoder.down.{lvl}.block.{ib}", bin_, bout))
                bin_ = bout
            if lvl != nres - 1:
                s[f"encoder.down.{lvl}.downsample.conv.weight"] = (bin_, bin_, 3, 3)
                s[f"encoder.down.{lvl}.downsample.conv.bias"] = (bin_,)
        s.update(_res_shapes("encoder.mid.block_1", bin_, bin_))
        s.update(_attn_shapes("encoder.mid.attn_1", bin_))
        s.update(_res_shapes("encoder.mid.block_2", bin_, bin_))
        s["encoder.norm_out.weight"] = (bin_,)
        s["encoder.norm_out.bias"] = (bin_,)
        s["encoder.conv_out.weight"] = (2 * cfg.z_channels, bin_, 3, 3)
        s["encoder.conv_out.bias"] = (2 * cfg.z_channels,)
    # decoder (ae_modules.py:604-659); keys are ordered mid -> up.0.. -> norm_out
    bin_ = cfg.ch * cfg.ch_mult[-1]
    s["decoder.conv_in.weight"] = (bin_, cfg.z_channels, 3, 3)
    s["decoder.conv_in.bias"] = (bin_,)
    s.update(_res_shapes("decoder.mid.block_1", bin_, bin_))
    s.update(_attn_shapes("decoder.mid.attn_1", bin_))
    s.update(_res_shapes("decoder.mid.block_2", bin_, bin_))
    ups = {}
    for lvl in reversed(range(nres)):
        bout = cfg.ch * cfg.ch_mult[lvl]
        u = OrderedDict()
        for ib in range(cfg.num_res_blocks + 1):
            u.update(_res_shapes(f"decoder.up.{lvl}.block.{ib}", bin_, bout))
            bin_ = bout
        if lvl != 0:
            u[f"decoder.up.{lvl}.upsample.conv.weight"] = (bin_, bin_, 3, 3)
            u[f"decoder.up.{lvl}.upsample.conv.bias"] = (bin_,)
        ups[lvl] = u
    for lvl in range(nres):
        s.update(ups[lvl])
    s["decoder.norm_out.weight"] = (bin_,)
    s["decoder.norm_out.bias"] = (bin_,)
    s["decoder.conv_out.weight"] = (cfg.out_ch, bin_, 3, 3)
    s["decoder.conv_out.bias"] = (cfg.out_ch,)
    if with_encoder:
        s["quant_conv.weight"] = (2 * cfg.embed_dim, 2 * cfg.z_channels, 1, 1)
        s["quant_conv.bias"] = (2 * cfg.embed_dim,)
    s["post_quant_conv.weight"] = (cfg.z_channels, cfg.embed_dim, 1, 1)
    s["post_quant_conv.bias"] = (cfg.z_channels,)
    if with_adaptor:
        a = cfg.adaptor_ch
        for ib in range(cfg.adaptor_res_blocks + 1):
            s.update(_res_shapes(f"decoder_adaptor.up.0.block.{ib}", a, a))
        s["decoder_adaptor.norm_out.weight"] = (a,)
        s["decoder_adaptor.norm_out.bias"] = (a,)
        s["decoder_adaptor.conv_out.weight"] = (cfg.adaptor_out_ch, a, 3, 3)
        s["decoder_adaptor.conv_out.bias"] = (cfg.adaptor_out_ch,)
    return s


# --------------------------------------------------------------------------- functional forward

def _norm(x, sd, p):
    return F.group_norm(x, 32, sd[f"{p}.weight"], sd[f"{p}.bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _resnet(x, sd, p):
    """ResnetBlock.forward ae_modules.py:228-248 with temb=None."""
    h = F.conv2d(_swish(_norm(x, sd, f"{p}.norm1")), sd[f"{p}.conv1.weight"],
                 sd[f"{p}.conv1.bias"], padding=1)
    h = F.conv2d(_swish(_norm(h, sd, f"{p}.norm2")), sd[f"{p}.conv2.weight"],
                 sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.nin_shortcut.weight"], sd[f"{p}.nin_shortcut.bias"])
    return x + h


def _attn(x, sd, p):
    """AttnBlock.forward ae_modules.py:53-78: single head, d = C."""
    h = _norm(x, sd, f"{p}.norm")
    q = F.conv2d(h, sd[f"{p}.q.weight"], sd[f"{p}.q.bias"])
    k = F.conv2d(h, sd[f"{p}.k.weight"], sd[f"{p}.k.bias"])
    v = F.conv2d(h, sd[f"{p}.v.weight"], sd[f"{p}.v.bias"])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** -0.5)
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    h = F.conv2d(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return x + h


@torch.no_grad()
def encoder_forward(cfg: VAEConfig, sd, x):
    """Encoder.forward ae_modules.py:537-580 (position_encoding=None)."""
    nres = len(cfg.ch_mult)
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for lvl in range(nres):
        for ib in range(cfg.num_res_blocks):
            h = _resnet(h, sd, f"encoder.down.{lvl}.block.{ib}")
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1))  # asymmetric pad, ae_modules.py:102-104
            h = F.conv2d(h, sd[f"encoder.down.{lvl}.downsample.conv.weight"],
                         sd[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = _resnet(h, sd, "encoder.mid.block_1")
    h = _attn(h, sd, "encoder.mid.attn_1")
    h = _resnet(h, sd, "encoder.mid.block_2")
    h = _swish(_norm(h, sd, "encoder.norm_out"))
    return F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)


@torch.no_grad()
def decoder_forward(cfg: VAEConfig, sd, z, give_pre_and_end=False):
    """Decoder.forward ae_modules.py:661-702."""
    nres = len(cfg.ch_mult)
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _resnet(h, sd, "decoder.mid.block_1")
    h = _attn(h, sd, "decoder.mid.attn_1")
    h = _resnet(h, sd, "decoder.mid.block_2")
    for lvl in reversed(range(nres)):
        for ib in range(cfg.num_res_blocks + 1):
            h = _resnet(h, sd, f"decoder.up.{lvl}.block.{ib}")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up.{lvl}.upsample.conv.weight"],
                         sd[f"decoder.up.{lvl}.upsample.conv.bias"], padding=1)
    pre = h
    h = _swish(_norm(h, sd, "decoder.norm_out"))
    h = F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
    return (h, pre) if give_pre_and_end else h


@torch.no_grad()
def adaptor_forward(cfg: VAEConfig, sd, pre):
    """VAEDecoderadaptor.forward autoencoder_adaptor.py:277-317."""
    h = pre
    for ib in range(cfg.adaptor_res_blocks + 1):
        h = _resnet(h, sd, f"decoder_adaptor.up.0.block.{ib}")
    h = _swish(_norm(h, sd, "decoder_adaptor.norm_out"))
    return F.conv2d(h, sd["decoder_adaptor.conv_out.weight"], sd["decoder_adaptor.conv_out.bias"],
                    padding=1)


@torch.no_grad()
def encode_moments(cfg: VAEConfig, sd, x):
    """AutoencoderKL.encode autoencoder.py:129-134 -> posterior parameters."""
    h = encoder_forward(cfg, sd, x)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def posterior_sample(moments, noise):
    """DiagonalGaussianDistribution distributions.py:24-40 (noise supplied)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise


@torch.no_grad()
def decode(cfg: VAEConfig, sd, z):
    """AutoencoderKL.decode autoencoder.py:136-139."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(cfg, sd, z)


@torch.no_grad()
def decode_with_conf_adaptor(cfg: VAEConfig, sd, z):
    """AutoencoderKL.decode_with_conf_adaptor autoencoder.py:120-127."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    rgb, pre = decoder_forward(cfg, sd, z, give_pre_and_end=True)
    conf = adaptor_forward(cfg, sd, pre)
    return torch.cat([rgb, conf], dim=1)
