"""B200-native AutoencoderKL behind the reference's `AutoencoderKL` seam.

Drop-in for lvdm/models/autoencoder.py:AutoencoderKL (constructor arguments =
`first_stage_config.params` / `pointmap_vae_config.params` of
configs/inference_geo4d.yaml:2-36,91-130; `encode`, `decode`,
`decode_with_conf_adaptor`; identical `state_dict()` keys for
encoder.*/decoder.*/quant_conv.*/post_quant_conv.*/decoder_adaptor.* and, so
that the reference vae.ckpt loads with strict=True, placeholders for the
unused encoder_adaptor.* tensors).

All convolutions run as tcgen05 tap-GEMMs on channels-last bf16 rows; the
3x3 taps are TMA-shifted boxes (no im2col) except the three stride-2
encoder convs.  The mid AttnBlock (1 head, d = C) is three batched matmuls on
the same kernel plus a row softmax.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from ._cabi import require_device
from .unet import _Node, _pad_k, _register


class DiagonalGaussianDistribution:
    """lvdm/distributions.py:24-65 (sample / mode only)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = torch.randn(self.mean.shape)  # CPU RNG, like the reference (distributions.py:37)
        return self.mean + self.std * noise.to(device=self.parameters.device)

    def mode(self) -> torch.Tensor:
        return self.mean


def _res_shapes(s, p, cin, cout):
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


def _attn_shapes(s, p, c):
    s[f"{p}.norm.weight"] = (c,)
    s[f"{p}.norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        s[f"{p}.{n}.weight"] = (c, c, 1, 1)
        s[f"{p}.{n}.bias"] = (c,)


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, test=False, logdir=None, input_dim=4, test_args=None,
                 adaptorconfig=None):
        super().__init__()
        dd = dict(ddconfig)
        assert dd["double_z"]
        if dd.get("attn_resolutions"):
            raise NotImplementedError("attn_resolutions other than [] are not used by Geo4D")
        self.ch = dd["ch"]
        self.ch_mult = tuple(dd["ch_mult"])
        self.num_res_blocks = dd["num_res_blocks"]
        self.in_channels = dd["in_channels"]
        self.out_ch = dd["out_ch"]
        self.z_channels = dd["z_channels"]
        self.embed_dim = embed_dim
        self.adaptorconfig = dict(adaptorconfig) if adaptorconfig is not None else None
        self.image_key = image_key
        for k, shp in self._param_shapes().items():
            _register(self, k, torch.empty(shp))
        self._packed = None

    # ------------------------------------------------------------------ parameter inventory
    def _param_shapes(self):
        s = OrderedDict()
        ch, mult, nrb, zc = self.ch, self.ch_mult, self.num_res_blocks, self.z_channels
        nres = len(mult)
        s["encoder.conv_in.weight"] = (ch, self.in_channels, 3, 3)
        s["encoder.conv_in.bias"] = (ch,)
        in_mult = (1,) + tuple(mult)
        bin_ = ch
        for lvl in range(nres):
            bin_ = ch * in_mult[lvl]
            bout = ch * mult[lvl]
            for ib in range(nrb):
                _res_shapes(s, f"encoder.down.{lvl}.block.{ib}", bin_, bout)
                bin_ = bout
            if lvl != nres - 1:
                s[f"encoder.down.{lvl}.downsample.conv.weight"] = (bin_, bin_, 3, 3)
                s[f"encoder.down.{lvl}.downsample.conv.bias"] = (bin_,)
        _res_shapes(s, "encoder.mid.block_1", bin_, bin_)
        _attn_shapes(s, "encoder.mid.attn_1", bin_)
        _res_shapes(s, "encoder.mid.block_2", bin_, bin_)
        s["encoder.norm_out.weight"] = (bin_,)
        s["encoder.norm_out.bias"] = (bin_,)
        s["encoder.conv_out.weight"] = (2 * zc, bin_, 3, 3)
        s["encoder.conv_out.bias"] = (2 * zc,)
        bin_ = ch * mult[-1]
        s["decoder.conv_in.weight"] = (bin_, zc, 3, 3)
        s["decoder.conv_in.bias"] = (bin_,)
        _res_shapes(s, "decoder.mid.block_1", bin_, bin_)
        _attn_shapes(s, "decoder.mid.attn_1", bin_)
        _res_shapes(s, "decoder.mid.block_2", bin_, bin_)
        for lvl in reversed(range(nres)):
            bout = ch * mult[lvl]
            for ib in range(nrb + 1):
                _res_shapes(s, f"decoder.up.{lvl}.block.{ib}", bin_, bout)
                bin_ = bout
            if lvl != 0:
                s[f"decoder.up.{lvl}.upsample.conv.weight"] = (bin_, bin_, 3, 3)
                s[f"decoder.up.{lvl}.upsample.conv.bias"] = (bin_,)
        s["decoder.norm_out.weight"] = (bin_,)
        s["decoder.norm_out.bias"] = (bin_,)
        s["decoder.conv_out.weight"] = (self.out_ch, bin_, 3, 3)
        s["decoder.conv_out.bias"] = (self.out_ch,)
        s["quant_conv.weight"] = (2 * self.embed_dim, 2 * zc, 1, 1)
        s["quant_conv.bias"] = (2 * self.embed_dim,)
        s["post_quant_conv.weight"] = (zc, self.embed_dim, 1, 1)
        s["post_quant_conv.bias"] = (zc,)
        if self.adaptorconfig is not None:
            ad = self.adaptorconfig
            a = ad["ch"] * ad["ch_mult"][-1]
            # encoder_adaptor (autoencoder_adaptor.py:92-199) is instantiated by the reference and therefore
            # present in vae.ckpt, but never executed at inference: keep its tensors as inert placeholders.
            ai = ad["in_channels"]
            s["encoder_adaptor.conv_in.weight"] = (a, ai, 3, 3)
            s["encoder_adaptor.conv_in.bias"] = (a,)
            for ib in range(ad["num_res_blocks"]):
                _res_shapes(s, f"encoder_adaptor.down.0.block.{ib}", a, a)
            s["encoder_adaptor.norm_out.weight"] = (a,)
            s["encoder_adaptor.norm_out.bias"] = (a,)
            s["encoder_adaptor.conv_out.weight"] = (ai, a, 3, 3)
            s["encoder_adaptor.conv_out.bias"] = (ai,)
            for ib in range(ad["num_res_blocks"] + 1):
                _res_shapes(s, f"decoder_adaptor.up.0.block.{ib}", a, a)
            s["decoder_adaptor.norm_out.weight"] = (a,)
            s["decoder_adaptor.norm_out.bias"] = (a,)
            s["decoder_adaptor.conv_out.weight"] = (ad["out_ch"], a, 3, 3)
            s["decoder_adaptor.conv_out.bias"] = (ad["out_ch"],)
        return s

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    # ------------------------------------------------------------------ packing
    @torch.no_grad()
    def prepare(self):
        require_device()
        sd = {k: v.detach() for k, v in self.state_dict().items() if not k.startswith("encoder_adaptor.")}
        dev = next(iter(sd.values())).device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKL.prepare(): move the model to a CUDA device first")
        P: Dict[str, torch.Tensor] = {}
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        f32 = lambda t: t.float().contiguous()
        for k, w in sd.items():
            if k.endswith(".weight"):
                p = k[:-7]
                if w.dim() == 1:
                    P[f"{p}.g"] = f32(w)
                    P[f"{p}.be"] = f32(sd[f"{p}.bias"])
                elif w.shape[2:] == (3, 3):
                    if ".downsample." in p:
                        P[f"{p}.w"] = bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
                    else:
                        w9 = w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1])
                        P[f"{p}.w"] = bf(_pad_k(w9, -(-w.shape[1] // 64) * 64))
                    P[f"{p}.b"] = f32(sd[f"{p}.bias"])
                elif w.shape[2:] == (1, 1):
                    w2 = w.reshape(w.shape[0], w.shape[1])
                    P[f"{p}.w"] = bf(_pad_k(w2, -(-w.shape[1] // 64) * 64))
                    P[f"{p}.b"] = f32(sd[f"{p}.bias"])
        for side in ("encoder", "decoder"):
            a = f"{side}.mid.attn_1"
            P[f"{a}.qkv.w"] = torch.cat([P[f"{a}.q.w"], P[f"{a}.k.w"], P[f"{a}.v.w"]], 0).contiguous()
            P[f"{a}.qkv.b"] = torch.cat([P[f"{a}.q.b"], P[f"{a}.k.b"], P[f"{a}.v.b"]], 0).contiguous()
        self._packed = P
        return self

    # ------------------------------------------------------------------ blocks (rows [N*H*W, C] bf16)
    def _resnet(self, h, p, N, H, W):
        """ResnetBlock.forward ae_modules.py:228-248 (temb=None)."""
        P = self._packed
        a = ops.groupnorm(h, N, H * W, P[f"{p}.norm1.g"], P[f"{p}.norm1.be"], 1e-6, True)
        h1 = ops.conv3x3(a, N, H, W, P[f"{p}.conv1.w"], P[f"{p}.conv1.b"])
        a2 = ops.groupnorm(h1, N, H * W, P[f"{p}.norm2.g"], P[f"{p}.norm2.be"], 1e-6, True)
        skip = h
        if f"{p}.nin_shortcut.w" in P:
            skip = ops.linear(h, P[f"{p}.nin_shortcut.w"], P[f"{p}.nin_shortcut.b"])
        return ops.conv3x3(a2, N, H, W, P[f"{p}.conv2.w"], P[f"{p}.conv2.b"], residual=skip)

    def _attn(self, h, p, N, H, W):
        """AttnBlock.forward ae_modules.py:53-78: softmax(q k^T / sqrt(C)) v, one head of width C."""
        P = self._packed
        Cc, L = h.shape[1], H * W
        if L % 64 or Cc % 64:
            raise NotImplementedError("VAE AttnBlock needs h*w and C to be multiples of 64 on the tcgen05 path")
        a = ops.groupnorm(h, N, L, P[f"{p}.norm.g"], P[f"{p}.norm.be"], 1e-6, False)
        qkv = ops.linear(a, P[f"{p}.qkv.w"], P[f"{p}.qkv.b"])  # [N*L, 3C]
        q = qkv[:, :Cc].unflatten(0, (N, L))
        k = qkv[:, Cc:2 * Cc]
        v = qkv[:, 2 * Cc:]
        # contiguous K (B operand of the batched matmul must be dense [N, L, C])
        kc = k.contiguous().unflatten(0, (N, L))
        s = ops.bmm_nt(q, kc, alpha=float(Cc) ** -0.5, out_dtype=torch.float32)  # [N, L, L]
        pr = ops.softmax_rows(s.flatten(0, 1)).unflatten(0, (N, L))
        vt = ops.transpose_bf16(v, N, L, Cc)  # [N, C, L]
        o = ops.bmm_nt(pr, vt)  # [N, L, C]
        return ops.linear(o.flatten(0, 1), P[f"{p}.proj_out.w"], P[f"{p}.proj_out.b"], residual=h)

    # ------------------------------------------------------------------ decoder
    @torch.no_grad()
    def _decode_rows(self, z: torch.Tensor, with_pre: bool):
        """z [N, zc, h, w] fp32 (already divided by scale_factor) -> (rgb rows fp32 [N*H*W, out_ch], pre rows)."""
        if self._packed is None:
            self.prepare()
        P = self._packed
        N, zc, h, w = z.shape
        nres = len(self.ch_mult)
        zr = ops.bcthw_to_rows(z.float().unsqueeze(0).transpose(1, 2).contiguous(), None, 64)  # b=1, t=N
        # post_quant_conv (1x1, autoencoder.py:137) into a zero-padded 64-channel buffer
        zq = torch.zeros((N * h * w, 64), device=z.device, dtype=torch.bfloat16)
        ops.linear(zr, P["post_quant_conv.w"], P["post_quant_conv.b"], out=zq[:, :zc])
        x = ops.conv3x3(zq, N, h, w, P["decoder.conv_in.w"], P["decoder.conv_in.b"])
        x = self._resnet(x, "decoder.mid.block_1", N, h, w)
        x = self._attn(x, "decoder.mid.attn_1", N, h, w)
        x = self._resnet(x, "decoder.mid.block_2", N, h, w)
        H, W = h, w
        for lvl in reversed(range(nres)):
            for ib in range(self.num_res_blocks + 1):
                x = self._resnet(x, f"decoder.up.{lvl}.block.{ib}", N, H, W)
            if lvl != 0:
                x = ops.upsample2x(x, N, H, W)
                H, W = 2 * H, 2 * W
                x = ops.conv3x3(x, N, H, W, P[f"decoder.up.{lvl}.upsample.conv.w"],
                                P[f"decoder.up.{lvl}.upsample.conv.b"])
        pre = x
        a = ops.groupnorm(x, N, H * W, P["decoder.norm_out.g"], P["decoder.norm_out.be"], 1e-6, True)
        rgb = ops.conv3x3(a, N, H, W, P["decoder.conv_out.w"], P["decoder.conv_out.b"], out_dtype=torch.float32)
        return rgb, (pre if with_pre else None), (N, H, W)

    def _rows_to_nchw(self, rows, Cc, N, H, W):
        return ops.rows_to_bcthw(rows, Cc, 1, N, H, W)[0].transpose(0, 1)  # [N, C, H, W] view

    @torch.no_grad()
    def decode(self, z, **kwargs):
        """AutoencoderKL.decode autoencoder.py:136-139."""
        rgb, _, (N, H, W) = self._decode_rows(z, False)
        return self._rows_to_nchw(rgb, self.out_ch, N, H, W)

    @torch.no_grad()
    def decode_with_conf_adaptor(self, z, **kwargs):
        """AutoencoderKL.decode_with_conf_adaptor autoencoder.py:120-127 (+ VAEDecoderadaptor.forward
        autoencoder_adaptor.py:277-317): cat([rgb, conf], dim=1)."""
        if self.adaptorconfig is None:
            raise RuntimeError("decode_with_conf_adaptor needs adaptorconfig")
        P = self._packed if self._packed is not None else self.prepare()._packed
        rgb, pre, (N, H, W) = self._decode_rows(z, True)
        x = pre
        for ib in range(self.adaptorconfig["num_res_blocks"] + 1):
            x = self._resnet(x, f"decoder_adaptor.up.0.block.{ib}", N, H, W)
        a = ops.groupnorm(x, N, H * W, P["decoder_adaptor.norm_out.g"], P["decoder_adaptor.norm_out.be"], 1e-6, True)
        conf = ops.conv3x3(a, N, H, W, P["decoder_adaptor.conv_out.w"], P["decoder_adaptor.conv_out.b"],
                           out_dtype=torch.float32)
        oc = self.adaptorconfig["out_ch"]
        return torch.cat([self._rows_to_nchw(rgb, self.out_ch, N, H, W), self._rows_to_nchw(conf, oc, N, H, W)], 1)

    # ------------------------------------------------------------------ encoder
    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder.forward (ae_modules.py:537-580) + quant_conv -> posterior parameters [N, 2*embed, h, w]."""
        if self._packed is None:
            self.prepare()
        P = self._packed
        N, ci, H, W = x.shape
        nres = len(self.ch_mult)
        xr = ops.bcthw_to_rows(x.float().unsqueeze(0).transpose(1, 2).contiguous(), None, 64)
        h = ops.conv3x3(xr, N, H, W, P["encoder.conv_in.w"], P["encoder.conv_in.b"])
        for lvl in range(nres):
            for ib in range(self.num_res_blocks):
                h = self._resnet(h, f"encoder.down.{lvl}.block.{ib}", N, H, W)
            if lvl != nres - 1:
                Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1  # pad (0,1,0,1), stride 2 (ae_modules.py:100-106)
                col = ops.im2col_s2(h, N, H, W, 0, Ho, Wo)
                p = f"encoder.down.{lvl}.downsample.conv"
                h = ops.linear(col, P[f"{p}.w"], P[f"{p}.b"])
                H, W = Ho, Wo
        h = self._resnet(h, "encoder.mid.block_1", N, H, W)
        h = self._attn(h, "encoder.mid.attn_1", N, H, W)
        h = self._resnet(h, "encoder.mid.block_2", N, H, W)
        a = ops.groupnorm(h, N, H * W, P["encoder.norm_out.g"], P["encoder.norm_out.be"], 1e-6, True)
        zc2 = 2 * self.z_channels
        m = torch.zeros((N * H * W, 64), device=x.device, dtype=torch.bfloat16)
        ops.conv3x3(a, N, H, W, P["encoder.conv_out.w"], P["encoder.conv_out.b"], out=m[:, :zc2])
        mom = ops.linear(m, P["quant_conv.w"], P["quant_conv.b"], out_dtype=torch.float32)
        return self._rows_to_nchw(mom, 2 * self.embed_dim, N, H, W)

    @torch.no_grad()
    def encode(self, x, **kwargs):
        """AutoencoderKL.encode autoencoder.py:129-134."""
        return DiagonalGaussianDistribution(self.encode_moments(x))
