"""fp32 CPU restatement of the reference spatio-temporal U-Net (test oracle).

Follows lvdm/modules/networks/openaimodel3d.py (UNetModel.__init__:311-556,
UNetModel.forward:558-633, ResBlock._forward:210-236,
TemporalConvBlock.forward:272-279, Downsample:75-77, Upsample:98-106) and
lvdm/modules/attention.py (CrossAttention.forward:81-144,
BasicTransformerBlock._forward:242-246, SpatialTransformer.forward:294-310,
TemporalTransformer.forward:365-412, GEGLU:415-422) of jzr99/Geo4D.

The network is described by a flat "plan" (list of block descriptors) derived
from the same constructor arguments as the reference's YAML
(configs/inference_geo4d.yaml:62-89).  `param_shapes` enumerates the
state-dict keys/shapes the reference module would own, `init_params` draws a
seeded synthetic state dict (no checkpoint exists offline), and `forward`
evaluates the network with plain torch.nn.functional ops in fp32.

Pinned by oracle/gen_golden.py against the imported reference module
(strict state-dict load + output comparison).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    # configs/inference_geo4d.yaml:62-89 (defaults = the shipped Geo4D model)
    in_channels: int = 20
    out_channels: int = 16
    model_channels: int = 320
    attention_resolutions: Sequence[int] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Sequence[int] = (1, 2, 4, 4)
    num_head_channels: int = 64
    transformer_depth: int = 1
    context_dim: int = 1024
    use_linear: bool = True
    temporal_conv: bool = True
    temporal_attention: bool = True
    temporal_length: int = 16
    addition_attention: bool = True
    image_cross_attention: bool = True
    default_fs: int = 24
    fs_condition: bool = True
    text_context_len: int = 77
    init_attn_heads: int = 8  # openaimodel3d.py:403 hard-codes n_heads=8

    @staticmethod
    def tiny(**kw) -> "UNetConfig":
        """Small config with the same topology, for fast parity tests."""
        base = dict(model_channels=64, context_dim=64, temporal_length=4)
        base.update(kw)
        return UNetConfig(**base)


# --------------------------------------------------------------------------- plan

def build_plan(cfg: UNetConfig) -> dict:
    """Mirror of UNetModel.__init__ (openaimodel3d.py:385-556) as data."""
    mc = cfg.model_channels
    plan = {"input_blocks": [], "middle_block": [], "output_blocks": []}
    plan["input_blocks"].append([("conv", "input_blocks.0.0", cfg.in_channels, mc)])
    chans = [mc]
    ch, ds = mc, 1
    idx = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", f"input_blocks.{idx}.0", ch, mult * mc, cfg.temporal_conv)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                heads = ch // cfg.num_head_channels
                layers.append(("st", f"input_blocks.{idx}.1", ch, heads))
                if cfg.temporal_attention:
                    layers.append(("tt", f"input_blocks.{idx}.2", ch, heads, cfg.use_linear))
            plan["input_blocks"].append(layers)
            chans.append(ch)
            idx += 1
        if level != len(cfg.channel_mult) - 1:
            plan["input_blocks"].append([("down", f"input_blocks.{idx}.0", ch, ch)])
            chans.append(ch)
            ds *= 2
            idx += 1
    heads = ch // cfg.num_head_channels
    mid = [("res", "middle_block.0", ch, ch, cfg.temporal_conv),
           ("st", "middle_block.1", ch, heads)]
    if cfg.temporal_attention:
        mid.append(("tt", "middle_block.2", ch, heads, cfg.use_linear))
    mid.append(("res", f"middle_block.{len(mid)}", ch, ch, cfg.temporal_conv))
    plan["middle_block"] = mid
    idx = 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", f"output_blocks.{idx}.0", ch + ich, mc * mult, cfg.temporal_conv)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                heads = ch // cfg.num_head_channels
                layers.append(("st", f"output_blocks.{idx}.1", ch, heads))
                if cfg.temporal_attention:
                    layers.append(("tt", f"output_blocks.{idx}.2", ch, heads, cfg.use_linear))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", f"output_blocks.{idx}.{len(layers)}", ch, ch))
                ds //= 2
            plan["output_blocks"].append(layers)
            idx += 1
    plan["final_ch"] = ch
    return plan


# --------------------------------------------------------------------------- parameter inventory

def _attn_shapes(p, dim, heads, dh, ctx_dim, ip):
    inner = heads * dh
    s = OrderedDict()
    s[f"{p}.to_q.weight"] = (inner, dim)
    s[f"{p}.to_k.weight"] = (inner, ctx_dim)
    s[f"{p}.to_v.weight"] = (inner, ctx_dim)
    s[f"{p}.to_out.0.weight"] = (dim, inner)
    s[f"{p}.to_out.0.bias"] = (dim,)
    if ip:
        s[f"{p}.to_k_ip.weight"] = (inner, ctx_dim)
        s[f"{p}.to_v_ip.weight"] = (inner, ctx_dim)
    return s


def _btb_shapes(p, dim, heads, dh, ctx_dim, ip):
    """BasicTransformerBlock (attention.py:214-233): attn1 self, attn2 cross."""
    s = OrderedDict()
    s.update(_attn_shapes(f"{p}.attn1", dim, heads, dh, dim, False))
    s[f"{p}.ff.net.0.proj.weight"] = (dim * 4 * 2, dim)
    s[f"{p}.ff.net.0.proj.bias"] = (dim * 4 * 2,)
    s[f"{p}.ff.net.2.weight"] = (dim, dim * 4)
    s[f"{p}.ff.net.2.bias"] = (dim,)
    s.update(_attn_shapes(f"{p}.attn2", dim, heads, dh, ctx_dim if ctx_dim else dim, ip))
    for n in ("norm1", "norm2", "norm3"):
        s[f"{p}.{n}.weight"] = (dim,)
        s[f"{p}.{n}.bias"] = (dim,)
    return s


def _res_shapes(p, cin, cout, emb_dim, tconv):
    s = OrderedDict()
    s[f"{p}.in_layers.0.weight"] = (cin,)
    s[f"{p}.in_layers.0.bias"] = (cin,)
    s[f"{p}.in_layers.2.weight"] = (cout, cin, 3, 3)
    s[f"{p}.in_layers.2.bias"] = (cout,)
    s[f"{p}.emb_layers.1.weight"] = (cout, emb_dim)
    s[f"{p}.emb_layers.1.bias"] = (cout,)
    s[f"{p}.out_layers.0.weight"] = (cout,)
    s[f"{p}.out_layers.0.bias"] = (cout,)
    s[f"{p}.out_layers.3.weight"] = (cout, cout, 3, 3)
    s[f"{p}.out_layers.3.bias"] = (cout,)
    if cin != cout:
        s[f"{p}.skip_connection.weight"] = (cout, cin, 1, 1)
        s[f"{p}.skip_connection.bias"] = (cout,)
    if tconv:
        for k, ci in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
            s[f"{p}.temopral_conv.{k}.0.weight"] = (cout,)
            s[f"{p}.temopral_conv.{k}.0.bias"] = (cout,)
            s[f"{p}.temopral_conv.{k}.{ci}.weight"] = (cout, cout, 3, 1, 1)
            s[f"{p}.temopral_conv.{k}.{ci}.bias"] = (cout,)
    return s


def _st_shapes(p, ch, heads, dh, cfg: UNetConfig):
    inner = heads * dh
    s = OrderedDict()
    s[f"{p}.norm.weight"] = (ch,)
    s[f"{p}.norm.bias"] = (ch,)
    if cfg.use_linear:
        s[f"{p}.proj_in.weight"] = (inner, ch)
    else:
        s[f"{p}.proj_in.weight"] = (inner, ch, 1, 1)
    s[f"{p}.proj_in.bias"] = (inner,)
    for d in range(cfg.transformer_depth):
        s.update(_btb_shapes(f"{p}.transformer_blocks.{d}", inner, heads, dh,
                             cfg.context_dim, cfg.image_cross_attention))
    s[f"{p}.proj_out.weight"] = (ch, inner) if cfg.use_linear else (ch, inner, 1, 1)
    s[f"{p}.proj_out.bias"] = (ch,)
    return s


def _tt_shapes(p, ch, heads, dh, use_linear, depth):
    inner = heads * dh
    s = OrderedDict()
    s[f"{p}.norm.weight"] = (ch,)
    s[f"{p}.norm.bias"] = (ch,)
    s[f"{p}.proj_in.weight"] = (inner, ch) if use_linear else (inner, ch, 1)
    s[f"{p}.proj_in.bias"] = (inner,)
    for d in range(depth):
        # only_self_att => context_dim None, no image branch (attention.py:347-348)
        s.update(_btb_shapes(f"{p}.transformer_blocks.{d}", inner, heads, dh, None, False))
    s[f"{p}.proj_out.weight"] = (ch, inner) if use_linear else (ch, inner, 1)
    s[f"{p}.proj_out.bias"] = (ch,)
    return s


def param_shapes(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    mc = cfg.model_channels
    emb = mc * 4
    dh = cfg.num_head_channels
    s = OrderedDict()
    for name in ("time_embed",) + (("fps_embedding",) if cfg.fs_condition else ()):
        s[f"{name}.0.weight"] = (emb, mc)
        s[f"{name}.0.bias"] = (emb,)
        s[f"{name}.2.weight"] = (emb, emb)
        s[f"{name}.2.bias"] = (emb,)
    plan = build_plan(cfg)

    def add(layers):
        for L in layers:
            kind, p = L[0], L[1]
            if kind == "conv":
                s[f"{p}.weight"] = (L[3], L[2], 3, 3)
                s[f"{p}.bias"] = (L[3],)
            elif kind == "res":
                s.update(_res_shapes(p, L[2], L[3], emb, L[4]))
            elif kind == "st":
                s.update(_st_shapes(p, L[2], L[3], dh, cfg))
            elif kind == "tt":
                s.update(_tt_shapes(p, L[2], L[3], dh, L[4], cfg.transformer_depth))
            elif kind == "down":
                s[f"{p}.op.weight"] = (L[3], L[2], 3, 3)
                s[f"{p}.op.bias"] = (L[3],)
            elif kind == "up":
                s[f"{p}.conv.weight"] = (L[3], L[2], 3, 3)
                s[f"{p}.conv.bias"] = (L[3],)

    for layers in plan["input_blocks"]:
        add(layers)
    if cfg.addition_attention:
        # init_attn: TemporalTransformer(mc, n_heads=8, d_head, use_linear=False);
        # registered after the input_blocks ModuleList (openaimodel3d.py:398-409)
        s.update(_tt_shapes("init_attn.0", mc, cfg.init_attn_heads, dh, False,
                            cfg.transformer_depth))
    add(plan["middle_block"])
    for layers in plan["output_blocks"]:
        add(layers)
    s["out.0.weight"] = (plan["final_ch"],)
    s["out.0.bias"] = (plan["final_ch"],)
    s["out.2.weight"] = (cfg.out_channels, mc, 3, 3)
    s["out.2.bias"] = (cfg.out_channels,)
    return s


def init_params(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int = 0,
                gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights.  A freshly constructed reference net outputs
    exactly zero (zero-initialised out convs, SURVEY.md 'Key facts'), so every
    tensor gets a non-trivial value: matrices U(-b, b) with b = gain/sqrt(fan_in),
    norm scales 1 + 0.1 N(0,1), biases 0.1-scaled.  Draw order = key order."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in shapes.items():
        if len(shp) == 1:
            is_norm_w = k.endswith(".weight")
            t = torch.randn(shp, generator=g) * 0.1
            if is_norm_w:
                t = t + 1.0
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            b = gain / math.sqrt(fan_in)
            t = (torch.rand(shp, generator=g) * 2 - 1) * b
        sd[k] = t
    return sd


# --------------------------------------------------------------------------- functional forward

def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """utils_diffusion.py:8-28 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None].to(timesteps.device)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, p, eps):
    return F.group_norm(x.float(), 32, sd[f"{p}.weight"], sd[f"{p}.bias"], eps)


def _lin(x, sd, p, bias=True):
    return F.linear(x, sd[f"{p}.weight"], sd.get(f"{p}.bias") if bias else None)


def _attention(q, k, v, heads):
    """softmax(q k^T / sqrt(d)) v per head; attention.py:101-125 (einsum path)."""
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    qh, kh, vh = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh) * (d ** -0.5)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", sim, vh)
    return out.permute(0, 2, 1, 3).reshape(b, n, inner)


def _cross_attention(x, sd, p, heads, context=None, ip=False, text_len=77):
    """CrossAttention.forward attention.py:81-144 (== efficient_forward 146-209)."""
    q = _lin(x, sd, f"{p}.to_q", bias=False)
    if context is None:
        k = _lin(x, sd, f"{p}.to_k", bias=False)
        v = _lin(x, sd, f"{p}.to_v", bias=False)
        out = _attention(q, k, v, heads)
    else:
        ctx_t = context[:, :text_len]
        k = _lin(ctx_t, sd, f"{p}.to_k", bias=False)
        v = _lin(ctx_t, sd, f"{p}.to_v", bias=False)
        out = _attention(q, k, v, heads)
        if ip:
            ctx_i = context[:, text_len:]
            k_ip = _lin(ctx_i, sd, f"{p}.to_k_ip", bias=False)
            v_ip = _lin(ctx_i, sd, f"{p}.to_v_ip", bias=False)
            out = out + 1.0 * _attention(q, k_ip, v_ip, heads)  # image_cross_attention_scale=1.0
    return _lin(out, sd, f"{p}.to_out.0")


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], 1e-5)


def _basic_block(x, sd, p, heads, context, ip, text_len):
    """attention.py:242-246."""
    x = _cross_attention(_ln(x, sd, f"{p}.norm1"), sd, f"{p}.attn1", heads) + x
    x = _cross_attention(_ln(x, sd, f"{p}.norm2"), sd, f"{p}.attn2", heads,
                         context=context, ip=ip, text_len=text_len) + x
    h = _lin(_ln(x, sd, f"{p}.norm3"), sd, f"{p}.ff.net.0.proj")
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    x = _lin(h, sd, f"{p}.ff.net.2") + x
    return x


def _spatial_transformer(x, sd, p, heads, context, cfg: UNetConfig):
    """attention.py:294-310 (use_linear path and conv path)."""
    n, c, hh, ww = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    if not cfg.use_linear:
        x = F.conv2d(x, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(n, hh * ww, -1)
    if cfg.use_linear:
        x = _lin(x, sd, f"{p}.proj_in")
    for d in range(cfg.transformer_depth):
        x = _basic_block(x, sd, f"{p}.transformer_blocks.{d}", heads, context,
                         cfg.image_cross_attention, cfg.text_context_len)
    if cfg.use_linear:
        x = _lin(x, sd, f"{p}.proj_out")
    x = x.reshape(n, hh, ww, -1).permute(0, 3, 1, 2)
    if not cfg.use_linear:
        x = F.conv2d(x, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return x + x_in


def _temporal_transformer(x, sd, p, heads, use_linear, b, depth):
    """attention.py:365-412, only_self_att branch.  x: (b*t, c, h, w)."""
    bt, c, hh, ww = x.shape
    t = bt // b
    x5 = x.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)  # b c t h w
    x_in = x5
    h = F.group_norm(x5, 32, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    h = h.permute(0, 3, 4, 1, 2).reshape(b * hh * ww, c, t)  # (b h w) c t
    if not use_linear:
        h = F.conv1d(h, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    h = h.permute(0, 2, 1)  # bhw t c
    if use_linear:
        h = _lin(h, sd, f"{p}.proj_in")
    for d in range(depth):
        h = _basic_block(h, sd, f"{p}.transformer_blocks.{d}", heads, None, False, 77)
    if use_linear:
        h = _lin(h, sd, f"{p}.proj_out")
        h = h.reshape(b, hh, ww, t, c).permute(0, 4, 3, 1, 2)
    else:
        h = h.permute(0, 2, 1)
        h = F.conv1d(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
        h = h.reshape(b, hh, ww, c, t).permute(0, 3, 4, 1, 2)
    out = h + x_in
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def _temporal_conv(x, sd, p, b):
    """TemporalConvBlock openaimodel3d.py:272-279; x: (b*t, c, h, w)."""
    bt, c, hh, ww = x.shape
    t = bt // b
    x5 = x.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)
    h = x5
    for k, ci in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
        h = F.group_norm(h, 32, sd[f"{p}.{k}.0.weight"], sd[f"{p}.{k}.0.bias"], 1e-5)
        h = F.silu(h)
        h = F.conv3d(h, sd[f"{p}.{k}.{ci}.weight"], sd[f"{p}.{k}.{ci}.bias"], padding=(1, 0, 0))
    out = x5 + h
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def _resblock(x, emb, sd, p, cin, cout, tconv, b):
    """ResBlock._forward openaimodel3d.py:210-236 (no up/down, no scale-shift)."""
    h = F.silu(_gn(x, sd, f"{p}.in_layers.0", 1e-5))
    h = F.conv2d(h, sd[f"{p}.in_layers.2.weight"], sd[f"{p}.in_layers.2.bias"], padding=1)
    e = _lin(F.silu(emb), sd, f"{p}.emb_layers.1")
    h = h + e[:, :, None, None]
    h = F.silu(_gn(h, sd, f"{p}.out_layers.0", 1e-5))
    h = F.conv2d(h, sd[f"{p}.out_layers.3.weight"], sd[f"{p}.out_layers.3.bias"], padding=1)
    if cin != cout:
        x = F.conv2d(x, sd[f"{p}.skip_connection.weight"], sd[f"{p}.skip_connection.bias"])
    h = x + h
    if tconv:
        h = _temporal_conv(h, sd, f"{p}.temopral_conv", b)
    return h


def _run_layers(layers, h, emb, context, sd, cfg: UNetConfig, b):
    for L in layers:
        kind, p = L[0], L[1]
        if kind == "conv":
            h = F.conv2d(h, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
        elif kind == "res":
            h = _resblock(h, emb, sd, p, L[2], L[3], L[4], b)
        elif kind == "st":
            h = _spatial_transformer(h, sd, p, L[3], context, cfg)
        elif kind == "tt":
            h = _temporal_transformer(h, sd, p, L[3], L[4], b, cfg.transformer_depth)
        elif kind == "down":
            h = F.conv2d(h, sd[f"{p}.op.weight"], sd[f"{p}.op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)
    return h


@torch.no_grad()
def forward(cfg: UNetConfig, sd: Dict[str, torch.Tensor], x: torch.Tensor,
            timesteps: torch.Tensor, context: torch.Tensor,
            fs: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """UNetModel.forward openaimodel3d.py:558-633.

    x [b, in_ch, t, h, w]; timesteps [b]; context [b, 77 + t*16, ctx_dim];
    fs [b] (long).  `taps`, if given, is filled with intermediate activations
    (keyed by block name) for layer-wise parity debugging.
    """
    b, _, t, _, _ = x.shape
    mc = cfg.model_channels
    t_emb = timestep_embedding(timesteps, mc)
    emb = _lin(F.silu(_lin(t_emb, sd, "time_embed.0")), sd, "time_embed.2")
    l_ctx = context.shape[1]
    if l_ctx == cfg.text_context_len + t * 16:  # hard-coded split, :575
        ctx_text = context[:, :cfg.text_context_len].repeat_interleave(t, dim=0)
        ctx_img = context[:, cfg.text_context_len:].reshape(b * t, 16, -1)
        context = torch.cat([ctx_text, ctx_img], dim=1)
    else:
        context = context.repeat_interleave(t, dim=0)
    emb = emb.repeat_interleave(t, dim=0)
    h = x.permute(0, 2, 1, 3, 4).reshape(b * t, x.shape[1], x.shape[3], x.shape[4])
    if cfg.fs_condition:
        if fs is None:
            fs = torch.full((b,), cfg.default_fs, dtype=torch.long)
        fs_emb = timestep_embedding(fs, mc)
        fs_embed = _lin(F.silu(_lin(fs_emb, sd, "fps_embedding.0")), sd, "fps_embedding.2")
        emb = emb + fs_embed.repeat_interleave(t, dim=0)
    plan = build_plan(cfg)
    hs = []
    for i, layers in enumerate(plan["input_blocks"]):
        h = _run_layers(layers, h, emb, context, sd, cfg, b)
        if i == 0 and cfg.addition_attention:
            h = _temporal_transformer(h, sd, "init_attn.0", cfg.init_attn_heads, False, b,
                                      cfg.transformer_depth)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_layers(plan["middle_block"], h, emb, context, sd, cfg, b)
    if taps is not None:
        taps["middle_block"] = h
    for i, layers in enumerate(plan["output_blocks"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(layers, h, emb, context, sd, cfg, b)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    y = F.silu(_gn(h, sd, "out.0", 1e-5))
    y = F.conv2d(y, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return y.reshape(b, t, cfg.out_channels, y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
