"""Image-token Resampler behind the reference's `image_proj_stage_config` seam
(lvdm/modules/encoders/resampler.py:96-164, PerceiverAttention :47-93, FeedForward :27-34): the OpenCLIP image
tokens [b, 257, 1280] -> `num_queries * video_length` conditioning tokens [b, 256, 1024] that the U-Net's image
cross-attention reads (infer_geo4d.py:140-156).  Same constructor arguments and state-dict keys as the reference
class; every matmul / LayerNorm / attention runs through the C-ABI kernels (bf16 tensor cores, fp32 accumulate):
fused q / kv projections, the d = 64 tcgen05 attention kernel (scale (1/d^0.25)^2 = d^-0.5), exact-erf GELU in the
GEMM epilogue, residual adds in the GEMM epilogue.  SURVEY.md 8(f) N3: the OpenCLIP ViT-H towers that feed it stay
outside this port (their output for the shipped settings is a constant)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._cabi import require_device


class Resampler(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, video_length=None):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("geo4d_b200.Resampler: the attention kernel is specialised for head dim 64")
        self.dim, self.depth, self.heads, self.dim_head = dim, depth, heads, dim_head
        self.num_queries, self.video_length = num_queries, video_length
        self.embedding_dim, self.output_dim, self.ff_mult = embedding_dim, output_dim, ff_mult
        nq = num_queries * (video_length or 1)
        inner = dim_head * heads
        P = lambda *s: nn.Parameter(torch.empty(*s), requires_grad=False)
        self.latents = P(1, nq, dim)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList()
        for _ in range(depth):
            attn = nn.Module()
            attn.norm1, attn.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
            attn.to_q = nn.Linear(dim, inner, bias=False)
            attn.to_kv = nn.Linear(dim, inner * 2, bias=False)
            attn.to_out = nn.Linear(inner, dim, bias=False)
            ff = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, int(dim * ff_mult), bias=False), nn.GELU(),
                               nn.Linear(int(dim * ff_mult), dim, bias=False))
            self.layers.append(nn.ModuleList([attn, ff]))
        for p in self.parameters():
            p.requires_grad_(False)
        self._packed = None

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    @torch.no_grad()
    def prepare(self):
        require_device()
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        P = {"proj_in.w": bf(self.proj_in.weight), "proj_in.b": f32(self.proj_in.bias),
             "proj_out.w": bf(self.proj_out.weight), "proj_out.b": f32(self.proj_out.bias),
             "norm_out.g": f32(self.norm_out.weight), "norm_out.b": f32(self.norm_out.bias),
             "latents": bf(self.latents[0])}
        for i, (attn, ff) in enumerate(self.layers):
            for n in ("norm1", "norm2"):
                P[f"{i}.{n}.g"], P[f"{i}.{n}.b"] = f32(getattr(attn, n).weight), f32(getattr(attn, n).bias)
            P[f"{i}.q"], P[f"{i}.kv"], P[f"{i}.out"] = bf(attn.to_q.weight), bf(attn.to_kv.weight), bf(attn.to_out.weight)
            P[f"{i}.ff.g"], P[f"{i}.ff.b"] = f32(ff[0].weight), f32(ff[0].bias)
            P[f"{i}.ff.w1"], P[f"{i}.ff.w2"] = bf(ff[1].weight), bf(ff[3].weight)
        self._packed = P
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [b, L, embedding_dim] (or [b, t, L, C] with per-frame image tokens) -> [b, nq, output_dim] fp32."""
        if self._packed is None:
            self.prepare()
        P = self._packed
        per_frame = x.dim() == 4
        if per_frame:
            B, T, L, C = x.shape
            x = x.reshape(B * T, L, C)
            nq = self.num_queries
            lat0 = P["latents"].reshape(T, nq, self.dim).repeat(B, 1, 1).reshape(B * T * nq, self.dim)
        else:
            nq = P["latents"].shape[0]
            lat0 = P["latents"].repeat(x.shape[0], 1)
        b, L, _ = x.shape
        inner = self.heads * 64
        dev = x.device
        xr = ops.linear(x.reshape(b * L, -1).to(torch.bfloat16).contiguous(), P["proj_in.w"], P["proj_in.b"])
        lat = lat0.contiguous().clone()
        kv = torch.empty((b * (L + nq), 2 * inner), device=dev, dtype=torch.bfloat16)
        kvin = torch.empty((b, L + nq, self.dim), device=dev, dtype=torch.bfloat16)
        o = torch.empty((b * nq, inner), device=dev, dtype=torch.bfloat16)
        for i in range(self.depth):
            xn = ops.layernorm(xr, P[f"{i}.norm1.g"], P[f"{i}.norm1.b"])
            ln = ops.layernorm(lat, P[f"{i}.norm2.g"], P[f"{i}.norm2.b"])
            q = ops.linear(ln, P[f"{i}.q"])
            kvin[:, :L] = xn.view(b, L, self.dim)              # keys / values over cat(x, latents) (resampler.py:79-80)
            kvin[:, L:] = ln.view(b, nq, self.dim)
            ops.linear(kvin.view(b * (L + nq), self.dim), P[f"{i}.kv"], out=kv)
            ops.attention(q, kv[:, :inner], kv[:, inner:], o, b, self.heads, nq, L + nq)
            lat = ops.linear(o, P[f"{i}.out"], residual=lat)
            h = ops.linear(ops.layernorm(lat, P[f"{i}.ff.g"], P[f"{i}.ff.b"]), P[f"{i}.ff.w1"], act=ops.ACT_GELU)
            lat = ops.linear(h, P[f"{i}.ff.w2"], residual=lat)
        out = ops.linear(lat, P["proj_out.w"], P["proj_out.b"])
        out = ops.layernorm(out, P["norm_out.g"], P["norm_out.b"]).float()
        if per_frame:
            return out.view(B, T * nq, self.output_dim)
        return out.view(b, nq, self.output_dim)
