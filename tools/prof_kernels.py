#!/usr/bin/env python
"""Representative launches of every hot kernel at the 320x512x16f U-Net shapes (for ncu captures).
Usage under ncu:  ncu --set full --clock-control none --import-source on -k regex:'tap_gemm|attn_fwd|gn_|layernorm|temporal_attn|align_iter' \
                      -o gpurun_out/prof python tools/prof_kernels.py"""
import os, sys
os.environ.setdefault("GEO4D_AUTOTUNE", "0")   # under ncu every tuning launch would be profiled too
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops

def main():
    dev = "cuda"
    torch.manual_seed(0)
    bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
    reps = int(os.environ.get("REPS", "1"))
    M = 40960
    for _ in range(reps):
        # linears (level 0)
        x320 = bf(M, 320); x1280 = bf(M, 1280)
        ops.linear(x320, bf(320, 320), torch.randn(320, device=dev))
        ops.linear(x320, bf(960, 320))
        ops.linear(x1280, bf(320, 1280), torch.randn(320, device=dev), residual=x320)
        ops.linear(x320, bf(2560, 320), torch.randn(2560, device=dev), act=ops.ACT_GEGLU)
        # linears (level 1, 2)
        x640 = bf(10240, 640); ops.linear(x640, bf(1920, 640)); ops.linear(x640, bf(5120, 640), torch.randn(5120, device=dev), act=ops.ACT_GEGLU)
        x12 = bf(2560, 1280); ops.linear(x12, bf(3840, 1280)); ops.linear(x12, bf(10240, 1280), torch.randn(10240, device=dev), act=ops.ACT_GEGLU)
        # convs
        ops.conv3x3(x320, 16, 40, 64, bf(9, 320, 320), torch.randn(320, device=dev))
        ops.conv3x3(bf(M, 960), 16, 40, 64, bf(9, 320, 960), torch.randn(320, device=dev))
        ops.conv3x3(x640, 16, 20, 32, bf(9, 640, 640), torch.randn(640, device=dev))
        ops.conv3x3(x12, 16, 10, 16, bf(9, 1280, 1280), torch.randn(1280, device=dev))
        ops.conv3x3(bf(640, 1280), 16, 5, 8, bf(9, 1280, 1280), torch.randn(1280, device=dev))
        ops.temporal_conv3(x320, 1, 16, 2560, bf(3, 320, 320), torch.randn(320, device=dev))
        # VAE-like conv
        ops.conv3x3(bf(4 * 320 * 512, 128), 4, 320, 512, bf(9, 128, 128), torch.randn(128, device=dev))
        # attention
        qkv = bf(M, 960); o = torch.empty(M, 320, device=dev, dtype=torch.bfloat16)
        ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], o, 16, 5, 2560, 2560)
        kv = bf(77, 640); ops.attention(qkv[:, :320], kv[:, :320], kv[:, 320:], o, 16, 5, 2560, 77, kv_batch_div=16)
        ops.temporal_attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], o, 1, 16, 2560, 5)
        # norms
        g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
        ops.groupnorm(x320, 16, 2560, g, b, 1e-5, True)
        ops.groupnorm(x320, 1, 16 * 2560, g, b, 1e-5, True)
        ops.layernorm(x320, g, b)
    torch.cuda.synchronize()
    print("done")

if __name__ == "__main__":
    main()
