#!/usr/bin/env python
"""GPU-bound per-op timings: each op is captured 20x into a CUDA graph and replayed (no CPU launch gaps).
Inputs stay the same between calls, i.e. L2-warm like in the real step where the producer just wrote them."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops

dev = torch.device("cuda")

def gtime(fn, n=20, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)

def bf(*s): return torch.randn(*s, device=dev).bfloat16()

def main():
    which = sys.argv[1:] or ["norm", "gemm", "attn"]
    if "norm" in which:
        for rows, C, S in [(40960, 320, 1), (40960, 320, 16), (10240, 640, 1), (10240, 640, 16), (2560, 1280, 1), (2560, 1280, 16), (640, 1280, 1), (640, 1280, 16), (40960, 960, 16), (40960, 640, 16)]:
            x = bf(rows, C); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); out = torch.empty_like(x)
            us = gtime(lambda: ops.groupnorm(x, S, rows // S, g, b, 1e-5, True, out=out))
            ref = torch.nn.functional.silu(torch.nn.functional.group_norm(x.float().view(S, rows // S, C).permute(0, 2, 1), 32, g, b, 1e-5)).permute(0, 2, 1).reshape(rows, C)
            err = float((out.float() - ref).abs().max())
            print(f"groupnorm+silu rows={rows} C={C} stats={S}: {us:7.2f} us  ({3 * rows * C * 2 / us * 1e-6:.2f} TB/s algorithmic)  max err {err:.3e}", flush=True)
        for rows, C in [(40960, 320), (10240, 640), (2560, 1280)]:
            x = bf(rows, C); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); out = torch.empty_like(x)
            us = gtime(lambda: ops.layernorm(x, g, b, out=out))
            err = float((out.float() - torch.nn.functional.layer_norm(x.float(), (C,), g, b)).abs().max())
            print(f"layernorm rows={rows} C={C}: {us:7.2f} us  ({2 * rows * C * 2 / us * 1e-6:.2f} TB/s)  max err {err:.3e}", flush=True)
    if "gemm" in which:
        def lin(M, K, N, act=0, residual=False, bias=True):
            x = bf(M, K); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
            b = torch.randn(N, device=dev) if bias else None
            r = bf(M, N // 2 if act == 2 else N) if residual else None
            out = torch.empty(M, N // 2 if act == 2 else N, device=dev, dtype=torch.bfloat16)
            us = gtime(lambda: ops.linear(x, w, b, act=act, residual=r, out=out))
            ust = gtime(lambda: torch.nn.functional.linear(x, w))
            print(f"linear {M}x{K}->{N} act{act} res{int(residual)}: {us:7.2f} us {2.0 * M * K * N / us * 1e-6:7.1f} TF/s | torch matmul (no epilogue) {ust:7.2f} us", flush=True)
        def conv(Nf, H, W, Cin, Cout):
            x = bf(Nf * H * W, Cin); w = (torch.randn(9, Cout, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16(); b = torch.randn(Cout, device=dev)
            out = torch.empty(Nf * H * W, Cout, device=dev, dtype=torch.bfloat16)
            us = gtime(lambda: ops.conv3x3(x, Nf, H, W, w, b, out=out))
            xt = x.view(Nf, H, W, Cin).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            wt = w.view(3, 3, Cout, Cin).permute(2, 3, 0, 1).contiguous(memory_format=torch.channels_last)
            ust = gtime(lambda: torch.nn.functional.conv2d(xt, wt, padding=1))
            print(f"conv3x3 {Nf}x{H}x{W} {Cin}->{Cout}: {us:7.2f} us {2.0 * Nf * H * W * 9 * Cin * Cout / us * 1e-6:7.1f} TF/s | cudnn {ust:7.2f} us", flush=True)
        def tconv(T, HW, Cc):
            x = bf(T * HW, Cc); w = (torch.randn(3, Cc, Cc, device=dev) / (3 * Cc) ** 0.5).bfloat16(); b = torch.randn(Cc, device=dev)
            out = torch.empty(T * HW, Cc, device=dev, dtype=torch.bfloat16)
            us = gtime(lambda: ops.temporal_conv3(x, 1, T, HW, w, b, out=out))
            print(f"temporal_conv3 {T}x{HW} {Cc}: {us:7.2f} us {2.0 * T * HW * 3 * Cc * Cc / us * 1e-6:7.1f} TF/s", flush=True)
        lin(40960, 320, 320); lin(40960, 320, 320, residual=True); lin(40960, 320, 960, bias=False); lin(40960, 320, 2560, act=2); lin(40960, 1280, 320, residual=True)
        lin(10240, 640, 640); lin(10240, 640, 1920, bias=False); lin(10240, 640, 5120, act=2); lin(10240, 2560, 640, residual=True)
        lin(2560, 1280, 1280); lin(2560, 1280, 3840, bias=False); lin(2560, 1280, 10240, act=2); lin(2560, 5120, 1280, residual=True)
        conv(16, 40, 64, 320, 320); conv(16, 40, 64, 640, 320); conv(16, 20, 32, 640, 640); conv(16, 10, 16, 1280, 1280); conv(16, 5, 8, 1280, 1280)
        tconv(16, 2560, 320); tconv(16, 640, 640); tconv(16, 160, 1280); tconv(16, 40, 1280)
    if "attn" in which:
        for (B, Hh, Lq, Lk) in [(16, 5, 2560, 2560), (16, 10, 640, 640), (16, 20, 160, 160), (16, 5, 2560, 77)]:
            inner = Hh * 64
            qkv = bf(B * Lq, 3 * inner); o = torch.empty(B * Lq, inner, device=dev, dtype=torch.bfloat16)
            if Lk == Lq:
                us = gtime(lambda: ops.attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], o, B, Hh, Lq, Lk))
            else:
                kv = bf(Lk, 2 * inner)
                us = gtime(lambda: ops.attention(qkv[:, :inner], kv[:, :inner], kv[:, inner:], o, B, Hh, Lq, Lk, kv_batch_div=B))
            q4 = qkv[:, :inner].reshape(B, Lq, Hh, 64).transpose(1, 2)
            k4 = (qkv[:, inner:2 * inner] if Lk == Lq else kv[:, :inner].expand(B * Lk, inner) if False else None)
            print(f"attention B={B} H={Hh} Lq={Lq} Lk={Lk}: {us:7.2f} us {4.0 * B * Hh * Lq * Lk * 64 / us * 1e-6:7.1f} TF/s", flush=True)

if __name__ == "__main__":
    main()
