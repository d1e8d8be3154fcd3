#!/usr/bin/env python
"""Where the time of a tap-GEMM launch goes: per-CTA %globaltimer stamps (geo4d_debug_gemm_trace) for the
hot U-Net shapes + back-to-back timings of the same launch (CUDA events, 50 launches)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops

dev = torch.device("cuda")
lib = ops.lib()

def trace(fn, label, flops):
    buf = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.geo4d_debug_gemm_trace(ops._vp(buf))
    fn()
    torch.cuda.synchronize()
    lib.geo4d_debug_gemm_trace(None)
    t = buf.cpu().double()
    used = t[:, 0] > 0
    t = t[used]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    def med(c):
        v = rel[:, c][t[:, c] > 0]
        return float(v.median()) if len(v) else float("nan")
    def cd(a, b):
        ok = (t[:, a] > 0) & (t[:, b] > 0)
        return float((t[ok][:, a] - t[ok][:, b]).median()) if bool(ok.any()) else float("nan")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(f"{label:44s} {us:7.1f} us/launch {flops / us * 1e-6:7.1f} TF/s | ctas {int(used.sum()):3d} entry-spread {float(rel[:,0].max()):5.1f} "
          f"setup {med(1):5.1f} first-operands {med(2):5.1f} tile0-mma-issued {med(3):5.1f} acc0-ready {med(4):5.1f} "
          f"epi0-done {med(5):5.1f} last-epi {med(6):5.1f} exit med {med(7):5.1f} max {float(rel[:,7].max()):5.1f} | chunk0 clk: ld {cd(8,13):5.0f} math {cd(9,8):5.0f} barA {cd(10,9):5.0f} sts+fence {cd(11,10):5.0f} barB+tma {cd(12,11):5.0f}", flush=True)

def lin(M, K, N, act=0, residual=False, bias=True):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N // 2 if act == 2 else N, device=dev).bfloat16() if residual else None
    trace(lambda: ops.linear(x, w, b, act=act, residual=r), f"linear {M}x{K}->{N} act{act} res{int(residual)}", 2.0 * M * K * N)

def conv(Nf, H, W, Cin, Cout):
    x = torch.randn(Nf * H * W, Cin, device=dev).bfloat16()
    w = (torch.randn(9, Cout, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device=dev)
    trace(lambda: ops.conv3x3(x, Nf, H, W, w, b), f"conv3x3 {Nf}x{H}x{W} {Cin}->{Cout}", 2.0 * Nf * H * W * 9 * Cin * Cout)

def tconv(B, T, HW, C):
    x = torch.randn(B * T * HW, C, device=dev).bfloat16()
    w = (torch.randn(3, C, C, device=dev) / (3 * C) ** 0.5).bfloat16()
    b = torch.randn(C, device=dev)
    trace(lambda: ops.temporal_conv3(x, B, T, HW, w, b), f"temporal_conv3 {T}x{HW} {C}", 2.0 * B * T * HW * 3 * C * C)

MODE = os.environ.get("DIRECT", "0")
lib.geo4d_debug_gemm_direct_store(int(MODE))
lib.geo4d_debug_gemm_pair_mode(int(os.environ.get("PAIR", "-1")))
print("direct_store =", MODE)
lin(40960, 320, 320); lin(40960, 320, 320, residual=True); lin(40960, 320, 960, bias=False); lin(40960, 320, 2560, act=2); lin(40960, 1280, 320, residual=True)
lin(10240, 640, 640); lin(10240, 640, 1920, bias=False); lin(10240, 640, 5120, act=2); lin(10240, 2560, 640, residual=True)
lin(2560, 1280, 1280); lin(2560, 1280, 3840, bias=False); lin(2560, 1280, 10240, act=2); lin(2560, 5120, 1280, residual=True)
conv(16, 40, 64, 320, 320); conv(16, 20, 32, 640, 640); conv(16, 10, 16, 1280, 1280); conv(16, 5, 8, 1280, 1280)
tconv(1, 16, 2560, 320); tconv(1, 16, 640, 640); tconv(1, 16, 160, 1280); tconv(1, 16, 40, 1280)
conv(4, 320, 512, 128, 128)
