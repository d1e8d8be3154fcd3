#!/usr/bin/env python
"""A few launches of one tap-GEMM shape for `ncu --set full --import-source on` captures.
usage: prof_one.py M K N [act] [residual]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops
M, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
res = len(sys.argv) > 5 and sys.argv[5] == "1"
dev = "cuda"
x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(); b = torch.randn(N, device=dev)
r = torch.randn(M, N // 2 if act == 2 else N, device=dev).bfloat16() if res else None
out = torch.empty(M, N // 2 if act == 2 else N, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.linear(x, w, b, act=act, residual=r, out=out)
torch.cuda.synchronize()
print("done")
