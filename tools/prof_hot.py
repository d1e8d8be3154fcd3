#!/usr/bin/env python
"""One launch of every hot kernel at the 320x512x16f shapes, in a fixed order, for `ncu --set full` captures:

    ncu --set full --clock-control none --import-source on -k regex:'attn_fwd|tap_gemm|gn_fused|align_loop|splitk' \
        -o gpurun_out/r2_hot python tools/prof_hot.py

Order (= launch index in the report): attention L=2560 self, L=640 self, L=160 self, text+image cross (77 + 16 keys)
at L=2560; tap-GEMM conv3x3 320->320 @40x64, conv3x3 640->640 @20x32, conv3x3 1280->1280 @5x8 (split-K + its reduce),
linear 40960x320->320, GEGLU 320->2560; GroupNorm+SiLU 16 x 2560 x 320; 3 iterations of the alignment loop (1 window).
GEO4D_AUTOTUNE is off here (every tuning launch would be profiled): tiles come from the library's cost model."""
import os
import sys
os.environ.setdefault("GEO4D_AUTOTUNE", "0")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops


def main():
    dev = torch.device("cuda")
    torch.manual_seed(0)
    bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
    M = 40960
    qkv = bf(M, 960); o = torch.empty(M, 320, device=dev, dtype=torch.bfloat16)
    ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], o, 16, 5, 2560, 2560)
    q1 = bf(10240, 1920); o1 = torch.empty(10240, 640, device=dev, dtype=torch.bfloat16)
    ops.attention(q1[:, :640], q1[:, 640:1280], q1[:, 1280:], o1, 16, 10, 640, 640)
    q2 = bf(2560, 3840); o2 = torch.empty(2560, 1280, device=dev, dtype=torch.bfloat16)
    ops.attention(q2[:, :1280], q2[:, 1280:2560], q2[:, 2560:], o2, 16, 20, 160, 160)
    kvt, kvi = bf(77, 640), bf(16 * 16, 640)
    ops.cross_attention2(qkv[:, :320], kvt[:, :320], kvt[:, 320:], 77, 16, kvi[:, :320], kvi[:, 320:], 16, 1, o, 16, 5, 2560)
    x = bf(M, 320)
    ops.conv3x3(x, 16, 40, 64, bf(9, 320, 320), torch.randn(320, device=dev))
    x1 = bf(10240, 640)
    ops.conv3x3(x1, 16, 20, 32, bf(9, 640, 640), torch.randn(640, device=dev))
    ops.conv3x3(bf(640, 1280), 16, 5, 8, bf(9, 1280, 1280), torch.randn(1280, device=dev))
    ops.linear(x, bf(320, 320), torch.randn(320, device=dev))
    ops.linear(x, bf(2560, 320), torch.randn(2560, device=dev), act=ops.ACT_GEGLU)
    ops.groupnorm(x, 16, 2560, torch.ones(320, device=dev), torch.zeros(320, device=dev), 1e-5, True)
    # alignment loop: one window, 3 iterations in one launch
    from geo4d_b200.cloud_opt import LightPointCloudGroupOptimizer
    T, H, W = 16, 320, 512
    pred = {"pts3d": torch.randn(T, H, W, 3, device=dev) + torch.tensor([0.0, 0.0, 4.0], device=dev),
            "conf": 1 + torch.rand(T, H, W, 1, device=dev), "inverse_depthmap": 0.1 + torch.rand(T, H, W, 1, device=dev),
            "traj": torch.eye(4, device=dev).repeat(T, 1, 1)}
    sc = LightPointCloudGroupOptimizer([[{"idx": (i,)} for i in range(T)]], [pred], conf="id", conf_optimize=True,
                                       verbose=False, shared_focal=True, num_total_iter=3, temporal_smoothing_weight=0.015,
                                       translation_weight=1.0, depth_traj_start_iter=3, shard_alignment=False, engine="loop")
    with torch.no_grad():
        sc.im_depthmaps.fill_(1.4)
    sc._global_alignment_loop(lr=0.03, niter=3, schedule="linear", lr_min=1e-3)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
