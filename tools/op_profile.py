#!/usr/bin/env python
"""Per-op GPU time of encode / one U-Net step / decode at 320x512x16f using CUDA events around every C-ABI op
(eager, so launch gaps are included in 'wall' but not in the per-op sums)."""
import collections, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops, synthetic

REC = []
def wrap(name, fn):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        desc = name
        if name in ("linear", "conv3x3", "temporal_conv3", "bmm_nt"):
            x = a[0]; w = a[1] if name == "linear" else (a[4] if name == "conv3x3" else (a[4] if name == "temporal_conv3" else a[1]))
            desc = f"{name} M={x.shape[0] if x.dim()==2 else tuple(x.shape)} K={x.shape[-1]} N={w.shape[-2]}"
        elif name == "groupnorm":
            desc = f"groupnorm rows={a[0].shape[0]} C={a[0].shape[1]} stats={a[1]}"
        elif name == "attention":
            desc = f"attention B={a[4]} H={a[5]} Lq={a[6]} Lk={a[7]}"
        REC.append((desc, e0, e1))
        return r
    return inner

for n in ["linear", "conv3x3", "temporal_conv3", "bmm_nt", "groupnorm", "layernorm", "attention", "temporal_attention",
          "bcthw_to_rows", "rows_to_bcthw", "concat_rows", "upsample2x", "im2col_s2", "softmax_rows", "transpose_bf16"]:
    setattr(ops, n, wrap(n, getattr(ops, n)))

def report(title, top=18):
    torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d, e0, e1 in REC:
        agg[d][0] += 1; agg[d][1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    print(f"== {title}: {len(REC)} ops, sum {tot:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {v[1]:8.3f} ms  n={v[0]:4d}  {k}")
    REC.clear()

def main():
    dev = torch.device("cuda")
    H, W = 320, 512
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    video = synthetic.synthetic_video(16, H, W, device=dev)
    unet = model.model.diffusion_model
    for rep in range(2):
        REC.clear(); torch.cuda.synchronize(); t0 = time.time()
        z = model.encode_first_stage(video)
        torch.cuda.synchronize(); wall = time.time() - t0
    report(f"encode 16 frames (wall {wall*1e3:.1f} ms)")
    x = torch.randn(1, 20, 16, H // 8, W // 8, device=dev)
    ctx = torch.cat([model.get_learned_conditioning([""]), model.get_image_conditioning(1)], 1)
    ts = torch.tensor([499], device=dev)
    for rep in range(2):
        REC.clear(); torch.cuda.synchronize(); t0 = time.time()
        y = unet(x, ts, context=ctx, fs=torch.tensor([24], device=dev))
        torch.cuda.synchronize(); wall = time.time() - t0
    report(f"one U-Net step eager (wall {wall*1e3:.1f} ms)", top=40)
    zz = torch.randn(16, 4, H // 8, W // 8, device=dev)
    for rep in range(2):
        REC.clear(); torch.cuda.synchronize(); t0 = time.time()
        d = pm_vae.decode_with_conf_adaptor(zz)
        torch.cuda.synchronize(); wall = time.time() - t0
    report(f"decode+conf 16 frames (wall {wall*1e3:.1f} ms)")

if __name__ == "__main__":
    main()
