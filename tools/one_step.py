#!/usr/bin/env python
"""One eager U-Net step at 16f 320x512 between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import synthetic

def main():
    dev = torch.device("cuda")
    H, W = 320, 512
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    unet = model.model.diffusion_model
    x = torch.randn(1, 20, 16, H // 8, W // 8, device=dev)
    ctx = torch.cat([model.get_learned_conditioning([""]), model.get_image_conditioning(1)], 1)
    ts = torch.tensor([499], device=dev)
    fs = torch.tensor([24], device=dev)
    for _ in range(2):
        y = unet(x, ts, context=ctx, fs=fs)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = unet(x, ts, context=ctx, fs=fs)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", float(y.abs().mean()))

if __name__ == "__main__":
    main()
