#!/usr/bin/env python
"""DDIM loop timing at 16f 320x512 (CUDA-graph path), S steps; prints ms per U-Net step and a checksum."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import synthetic
from geo4d_b200.pipeline import Geo4DPipeline

def main():
    H, W, S = 320, 512, int(os.environ.get("DDIM", 20))
    dev = torch.device("cuda")
    if os.environ.get("POISON") == "1":
        # fill the caching allocator's pool with bf16 NaNs: any read of memory the pipeline never wrote shows up
        junk = [torch.full((1 << 28,), 0x7FC07FC0, dtype=torch.int32, device=dev) for _ in range(60)]
        del junk
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=S, postprocess=dict(cfg["postprocess"]))
    video = synthetic.synthetic_video(16, H, W, device=dev)
    def bits(t):
        v = t.contiguous().view(torch.int32).to(torch.int64)
        return int(v.sum().item()) ^ int((v * 31 % 1000003).sum().item())
    z = model.encode_first_stage(video)
    print("bits video", bits(video), "z", bits(z.float()), "cpu-rng probe", float(torch.randn(1)), flush=True)
    g = torch.Generator(device=dev).manual_seed(123)
    x_T = torch.randn((1, 16, 16, H // 8, W // 8), device=dev, generator=g)
    cond = {"c_crossattn": [torch.cat([model.get_learned_conditioning([""]), model.get_image_conditioning(1)], 1)], "c_concat": [z]}
    fs = torch.tensor([24], device=dev)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        samples, _ = pipe.sampler.sample(S=S, conditioning=cond, batch_size=1, shape=(16, 16, H // 8, W // 8), verbose=False, eta=0.0, x_T=x_T, fs=fs, timestep_spacing="uniform_trailing")
        e1.record(); torch.cuda.synchronize()
        print(f"PDL={os.environ.get('GEO4D_PDL', '1')} rep{rep}: {e0.elapsed_time(e1) / S:.3f} ms/step  checksum {float(samples.double().abs().mean()):.6f} bits {bits(samples.float())} nan {int(torch.isnan(samples).sum())}", flush=True)

if __name__ == "__main__":
    main()
