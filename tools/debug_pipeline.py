#!/usr/bin/env python
"""Full-size single-window pipeline on the GPU with per-stage statistics and timings (debug aid)."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import synthetic, ops
from geo4d_b200.pipeline import Geo4DPipeline

def stats(name, t):
    t = t.float()
    print(f"{name:28s} shape={tuple(t.shape)} nan={int(torch.isnan(t).sum())} inf={int(torch.isinf(t).sum())} "
          f"min={float(t.nan_to_num().min()):.4g} max={float(t.nan_to_num().max()):.4g} rms={float(t.nan_to_num().pow(2).mean().sqrt()):.4g}", flush=True)

def main():
    H, W = int(os.environ.get("H", 320)), int(os.environ.get("W", 512))
    steps = int(os.environ.get("DDIM", 50))
    dev = torch.device("cuda")
    t0 = time.time()
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    torch.cuda.synchronize(); print("build_model s", time.time() - t0, "mem GB", torch.cuda.memory_allocated() / 1e9, flush=True)
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=steps, postprocess=dict(cfg["postprocess"], silent=False))
    video = synthetic.synthetic_video(16, H, W, device=dev)
    stats("video", video)
    z = model.encode_first_stage(video)
    stats("cond latent z", z)
    g = torch.Generator(device=dev).manual_seed(123)
    x_T = torch.randn((1, 16, 16, H // 8, W // 8), device=dev, generator=g)
    cond = {"c_crossattn": [torch.cat([model.get_learned_conditioning([""]), model.get_image_conditioning(1)], 1)], "c_concat": [z]}
    fs = torch.tensor([24], device=dev)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        samples, inter = pipe.sampler.sample(S=steps, conditioning=cond, batch_size=1, shape=(16, 16, H // 8, W // 8), verbose=False, eta=0.0, x_T=x_T, fs=fs, timestep_spacing="uniform_trailing")
        torch.cuda.synchronize(); print(f"sample rep{rep} s", time.time() - t0, flush=True)
    stats("samples", samples)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        maps = pipe.decode_latents(samples)
        torch.cuda.synchronize(); print(f"decode rep{rep} s", time.time() - t0, "peak mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
    for c, n in enumerate(["x", "y", "z", "conf_raw", "rdx", "rdy", "rdz", "rmx", "rmy", "rmz", "invd"]):
        stats("map " + n, maps[:, c])
    pred = pipe.window_predictions(maps)
    for k, v in pred.items():
        stats("pred " + k, v)
    print("valid frac", float((pred["conf"] > 0.5).float().mean()), flush=True)
    views = [[{"idx": (i,)} for i in range(16)]]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        with torch.enable_grad():
            scene = pipe.post_optimization(views, [pred])
        torch.cuda.synchronize(); print(f"align rep{rep} s", time.time() - t0, flush=True)
    stats("depth", torch.stack(scene.get_depthmaps()))
    print("focal", float(scene.get_focals()[0]), "launches", ops.launch_count(), flush=True)

if __name__ == "__main__":
    main()
