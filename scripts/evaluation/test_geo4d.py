#!/usr/bin/env python
"""Demo entry point with the reference's flag surface (scripts/evaluation/test_geo4d.py:571-604 /
scripts/infer_geo4d.sh): video -> 4D reconstruction on the B200-native path.

    python scripts/evaluation/test_geo4d.py --config configs/inference_geo4d.yaml \
        --ckpt_path checkpoints/geo4d/model.ckpt --video_path data/demo/drift-turn.mp4 --savedir results \
        --height 320 --width 512 --ddim_steps 5 --stride 4

Without a checkpoint (none exists offline) pass --synthetic to run seeded random weights, e.g. for timing.
The OpenCLIP conditioning towers are outside this port: pass --cond_path with the constant conditioning
tensors saved by the reference ({'text': [1,77,1024], 'img': [1,256,1024]}); see INTEGRATION.md.
"""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def get_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--savedir", type=str, default="results", help="results saving path")
    p.add_argument("--ckpt_path", type=str, default=None, help="checkpoint path (model.ckpt)")
    p.add_argument("--config", type=str, default=os.path.join(REPO, "configs", "inference_geo4d.yaml"))
    p.add_argument("--video_path", type=str, default=None)
    p.add_argument("--cond_path", type=str, default=None, help="cached conditioning tensors (text, img)")
    p.add_argument("--synthetic", action="store_true", help="seeded random weights / conditioning / video")
    p.add_argument("--n_samples", type=int, default=1)
    p.add_argument("--ddim_steps", type=int, default=5)
    p.add_argument("--ddim_eta", type=float, default=0.0)
    p.add_argument("--bs", type=int, default=1)
    p.add_argument("--height", type=int, default=320)
    p.add_argument("--width", type=int, default=512)
    p.add_argument("--unconditional_guidance_scale", type=float, default=1.0)
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--video_length", type=int, default=16)
    p.add_argument("--stride", type=int, default=4)
    p.add_argument("--timestep_spacing", type=str, default="uniform_trailing")
    p.add_argument("--guidance_rescale", type=float, default=0.0)
    p.add_argument("--perframe_ae", action="store_true", default=True)
    p.add_argument("--max_frames", type=int, default=64)
    return p


def load_video(path, H, W, max_frames):
    import cv2
    import numpy as np
    import torch
    cap = cv2.VideoCapture(path)
    frames = []
    while len(frames) < max_frames:
        ok, fr = cap.read()
        if not ok:
            break
        fr = cv2.cvtColor(cv2.resize(fr, (W, H), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2RGB)
        frames.append(fr)
    v = torch.from_numpy(np.stack(frames)).float().permute(3, 0, 1, 2) / 127.5 - 1.0  # c t h w in [-1, 1]
    return v.unsqueeze(0)


def main():
    args = get_parser().parse_args()
    import torch
    from geo4d_b200 import synthetic
    from geo4d_b200.config import instantiate_from_config, load_yaml
    from geo4d_b200.pipeline import Geo4DPipeline
    assert args.height % 16 == 0 and args.width % 16 == 0, "Error: image size [h,w] should be multiples of 16!"
    assert args.bs == 1, "Current implementation only support [batch size = 1]!"
    torch.manual_seed(args.seed)
    dev = torch.device("cuda")
    if args.synthetic or args.ckpt_path is None:
        model, pm_vae, cfg = synthetic.build_model(args.config, device=dev, seed=args.seed)
    else:
        cfg = load_yaml(args.config)
        model = instantiate_from_config(cfg["model"]).to(dev)
        sd = torch.load(args.ckpt_path, map_location="cpu")
        model.load_state_dict(sd["state_dict"] if "state_dict" in sd else sd, strict=True)
        pm_vae = instantiate_from_config(cfg["pointmap_vae_config"]).to(dev)
        vsd = torch.load(cfg["vae_path"], map_location="cpu")["state_dict"]
        # infer_geo4d.py:343-347: only the keys under 'model.' belong to the fine-tuned VAE, the rest is dropped
        pm_vae.load_state_dict({k[6:]: v for k, v in vsd.items() if k.startswith("model.")}, strict=True)
        model.prepare(); pm_vae.prepare()
        if args.cond_path is None:
            raise SystemExit("--cond_path is required with a real checkpoint (conditioning towers are out of scope)")
        c = torch.load(args.cond_path, map_location=dev)
        model.set_cached_conditioning(c["text"], c["img"])
    if args.video_path and os.path.exists(args.video_path):
        video = load_video(args.video_path, args.height, args.width, args.max_frames).to(dev)
    else:
        video = synthetic.synthetic_video(max(args.video_length, 16), args.height, args.width, device=dev,
                                          seed=args.seed)
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=args.ddim_steps, ddim_eta=args.ddim_eta,
                         unconditional_guidance_scale=args.unconditional_guidance_scale,
                         timestep_spacing=args.timestep_spacing, guidance_rescale=args.guidance_rescale,
                         postprocess=cfg.get("postprocess"), seed=args.seed)
    t0 = time.time()
    scene, _ = pipe.reconstruct(video, stride=args.stride)
    torch.cuda.synchronize()
    dt = time.time() - t0
    os.makedirs(args.savedir, exist_ok=True)
    scene.save_tum_poses(os.path.join(args.savedir, "pred_traj.txt"))
    scene.save_focals(os.path.join(args.savedir, "pred_focal.txt"))
    scene.save_intrinsics(os.path.join(args.savedir, "pred_intrinsics.txt"))
    scene.save_depth_maps(args.savedir)
    with open(os.path.join(args.savedir, "time_cost.txt"), "w") as f:  # infer_geo4d.py:640-648
        f.write(f"time_for_each_frames: {dt / video.shape[2]}\n")
    print(f"{video.shape[2]} frames in {dt:.2f}s -> {video.shape[2] / dt:.2f} frames/s; results in {args.savedir}")


if __name__ == "__main__":
    main()
