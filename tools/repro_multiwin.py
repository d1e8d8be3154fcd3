"""Multi-window reconstruction on ONE GPU (C3-like): T frames, stride 8, full alignment.  Prints per-phase
times and the alignment profile; any exception is printed with its traceback (used to reproduce the
round-1 N = 8 crash of the replicated alignment at G = 8 windows without paying for 8 GPUs).

    GEO4D_ALIGN_PROFILE=1 python tools/repro_multiwin.py --frames 72 --ddim-steps 4 --runs 2
"""
import argparse
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=72)
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--ddim-steps", type=int, default=4)
    ap.add_argument("--align-iters", type=int, default=500)
    ap.add_argument("--runs", type=int, default=2)
    args = ap.parse_args()
    import torch
    from geo4d_b200 import synthetic
    from geo4d_b200.pipeline import Geo4DPipeline, sliding_windows
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=args.ddim_steps,
                         postprocess=dict(cfg["postprocess"], silent=True, n_iter=args.align_iters))
    T, H, W = args.frames, args.height, args.width
    video = synthetic.synthetic_video(T, H, W, device=dev, seed=123)
    print("windows:", sliding_windows(T, 8), flush=True)
    for r in range(args.runs):
        pipe.events = []
        torch.cuda.synchronize()
        t0 = time.time()
        try:
            scene, preds = pipe.reconstruct(video, stride=8)
            torch.cuda.synchronize()
            dt = time.time() - t0
            dm = torch.stack(scene.get_depthmaps())
            print(f"run {r}: {dt:.3f} s  ({T / dt:.2f} frames/s)  phases(ms)={ {k: round(v, 1) for k, v in pipe.phase_ms().items()} }"
                  f"  depth finite={bool(torch.isfinite(dm).all())} focal={float(scene.get_focals()[0]):.2f}"
                  f" invalid_depth={scene.invalid_depth_group} valid_traj={scene.valid_traj_group_list}", flush=True)
        except Exception:
            traceback.print_exc()
            print("REPRO: exception above", flush=True)
            return 1
    print(f"max memory allocated: {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
