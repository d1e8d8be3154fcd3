#!/usr/bin/env python
"""Evaluation entry point with the reference's flag surface (scripts/evaluation/infer_geo4d.py:688-718 /
scripts/eval_geo4d.sh) on the B200-native path: for every sequence of a dataset -> sliding 16-frame windows ->
DDIM + decode per window -> global alignment -> the repo's own depth and pose metrics and result files.

    python scripts/evaluation/infer_geo4d.py --config configs/inference_geo4d.yaml --ckpt_path model.ckpt \
        --dataset folder:/data/my_seqs --savedir results --height 320 --width 512 --ddim_steps 5 --stride 8

What is written (same names as the reference's run_evaluation, :314-647): per sequence `pred_traj.txt` (TUM,
wxyz), `pred_focal.txt`, `pred_intrinsics.txt`, `frame_%04d.npy` + `frame_colordepth_%04d.png` +
`colored_depth_maps.gif`, `conf_%d.npy`, `init_conf_%d.npy`, `frame_%04d.png`, `{seq}_error_%d.png`,
`_error_log_depth.txt`, `{seq}_eval_metric.txt`, `_error_log.txt`; per run `_error_log_all.txt`, `time_cost.txt`.
Not reproduced: the GLB export (trimesh, `get_3D_model_from_scene`) and the trajectory plot (evo/matplotlib).

Datasets.  The reference reads Sintel / Bonn / KITTI / ... through lvdm/data/eval_dataset_geo4d.py; none of
them (nor a checkpoint) exists offline, so this script accepts
  --dataset synthetic[:n_seq[:n_frames]]   seeded synthetic clips with synthetic ground truth (plumbing / timing)
  --dataset folder:<dir>                   <dir>/<seq>/{rgb/*.png|*.jpg, depth/*.npy (optional), traj.txt (optional, TUM wxyz)}
and `--synthetic_weights` runs seeded random weights when no --ckpt_path is given.
"""
import argparse
import datetime
import glob
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def get_parser():
    """same options, defaults and help as the reference parser (infer_geo4d.py:688-718)"""
    p = argparse.ArgumentParser()
    p.add_argument("--savedir", type=str, default=None, help="results saving path")
    p.add_argument("--ckpt_path", type=str, default=None, help="checkpoint path")
    p.add_argument("--config", type=str, default=os.path.join(REPO, "configs", "inference_geo4d.yaml"), help="config (yaml) path")
    p.add_argument("--prompt_dir", type=str, default=None, help="a data dir containing videos and prompts")
    p.add_argument("--n_samples", type=int, default=1, help="num of samples per prompt")
    p.add_argument("--ddim_steps", type=int, default=50, help="steps of ddim if positive, otherwise use DDPM")
    p.add_argument("--ddim_eta", type=float, default=1.0, help="eta for ddim sampling (0.0 yields deterministic sampling)")
    p.add_argument("--bs", type=int, default=1, help="batch size for inference, should be one")
    p.add_argument("--height", type=int, default=512, help="image height, in pixel space")
    p.add_argument("--width", type=int, default=512, help="image width, in pixel space")
    p.add_argument("--frame_stride", type=int, default=3, help="frame stride control")
    p.add_argument("--unconditional_guidance_scale", type=float, default=1.0, help="prompt classifier-free guidance")
    p.add_argument("--seed", type=int, default=123, help="seed for seed_everything")
    p.add_argument("--video_length", type=int, default=16, help="inference video length")
    p.add_argument("--negative_prompt", action="store_true", default=False, help="negative prompt")
    p.add_argument("--text_input", action="store_true", default=False, help="input text to I2V model or not")
    p.add_argument("--multiple_cond_cfg", action="store_true", default=False, help="use multi-condition cfg or not")
    p.add_argument("--cfg_img", type=float, default=None, help="guidance scale for image conditioning")
    p.add_argument("--timestep_spacing", type=str, default="uniform", help="timestep spacing (uniform | uniform_trailing)")
    p.add_argument("--guidance_rescale", type=float, default=0.0, help="guidance rescale")
    p.add_argument("--perframe_ae", action="store_true", default=False, help="per-frame AE decoding (accepted; frames are batched here)")
    p.add_argument("--dataset", type=str, default=None, help="Evaluation Dataset")
    p.add_argument("--full_seq", action="store_true", default=False, help="Evaluation Dataset")
    p.add_argument("--stride", type=int, default=4, help="Sliding window stride for video")
    p.add_argument("--loop", action="store_true", default=False, help="generate looping videos or not")
    p.add_argument("--interp", action="store_true", default=False, help="generate generative frame interpolation or not")
    # additions of this port (no checkpoints / datasets offline)
    p.add_argument("--synthetic_weights", action="store_true", help="seeded random weights when no checkpoint is given")
    p.add_argument("--cond_path", type=str, default=None, help="cached conditioning tensors {'text','img'} (see INTEGRATION.md)")
    return p


# ----------------------------------------------------------------------------------------------- datasets
def synthetic_sequences(spec, H, W, seed):
    """seeded clips with a synthetic ground truth (smooth positive depth, smooth camera path): plumbing only"""
    import numpy as np
    import torch
    from geo4d_b200 import synthetic
    parts = spec.split(":")
    n_seq = int(parts[1]) if len(parts) > 1 else 1
    T = int(parts[2]) if len(parts) > 2 else 24
    for s in range(n_seq):
        video = synthetic.synthetic_video(T, H, W, device="cpu", seed=seed + s)
        ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        depth = torch.stack([3.0 + 1.5 * torch.sin(xs / W * 3.1 + 0.1 * t) * torch.cos(ys / H * 2.3) for t in range(T)])
        traj = np.zeros((T, 7))
        traj[:, 0] = 0.02 * np.arange(T)
        traj[:, 1] = 0.01 * np.sin(0.3 * np.arange(T))          # not collinear: the Sim(3) alignment of ATE needs rank >= 2
        traj[:, 2] = 0.01 * np.arange(T)
        traj[:, 3] = 1.0                                        # wxyz identity
        yield {"seq": f"synthetic_{s:02d}", "video": video, "depth": depth, "gt_traj": [traj, np.arange(T).astype(float)]}


def folder_sequences(root, H, W, max_frames=None):
    import cv2
    import numpy as np
    import torch
    from geo4d_b200 import metrics
    for seq_dir in sorted(d for d in glob.glob(os.path.join(root, "*")) if os.path.isdir(d)):
        files = sorted(glob.glob(os.path.join(seq_dir, "rgb", "*.png")) + glob.glob(os.path.join(seq_dir, "rgb", "*.jpg")))
        if max_frames:
            files = files[:max_frames]
        if len(files) < 16:
            continue
        frames = [cv2.cvtColor(cv2.resize(cv2.imread(f), (W, H), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2RGB) for f in files]
        video = (torch.from_numpy(np.stack(frames)).float().permute(3, 0, 1, 2) / 127.5 - 1.0).unsqueeze(0)
        item = {"seq": os.path.basename(seq_dir), "video": video}
        dfiles = sorted(glob.glob(os.path.join(seq_dir, "depth", "*.npy")))[:len(files)]
        if len(dfiles) == len(files):
            item["depth"] = torch.from_numpy(np.stack([np.load(f) for f in dfiles])).float()
        tfile = os.path.join(seq_dir, "traj.txt")
        if os.path.exists(tfile):
            item["gt_traj"] = metrics.load_tum_trajectory(tfile)
        yield item


def load_model(args, dev):
    import torch
    from geo4d_b200 import synthetic
    from geo4d_b200.config import instantiate_from_config, load_yaml
    cfg = load_yaml(args.config)
    if args.ckpt_path is None:
        if not args.synthetic_weights:
            raise SystemExit("Error: checkpoint Not Found! (pass --ckpt_path, or --synthetic_weights for seeded random weights)")
        return synthetic.build_model(args.config, device=dev, seed=max(args.seed, 0))
    assert os.path.exists(args.ckpt_path), "Error: checkpoint Not Found!"
    model = instantiate_from_config(cfg["model"]).to(dev)
    sd = torch.load(args.ckpt_path, map_location="cpu")
    model.load_state_dict(sd["state_dict"] if "state_dict" in sd else sd, strict=True)
    pm_vae = None
    if "vae_path" in cfg:
        pm_vae = instantiate_from_config(cfg["pointmap_vae_config"]).to(dev)
        vsd = torch.load(cfg["vae_path"], map_location="cpu")["state_dict"]
        pm_vae.load_state_dict({k[6:]: v for k, v in vsd.items() if k.startswith("model.")}, strict=True)   # :343-347
        pm_vae.prepare()
    model.prepare()
    if args.cond_path is None:
        raise SystemExit("--cond_path is required with a real checkpoint (the OpenCLIP towers are outside this port)")
    c = torch.load(args.cond_path, map_location=dev)
    model.set_cached_conditioning(c["text"], c["img"])
    return model, pm_vae, cfg


def run_evaluation(args):
    import cv2
    import numpy as np
    import torch
    import torch.nn.functional as F
    from geo4d_b200 import metrics
    from geo4d_b200.pipeline import Geo4DPipeline
    assert (args.height % 16 == 0) and (args.width % 16 == 0), "Error: image size [h,w] should be multiples of 16!"
    assert args.bs == 1, "Current implementation only support [batch size = 1]!"
    if args.loop or args.interp:
        raise NotImplementedError
    dev = torch.device("cuda")
    model, pm_vae, cfg = load_model(args, dev)
    model.perframe_ae = args.perframe_ae
    post = dict(cfg.get("postprocess") or {})
    post["use_gt_focal"] = False                                                     # :371
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=args.ddim_steps, ddim_eta=args.ddim_eta,
                         unconditional_guidance_scale=args.unconditional_guidance_scale,
                         timestep_spacing=args.timestep_spacing, guidance_rescale=args.guidance_rescale, postprocess=post,
                         seed=max(args.seed, 0), multiple_cond_cfg=args.multiple_cond_cfg, cfg_img=args.cfg_img)
    name = args.dataset or "synthetic"
    if name.startswith("folder:"):
        seqs, dataset_name = folder_sequences(name[len("folder:"):], args.height, args.width), "folder"
    elif name.startswith("synthetic"):
        seqs, dataset_name = synthetic_sequences(name, args.height, args.width, max(args.seed, 0)), "synthetic"
    else:
        raise SystemExit(f"dataset {name!r}: the reference's dataset loaders need data that is not available offline; "
                         "use synthetic[:n[:T]] or folder:<dir>")
    orivae = pm_vae is None
    save_dir = os.path.join(args.savedir or "results", f"{dataset_name}" +
                            f"raydir_cross_depth_seq_stride{args.stride}_cameraopt_rot1.0_depth2_ddimstep{args.ddim_steps}"
                            f"_ddimeta{args.ddim_eta}_fastlr001_temp{post.get('temporal_smoothing_weight')}"
                            f"_cfg{args.unconditional_guidance_scale}_same_time_orivae{orivae}_robustfocal_gtfocalFalse_clean")
    os.makedirs(save_dir, exist_ok=True)
    ate_list, rpe_trans_list, rpe_rot_list, depth_metrics, time_list = [], [], [], [], []
    total_frames = 0
    for batch in seqs:
        seq = batch["seq"]
        video = batch["video"].to(dev)
        T = video.shape[2]
        total_frames += T
        torch.cuda.synchronize()
        t0 = time.time()                                         # :437,462,503-510: diffusion + decode + alignment
        scene, preds = pipe.reconstruct(video, stride=args.stride, keep_images=True)
        torch.cuda.synchronize()
        time_list.append(time.time() - t0)
        print(f"Diffusion + Optimization time: {time_list[-1]:.2f}s")
        out = f"{save_dir}/{seq}"
        os.makedirs(out, exist_ok=True)
        depthmap = torch.stack(scene.get_depthmaps(), 0)
        if batch.get("depth") is not None:
            gt = batch["depth"].to(dev).float()
            OH, OW = gt.shape[-2:]
            dm = F.interpolate(depthmap[None], size=(OH, OW), mode="bicubic", align_corners=False, antialias=True)[0].detach()
            # validity of every frame = "not masked (sky / far) in any window that saw it" (:476-484)
            vm = torch.ones(T, args.height, args.width, device=dev)
            for sl, p in zip(pipe.last_windows, preds):
                vm[sl] = torch.minimum(vm[sl], (p["conf"][..., 0] > 0).float())
            cm = (F.interpolate(vm[None], size=(OH, OW), mode="bicubic", align_corners=False, antialias=True)[0] > 0.8).reshape(-1)
            if dataset_name == "kitti":
                res, err, _, _ = metrics.depth_evaluation(dm.reshape(-1), gt.reshape(-1), max_depth=None, align_with_lad2=True)
            else:
                res, err, _, _ = metrics.depth_evaluation(dm.reshape(-1), gt.reshape(-1), max_depth=70, align_with_lad2=True,
                                                          post_clip_max=70, lr=1e-2, max_iters=5000, align_mask=cm)
            err = err.reshape(T, OH, OW)
            for i in range(T):
                cv2.imwrite(os.path.join(out, f"{seq}_error_{i}.png"),
                            np.clip(err[i].detach().cpu().numpy() * 255, 0, 255).astype(np.uint8))
            print(res)
            depth_metrics.append(res)
            with open(f"{out}/_error_log_depth.txt", "a") as f:
                f.write(f"{seq}_{res}\n")
        pred_traj = scene.get_tum_poses()
        scene.save_tum_poses(f"{out}/pred_traj.txt")
        scene.save_focals(f"{out}/pred_focal.txt")
        scene.save_intrinsics(f"{out}/pred_intrinsics.txt")
        scene.save_depth_maps(out)
        scene.save_conf_maps(out)
        scene.save_init_conf_maps(out)
        if scene.imgs is not None:
            scene.save_rgb_imgs(out)
        if batch.get("gt_traj") is not None:
            try:
                ate, rpe_trans, rpe_rot = metrics.eval_metrics(pred_traj, batch["gt_traj"], seq=seq,
                                                               filename=f"{save_dir}/{seq}_eval_metric.txt", sample_stride=1)
            except Exception as e:   # the reference logs zeros for a failed sequence (:588-593)
                print(f"Error: {e}")
                ate, rpe_trans, rpe_rot = 0, 0, 0
            ate_list.append(ate); rpe_trans_list.append(rpe_trans); rpe_rot_list.append(rpe_rot)
            with open(f"{out}/_error_log.txt", "a") as f:
                f.write(f"{post.get('eval_dataset')}-{seq: <16} | ATE: {ate:.5f}, RPE trans: {rpe_trans:.5f}, RPE rot: {rpe_rot:.5f}\n")
                f.write(f"{ate:.5f}\n{rpe_trans:.5f}\n{rpe_rot:.5f}\n")
            print(f"ATE: {ate:.5f}, RPE trans: {rpe_trans:.5f}, RPE rot: {rpe_rot:.5f}")
    if depth_metrics:
        avg = metrics.average_depth_metrics(depth_metrics)
        print("Average depth evaluation metrics:", avg)
        with open(f"{save_dir}/_error_log_all.txt", "a") as f:
            f.write(f"Average depth evaluation metrics: {avg}\n")
    if ate_list:
        nz = lambda v: float(np.asarray(v)[np.nonzero(np.asarray(v))].mean()) if np.any(np.asarray(v)) else 0.0
        print(f"ATE: {nz(ate_list)}, rpe_trans: {nz(rpe_trans_list)}, rpe_rot: {nz(rpe_rot_list)}")
        with open(f"{save_dir}/_error_log_all.txt", "a") as f:
            f.write(f"ATE: {nz(ate_list)}, rpe_trans: {nz(rpe_trans_list)}, rpe_rot: {nz(rpe_rot_list)}")
    tl = np.array(time_list)
    per_frame = tl.sum() / max(total_frames, 1)
    print("time_list", tl); print("total_times", tl.sum()); print("time_for_each_frames", per_frame)
    with open(f"{save_dir}/time_cost.txt", "a") as f:                                  # :640-648
        f.write(f"total_times: {tl.sum()}\n")
        f.write(f"time_for_each_frames: {per_frame}\n")
        f.write(f"time_list: {tl}\n")
    return save_dir


if __name__ == "__main__":
    print("@Geo4D cond-Inference: %s" % datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S"))
    a = get_parser().parse_args()
    if a.seed < 0:
        import random
        a.seed = random.randint(0, 2 ** 31)
    import torch
    torch.manual_seed(a.seed)
    run_evaluation(a)
