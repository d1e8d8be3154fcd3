#!/usr/bin/env python
"""Headline benchmark: 4D-reconstruction frames/sec, 320x512x16f windows, 50-step DDIM, synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full pass of the hot path over a synthetic clip: per window VAE-encode of the 16
conditioning frames -> 50 DDIM steps of the spatio-temporal U-Net (CUDA graph) -> 4 VAE decodes (point map +
confidence, ray directions, ray moments, inverse depth) -> per-window post-processing -> sliding-window global
alignment (init + 500 fused iterations + LAD / trajectory sub-alignments).  At N = 1 this is BASELINE.json
configs[1] (one 16-frame window); at N > 1 rank r owns window r of a 8(N+1)-frame clip (stride 8), the
per-window predictions are all-gathered over NCCL and the global alignment runs replicated (weak scaling).

`value` is timed with the video already in HBM; `e2e` includes the pinned-host -> device copy of the video
and the device -> host read of depth maps / poses / focal every step.  `--impl reference` times the CPU
restatement of the reference (oracle/, fp32 PyTorch, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# algorithmic work, traced from the reference modules, 2*MAC
UNET_TFLOP = {(320, 512): 12.61, (256, 256): 4.91, (576, 1024): 52.36}
VAE_TFLOP = {(320, 512): dict(dec=1.564, dec_conf=1.757, enc=0.690)}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"],
                    source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def unet_step_traffic():
    """DRAM bytes (read + write) of the roofline unit from the committed ncu pass, or None."""
    p = os.path.join(REPO, "profiles", "r1_unet_step_traffic.json")
    try:
        return json.load(open(p))["dram_bytes"]
    except Exception:
        return None


def kernel_rooflines(dev, pk):
    """Live per-kernel numbers for the dominant kernel (tap_gemm_kernel) on its heaviest U-Net shapes and for the
    attention kernel: each launch is timed GPU-bound (20 launches in a CUDA graph, CUDA events on the launching
    stream) and set against the measured dense-bf16 burst peak (a kernel timed alone)."""
    import torch
    from geo4d_b200 import ops

    def gtime(fn, n=20):
        fn()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with ops.capture_graph(g):
                for _ in range(n):
                    fn()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(3):
                g.replay()
            e1.record(side)
            e1.synchronize()
        torch.cuda.current_stream().wait_stream(side)
        return e0.elapsed_time(e1) * 1e-3 / (3 * n)

    bf = lambda *sh: torch.randn(*sh, device=dev).bfloat16()
    out = []

    def add(name, flops, sec, launches_per_step):
        ach = flops / sec * 1e-12
        out.append({"kernel": name, "us": round(sec * 1e6, 2), "achieved": round(ach, 1), "unit": "TFLOP/s",
                    "peak": pk["bf16_burst"], "frac": round(ach / pk["bf16_burst"], 3),
                    "launches_per_unet_step": launches_per_step})

    x = bf(40960, 320); w9 = bf(9, 320, 320); b = torch.randn(320, device=dev); o = torch.empty(40960, 320, device=dev, dtype=torch.bfloat16)
    add("tap_gemm conv3x3 16x40x64 320->320", 2.0 * 40960 * 2880 * 320, gtime(lambda: ops.conv3x3(x, 16, 40, 64, w9, b, out=o)), 7)
    x1 = bf(10240, 640); w1 = bf(9, 640, 640); b1 = torch.randn(640, device=dev); o1 = torch.empty(10240, 640, device=dev, dtype=torch.bfloat16)
    add("tap_gemm conv3x3 16x20x32 640->640", 2.0 * 10240 * 5760 * 640, gtime(lambda: ops.conv3x3(x1, 16, 20, 32, w1, b1, out=o1)), 6)
    wl = bf(320, 320)
    add("tap_gemm linear 40960x320->320 (+bias)", 2.0 * 40960 * 320 * 320, gtime(lambda: ops.linear(x, wl, b, out=o)), 45)
    wg = bf(2560, 320); bg = torch.randn(2560, device=dev); og = torch.empty(40960, 1280, device=dev, dtype=torch.bfloat16)
    add("tap_gemm linear 40960x320->2560 GEGLU", 2.0 * 40960 * 320 * 2560, gtime(lambda: ops.linear(x, wg, bg, act=ops.ACT_GEGLU, out=og)), 10)
    qkv = bf(40960, 960)
    add("attn_fwd B16 H5 L2560 d64", 4.0 * 16 * 5 * 2560 * 2560 * 64,
        gtime(lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], o, 16, 5, 2560, 2560)), 5)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active") and not v.lower().startswith("not"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """CPU restatement of the reference (oracle port), bounded sample, extrapolated to the workload."""
    if rank != 0:
        return
    import torch
    from oracle import unet as ou, vae as ov
    H, W = args.height, args.width
    # threads actually used: the sample's operators are small (one frame at 128x256); on the 128-thread GPU hosts an
    # OpenMP team of 128 spends its time in fork/join (measured: ~15 busy cores, > 100 s per sample), so cap at 32
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    # Bounded sample (~20-30 s of CPU work): ONE frame of the 16-frame window at 128x256 through one U-Net step and
    # each VAE pass.  Every stage's cost is linear in frames and (to first order) in pixels -- the spatial
    # self-attention is quadratic, so extrapolating linearly to HxW favours the CPU arm -- hence the factors below.
    t_sample, Hs, Ws = 1, 128, 256
    px = (H * W) / float(Hs * Ws)
    cfg = ou.UNetConfig(temporal_length=t_sample)
    g = torch.Generator().manual_seed(0)
    torch.set_flush_denormal(True)   # timing must not depend on denormal slow paths

    _base = torch.randn(1 << 25, generator=torch.Generator().manual_seed(7))
    _scaled = {}

    def cheap_params(shapes):
        # 1.4e9 weights: a seeded randn fill of every tensor costs minutes of single-threaded RNG and 5.6 GB of page
        # faults.  Timing does not depend on the values as long as activations stay well scaled, so every weight is
        # a read-only VIEW into one 32M-sample normal block pre-scaled by 2^-k/2 with 2^k ~ fan_in (no copies);
        # norm gains are ones, biases zeros
        out = {}
        off = 0
        for name, shp in shapes.items():
            n = 1
            for d in shp:
                n *= int(d)
            if len(shp) == 1:
                out[name] = torch.ones(shp) if name.endswith("weight") else torch.zeros(shp)
                continue
            fan = max(1, n // max(1, int(shp[0])))
            k = max(0, int(round(__import__("math").log2(fan))))
            if k not in _scaled:
                _scaled[k] = _base * (2.0 ** (-k / 2))
            blk = _scaled[k]
            assert n <= blk.numel(), (name, shp)
            off = (off + 7919 * 4) % (blk.numel() - n + 1)
            out[name] = blk[off:off + n].view(shp)
        return out
    sd = cheap_params(ou.param_shapes(cfg))
    x = torch.randn(1, 20, t_sample, Hs // 8, Ws // 8, generator=g)
    ctx = torch.randn(1, 77 + 16 * t_sample, 1024, generator=g)
    ts = torch.tensor([499])
    vcfg = ov.VAEConfig()
    vsd = cheap_params(ov.param_shapes(vcfg))
    z = torch.randn(1, 4, Hs // 8, Ws // 8, generator=g)
    img = torch.randn(1, 3, Hs, Ws, generator=g)

    def sample_once():
        with torch.no_grad():
            t0 = time.time(); ou.forward(cfg, sd, x, ts, ctx, None); t_unet = (time.time() - t0) * (16 / t_sample) * px
            t0 = time.time(); ov.decode_with_conf_adaptor(vcfg, vsd, z); t_dc = (time.time() - t0) * px
            t0 = time.time(); ov.decode(vcfg, vsd, z); t_d = (time.time() - t0) * px
            t0 = time.time(); ov.encode_moments(vcfg, vsd, img); t_e = (time.time() - t0) * px
        return t_unet, t_dc, t_d, t_e

    for _ in range(min(args.warmup, 1)):
        sample_once()
    acc = [sample_once() for _ in range(max(1, min(args.steps, 2)))]
    t_unet, t_dc, t_d, t_e = [sum(a[i] for a in acc) / len(acc) for i in range(4)]
    window_s = args.ddim_steps * t_unet + 16 * (t_dc + 3 * t_d) + 16 * t_e
    n_windows = max(1, world)
    frames = 16 if world <= 1 else 8 * (world + 1)
    value = frames / (window_s * n_windows)
    sample = (f"1 U-Net step + 1 decode+conf + 1 plain decode + 1 encode of ONE frame at {Hs}x{Ws}, fp32 PyTorch on "
              f"{cores} threads; extrapolated linearly in pixels to {H}x{W} and to {args.ddim_steps} steps x 16 frames x "
              f"{n_windows} window(s); alignment excluded")
    line = {"impl": "reference", "metric": "4D-recon frames/sec", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": window_s * n_windows * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{frames}f {H}x{W}, {args.ddim_steps}-step DDIM, {n_windows} window(s)",
                       "extrapolated": True},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "phases_s": {"unet_step": t_unet, "decode_conf_frame": t_dc, "decode_frame": t_d, "encode_frame": t_e}}
    print(json.dumps(line))
    return line


# ----------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--align-iters", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from geo4d_b200 import ops, sharding, synthetic
    from geo4d_b200.pipeline import Geo4DPipeline, sliding_windows
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W = args.height, args.width
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    pipe = Geo4DPipeline(model, pm_vae, ddim_steps=args.ddim_steps, postprocess=dict(cfg["postprocess"], silent=True,
                                                                                     n_iter=args.align_iters))
    T = 16 if world == 1 else 8 * (world + 1)
    windows = sliding_windows(T, 8)
    assert len(windows) == world
    video_host = synthetic.synthetic_video(T, H, W, device="cpu", seed=123).pin_memory()
    video_dev = video_host.to(dev, non_blocking=True)
    my = windows[rank]

    def step(video):
        """this rank's window -> predictions -> (all-gather) -> global alignment; returns the scene"""
        _, preds = pipe.reconstruct(video[:, :, my], stride=8, windows=[slice(0, 16, 1)], align=False,
                                    x_T_fn=lambda wi: torch.randn((1, 16, 16, H // 8, W // 8), device=dev,
                                                                  generator=torch.Generator(device=dev).manual_seed(123 + rank)))
        preds = sharding.gather_predictions({rank: preds[0]}, world, 16, H, W)
        views = [[{"idx": (i,)} for i in range(w.start, w.stop)] for w in windows]
        with torch.enable_grad():
            scene = pipe.post_optimization(views, preds)
        return scene

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(video_dev)
    # ---- timed region 1: device-resident input
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pipe.events = []
    e0.record()
    for _ in range(args.steps):
        scene = step(video_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - n0
    phase_ms = pipe.phase_ms()
    # ---- timed region 2: end to end from pinned host memory, results read back
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    for _ in range(args.steps):
        vd = video_host.to(dev, non_blocking=True)
        sc = step(vd)
        outs = [torch.stack(sc.get_depthmaps()).cpu(), sc.get_im_poses().detach().cpu(), sc.get_focals().detach().cpu()]
        d2h = sum(o.numel() * o.element_size() for o in outs)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    unet_ms = phase_ms.get("ddim", 0.0) / max(1, args.steps * args.ddim_steps)
    tflop = UNET_TFLOP.get((H, W))
    roof = None
    if tflop and unet_ms > 0:
        ach = tflop / (unet_ms * 1e-3)
        roof = {"bound": "tensor", "kernel": "U-Net step (1 CUDA-graph launch; tap_gemm_kernel = 93% of its FLOPs)",
                "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                "peak_source": pk["source"] + " (sustained: timed inside a long step)", "traffic": unet_step_traffic(),
                "algorithmic_tflop_per_launch": tflop, "ms_per_launch": unet_ms}
        try:
            roof["kernels"] = kernel_rooflines(dev, pk)
        except Exception as ex:  # pragma: no cover
            roof["kernels"] = {"error": repr(ex)}
    value = T * args.steps / (ms * 1e-3)
    line = {"metric": "4D-recon frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{T}f {H}x{W}, {args.ddim_steps}-step DDIM (cfg 1, eta 0, uniform_trailing), "
                                   f"{len(windows)} window(s) stride 8, {args.align_iters}-iter alignment",
                       "l2": "weights 2.9 GB + activations >> 126 MB L2 (no flush needed)",
                       "weights": "seeded synthetic", "parallelism": f"window-parallel x{world}",
                       "gemm_autotune": f"{len(ops.tuned_configs())} shapes pinned during warm-up"},
            "e2e": {"value": T * args.steps / (ms_e2e * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": video_host.numel() * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk, "roofline": roof,
            "phases_ms_per_step": {k: v / args.steps for k, v in phase_ms.items()}}
    if not args.no_cpu_baseline:
        try:
            old = sys.stdout
            sys.stdout = open(os.devnull, "w")
            ref = run_reference(argparse.Namespace(**{**vars(args), "steps": 1, "warmup": 0}), 0, world)
            sys.stdout = old
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as ex:  # pragma: no cover
            sys.stdout = old
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
