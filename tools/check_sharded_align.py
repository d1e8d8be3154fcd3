"""Multi-GPU check of the sharded alignment loop (torchrun, one process per GPU, NCCL):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded_align.py [--big]

Every rank builds the same seeded synthetic scene (oracle.align.synthetic_scene: test infrastructure), runs
the alignment (a) replicated (every rank optimises the whole clip, shard_alignment=False) and (b) sharded
(images split over the ranks, gradient records exchanged inside the kernel over NVLink) and checks that
(1) the sharded result is bit-identical on every rank, (2) it agrees with the replicated one to fp32
rounding of the reductions, (3) the loop survives repeated calls (monotonic flags, cached peer buffers).
Prints timings of both variants (CUDA events, max over ranks).  Exit code 0 = all checks passed.
"""
import argparse
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also time a 72-frame 320x512 clip (8 windows)")
    ap.add_argument("--iters", type=int, default=60)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from oracle import align as oa
    from geo4d_b200.cloud_opt import LightPointCloudGroupOptimizer
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def run(T, H, W, niter, start_b, lad, shard, repeat=1):
        groups, preds, _ = oa.synthetic_scene(T=T, H=H, W=W, noise=0.003)
        views = [[{"idx": (i,)} for i in g] for g in groups]
        preds_d = [{k: v.to(dev) for k, v in p.items()} for p in preds]
        best, sc = None, None
        for _ in range(repeat):
            sc = LightPointCloudGroupOptimizer(views, preds_d, conf="id", conf_optimize=True, verbose=False,
                                               shared_focal=True, num_total_iter=niter, temporal_smoothing_weight=0.015,
                                               translation_weight=1.0, depth_traj_start_iter=start_b, lad_max_iters=lad,
                                               engine="loop", shard_alignment=shard)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.enable_grad():
                loss = sc.compute_global_alignment(init="group", niter=niter, schedule="linear", lr=0.03)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return sc, loss, best

    try:
        rep, l_rep, t_rep = run(40, 32, 48, args.iters, 20, 300, shard=False)
        shd, l_shd, t_shd = run(40, 32, 48, args.iters, 20, 300, shard=True, repeat=3)
        info = shd._shard
        if rank == 0:
            print(f"small scene: replicated {t_rep:.1f} ms, sharded {t_shd:.1f} ms, shard info {info}", flush=True)
        if not info or info["world"] != world:
            print(f"rank {rank}: alignment was NOT sharded ({info})", flush=True)
            ok = False
        names = ["im_poses", "im_focals", "pw_poses", "s_depth", "t_depth", "traj_align_poses", "im_depthmaps"]
        for n in names:
            a = getattr(shd, n).detach()
            ref = a.clone()
            dist.broadcast(ref, src=0)
            if not torch.equal(a, ref):
                print(f"rank {rank}: {n} differs from rank 0 (max {float((a - ref).abs().max()):.3e})", flush=True)
                ok = False
        # phase A only: trajectories comparable parameter by parameter
        rep_a, _, _ = run(40, 32, 48, 20, 20, 300, shard=False)
        shd_a, _, _ = run(40, 32, 48, 20, 20, 300, shard=True)
        for n in ("im_poses", "im_focals", "pw_poses", "im_depthmaps"):
            d = float((getattr(rep_a, n).detach() - getattr(shd_a, n).detach()).abs().max())
            if not d < 2e-4:
                print(f"rank {rank}: sharded vs replicated {n}: {d:.3e}", flush=True)
                ok = False
        da, db = torch.stack(rep.get_depthmaps()), torch.stack(shd.get_depthmaps())
        rel = float(((da - db).abs() / da).mean())
        if not (rel < 2e-2 and abs(l_rep - l_shd) / l_rep < 0.1):
            print(f"rank {rank}: full run differs: depth {rel:.3e}, loss {l_rep} vs {l_shd}", flush=True)
            ok = False
        if args.big:
            _, _, t1 = run(72, 320, 512, 500, 150, 5000, shard=False, repeat=2)
            _, _, t2 = run(72, 320, 512, 500, 150, 5000, shard=True, repeat=2)
            if rank == 0:
                print(f"72 frames 320x512, 8 windows, 500 iterations: replicated {t1:.1f} ms, sharded x{world} {t2:.1f} ms",
                      flush=True)
    except Exception:
        traceback.print_exc()
        ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_ALIGN_OK" if int(flag) else "SHARDED_ALIGN_FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag) else 1


if __name__ == "__main__":
    sys.exit(main())
