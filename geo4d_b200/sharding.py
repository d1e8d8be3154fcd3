"""Window sharding across ranks: windows are independent through diffusion + decode +
per-window post-processing; the only exchange is one all-gather of the fixed-size per-window record
{pts3d 16 HW 3, conf 16 HW, inverse depth 16 HW, traj 16x16} before the global alignment.  One process per
GPU; NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def windows_for_rank(n_windows: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment w -> rank w mod world (contiguous when n_windows == world)."""
    return [w for w in range(n_windows) if w % world == rank]


def record_numel(T: int, H: int, W: int) -> int:
    return T * H * W * 5 + T * 16


def pack_pred(pred: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([pred["pts3d"].reshape(-1), pred["conf"].reshape(-1), pred["inverse_depthmap"].reshape(-1),
                      pred["traj"].reshape(-1)]).float().contiguous()


def unpack_pred(flat: torch.Tensor, T: int, H: int, W: int) -> Dict[str, torch.Tensor]:
    n = T * H * W
    return {"pts3d": flat[:3 * n].view(T, H, W, 3), "conf": flat[3 * n:4 * n].view(T, H, W, 1),
            "inverse_depthmap": flat[4 * n:5 * n].view(T, H, W, 1), "traj": flat[5 * n:5 * n + T * 16].view(T, 4, 4)}


def gather_predictions(local: Dict[int, Dict[str, torch.Tensor]], n_windows: int, T: int, H: int, W: int,
                       group=None) -> List[Dict[str, torch.Tensor]]:
    """local: {window index: pred} owned by this rank -> list of all n_windows preds on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per_rank = -(-n_windows // world)
    rec = record_numel(T, H, W)
    any_pred = next(iter(local.values()))
    dev = any_pred["pts3d"].device
    send = torch.zeros(per_rank * rec, device=dev)
    mine = windows_for_rank(n_windows, rank, world)
    for slot, w in enumerate(mine):
        send[slot * rec:(slot + 1) * rec] = pack_pred(local[w])
    if world == 1:
        recv = send
    else:
        recv = torch.empty(world * per_rank * rec, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
    out: List[Dict[str, torch.Tensor]] = [None] * n_windows
    for r in range(world):
        for slot, w in enumerate(windows_for_rank(n_windows, r, world)):
            off = (r * per_rank + slot) * rec
            out[w] = unpack_pred(recv[off:off + rec], T, H, W)
    return out


def gather_group_records(local: torch.Tensor, n_groups: int, group=None) -> torch.Tensor:
    """Per-window results computed round-robin over the ranks -> the full table on every rank.
    local [ceil(n_groups / world), k]: row j holds this rank's j-th window (windows_for_rank order), unused rows
    are ignored.  Returns [n_groups, k] in window order, identical on every rank (used for the LAD scale / shift
    fits of the replicated global alignment)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = -(-n_groups // world)
    assert local.shape[0] == per, (local.shape, per)
    if world == 1:
        return local[:n_groups].clone()
    allrec = torch.empty((world * per,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(allrec, local.contiguous(), group=group)
    allrec = allrec.view((world, per) + tuple(local.shape[1:]))
    return torch.stack([allrec[g % world, g // world] for g in range(n_groups)])
