"""Window sharding across ranks: windows are independent through diffusion + decode +
per-window post-processing; the only exchange is one all-gather of the fixed-size per-window record
{pts3d 16 HW 3, conf 16 HW, inverse depth 16 HW, traj 16x16} before the global alignment.  One process per
GPU; NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

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


# ----------------------------------------------------------------------------------------------- alignment sharding
def partition_images(edges_per_image: Sequence[int], world: int) -> List[int]:
    """Contiguous image ranges per rank for the sharded alignment loop, balanced by per-pixel work
    (1 + number of windows observing the image).  Returns img_lo with world + 1 entries."""
    cost = [1 + int(e) for e in edges_per_image]
    n, total = len(cost), float(sum(cost))
    lo, acc, r = [0], 0.0, 1
    for i, c in enumerate(cost):
        acc += c
        while r < world and acc >= total * r / world - 1e-9 and (n - (i + 1)) >= 0:
            lo.append(i + 1)
            r += 1
    while len(lo) < world + 1:
        lo.append(n)
    lo[world] = n
    return lo


class PeerExchange:
    """One receive buffer per rank, mapped into every rank's address space so that a kernel can store straight
    into its peers' memory over NVLink (the exchange step of geo4d_align_loop).  Mapping goes through
    torch.distributed._symmetric_memory (CUDA VMM handles) and, if that is unavailable, through classic CUDA IPC
    handles (torch storage sharing).  `ptrs[r]` is the address of rank r's buffer as seen from THIS process."""

    def __init__(self, nbytes: int, device: torch.device, group=None):
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = nbytes
        self.how = None
        self._keep = []
        err = []
        for how in ("symm", "ipc"):
            try:
                getattr(self, "_open_" + how)(nbytes, device, group)
                self.how = how
                break
            except Exception as ex:  # noqa: BLE001 -- any failure selects the next transport
                err.append(f"{how}: {type(ex).__name__}: {ex}")
        ok = torch.tensor([1 if self.how else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0 or self.how is None:
            raise RuntimeError("no peer-memory transport available: " + " | ".join(err))
        self.local.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def _open_symm(self, nbytes, device, group):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
        self.local = t
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        self._keep.append(hdl)
        assert self.ptrs[self.rank] == t.data_ptr()

    def _open_ipc(self, nbytes, device, group):
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        handle = t.untyped_storage()._share_cuda_()
        off = t.storage_offset()
        handles = [None] * self.world
        dist.all_gather_object(handles, (handle, off), group=group)
        self.local = t
        self.ptrs = []
        for r, (h, o) in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(t.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*h)
            peer = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st, o, (nbytes,))
            torch.zeros(8, dtype=torch.uint8, device=device).copy_(peer[:8])   # makes torch enable P2P access
            self._keep.append((st, peer))
            self.ptrs.append(peer.data_ptr())


_exchange_cache: Dict[tuple, "PeerExchange"] = {}
_flag_epoch = 0


def peer_exchange(nbytes: int, device: torch.device, group=None) -> Optional["PeerExchange"]:
    """Cached PeerExchange of at least nbytes, or None when no transport works on this box (the caller then runs
    the alignment replicated).  Collective: every rank must call it with the same size."""
    key = (device.index, dist.get_world_size(group), nbytes)
    if key not in _exchange_cache:
        try:
            _exchange_cache[key] = PeerExchange(nbytes, device, group)
        except Exception as ex:  # noqa: BLE001
            import warnings
            warnings.warn(f"geo4d_b200: sharded alignment disabled ({ex})")
            _exchange_cache[key] = None
    return _exchange_cache[key]


def reserve_flags(count: int) -> int:
    """Monotonic flag values for the next `count` iterations (identical on every rank: same call sequence)."""
    global _flag_epoch
    base = _flag_epoch
    _flag_epoch += count + 8
    return base
