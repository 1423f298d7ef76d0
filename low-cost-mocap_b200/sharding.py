"""Frame-set sharding across GPUs.

S1-S3 have no cross-frame state (the first temporal state in the reference is the Kalman
filter, helpers.py:109), so the stream shards by frame-set index: frame-set f belongs to
rank ``f % world`` (round-robin keeps every rank in lock-step with a live stream).  No
collective sits on the data path; the only exchange is ONE all-gather of fixed-size track
records per batch so that every rank (and the sequential consumer on rank 0: locate_objects /
Kalman, helpers.py:107-109) sees all 3D tracks in frame order.

One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_frame_sets: int, rank: int, world: int) -> np.ndarray:
    """Global frame-set indices owned by ``rank`` (round-robin)."""
    return np.arange(rank, n_frame_sets, world)


def shard_size(n_frame_sets: int, rank: int, world: int) -> int:
    return (n_frame_sets - rank + world - 1) // world if n_frame_sets > rank else 0


def pack_tracks(obj, err, n):
    """[B,R,3] f64, [B,R] f64, [B] i32  ->  one [B, R*4+1] f64 record tensor (x,y,z,err per root, count last).
    A single dense record per frame-set keeps the exchange one collective."""
    import torch
    B, R = err.shape
    rec = torch.empty((B, R * 4 + 1), dtype=torch.float64, device=obj.device)
    rec[:, : R * 3] = obj.reshape(B, R * 3)
    rec[:, R * 3: R * 4] = err
    rec[:, R * 4] = n.to(torch.float64)
    return rec


def unpack_tracks(rec, R):
    import torch
    B = rec.shape[0]
    obj = rec[:, : R * 3].reshape(B, R, 3)
    err = rec[:, R * 3: R * 4]
    n = rec[:, R * 4].to(torch.int32)
    return obj, err, n


def all_gather_tracks(rec_local, n_total: int, group=None):
    """One all-gather of the local track records; returns the records of ALL frame-sets in
    global frame order [n_total, rec].  Ranks may own unequal shard sizes (n_total % world != 0):
    shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = rec_local.shape[1]
    per = (n_total + world - 1) // world
    if rec_local.shape[0] != shard_size(n_total, rank, world):
        raise ValueError("local shard has the wrong number of frame-sets")
    if rec_local.shape[0] < per:
        pad = torch.zeros((per - rec_local.shape[0], width), dtype=rec_local.dtype, device=rec_local.device)
        rec_local = torch.cat([rec_local, pad], dim=0)
    gathered = torch.empty((world, per, width), dtype=rec_local.dtype, device=rec_local.device)
    dist.all_gather_into_tensor(gathered.view(world * per, width), rec_local.contiguous(), group=group)
    # frame f lives at [f % world, f // world]  ->  transpose restores global order
    ordered = gathered.permute(1, 0, 2).reshape(per * world, width)
    return ordered[:n_total]


class TrackBuffer:
    """Tracks of one rank's shard in ONE flat device allocation, laid out so that the matcher writes
    straight into the all-gather send buffer (no pack kernel) and the gathered result is read through
    views (no reorder copy):  [ obj f64 B*R*3 | err f64 B*R | n i32 B | flags i32 B ]."""

    def __init__(self, n_sets: int, max_roots: int, device):
        import torch
        self.B, self.R = n_sets, max_roots
        self.sizes = [n_sets * max_roots * 3 * 8, n_sets * max_roots * 8, n_sets * 4, n_sets * 4]
        self.offsets = [0]
        for sz in self.sizes:
            self.offsets.append((self.offsets[-1] + sz + 15) // 16 * 16)
        self.nbytes = self.offsets[-1]
        self.flat = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.views = self._views(self.flat)
        self.gathered = None

    def _views(self, flat):
        import torch
        o, B, R = self.offsets, self.B, self.R
        return {"obj": flat[o[0]: o[0] + self.sizes[0]].view(torch.float64).view(B, R, 3),
                "err": flat[o[1]: o[1] + self.sizes[1]].view(torch.float64).view(B, R),
                "n": flat[o[2]: o[2] + self.sizes[2]].view(torch.int32),
                "flags": flat[o[3]: o[3] + self.sizes[3]].view(torch.int32)}

    def all_gather(self, group=None, async_op=False):
        """ONE collective over the raw bytes; returns per-rank views [world] of dicts like ``views``.
        Global frame-set f lives at rank f % world, local index f // world (round-robin ownership).

        With ``async_op=True`` returns ``(views, work)``: the collective is only enqueued (after the work
        already on the current stream), so the next batch's kernels can run while the tracks travel; call
        ``work.wait()`` before reading the views or writing this buffer again (use two buffers in turn).

        Every rank must hold a buffer of the SAME size (``all_gather_into_tensor``): size the buffers for
        ``ceil(n_total / world)`` frame-sets on every rank and ignore the padding frame-sets of the short shards
        (their ``n`` stays 0); ``all_gather_tracks`` above does that padding itself for ragged totals."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if self.gathered is None or self.gathered.shape[0] != world:
            self.gathered = torch.empty((world, self.nbytes), dtype=torch.uint8, device=self.flat.device)
        work = dist.all_gather_into_tensor(self.gathered.view(-1), self.flat, group=group, async_op=async_op)
        views = [self._views(self.gathered[r]) for r in range(world)]
        return (views, work) if async_op else views
