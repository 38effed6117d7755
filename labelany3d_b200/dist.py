"""Images shard across the GPUs of one node; one all-gather of the packed box records at the end.

The reference scales by running independent processes over disjoint image ranges
(``--start_index/--end_index/--gpu_idx``, ``src/batch_scripts/whole.py:25-27`` of the reference) and
merging JSON files afterwards.  Here: one process per GPU (``torch.distributed``, NCCL over
NVLink / NVSwitch), rank ``r`` owns a contiguous block of images, nothing is exchanged on the data
path, and the ``[B/G, I, 64]`` record tensors are all-gathered once so every rank holds all boxes.
The per-image RNG seed (``seed + global image index``) makes the result independent of ``G``.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from .records import O_STATUS, REC


def shard_range(total, rank, world):
    """Contiguous block ``[start, stop)`` of rank ``rank``; every rank gets ``ceil(total/world)`` slots,
    the last ranks may own fewer (or no) real images."""
    per = (total + world - 1) // world
    start = min(rank * per, total)
    return start, min(start + per, total), per


def all_gather_records(local, total=None, group=None):
    """``local[n_r, I, 64]`` on every rank -> ``[total, I, 64]`` on every rank (rank order = image order).

    Ranks may hold different ``n_r`` only through :func:`shard_range` padding: ``local`` is padded to the
    common slot count with ``status = -1`` rows, gathered with one ``all_gather_into_tensor`` (NCCL
    all-gather on CUDA tensors, gloo on CPU tensors) and trimmed to ``total``.
    """
    world = dist.get_world_size(group)
    n, I, rec = local.shape
    assert rec == REC
    if total is None:
        total = n * world
    per = (total + world - 1) // world
    if n != per:
        pad = torch.full((per, I, REC), float("nan"), dtype=local.dtype, device=local.device)
        pad[..., O_STATUS] = -1
        pad[:n] = local
        local = pad
    out = torch.empty((world * per, I, REC), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:total]


class ShardedBoxFitter:
    """``BoxFitter`` for this rank's block of a ``B_total``-image batch plus the final all-gather."""

    def __init__(self, B_total, I, H, W, device=None, out_dtype=torch.float64, group=None):
        from .ops import BoxFitter
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B_total = int(B_total)
        self.start, self.stop, self.per = shard_range(self.B_total, self.rank, self.world)
        self.n_local = self.stop - self.start
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.local = BoxFitter(max(self.n_local, 1), I, H, W, device=device, out_dtype=out_dtype)
        self.gathered = torch.empty((self.world * self.per, I, REC), dtype=out_dtype, device=device)
        self.slot = torch.full((self.per, I, REC), float("nan"), dtype=out_dtype, device=device)
        self.slot[..., O_STATUS] = -1

    def __call__(self, depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, events=None):
        """Inputs are this rank's block (``[n_local, ...]``).  Returns ``[B_total, I, 64]`` on every rank."""
        if self.n_local:
            self.local(depth, K, masks, ground, method, yaw_steps, seed, image_offset=self.start,
                       out=self.slot[:self.n_local], events=events)
        if self.world == 1:
            return self.slot[:self.B_total]
        dist.all_gather_into_tensor(self.gathered, self.slot, group=self.group)
        return self.gathered[:self.B_total]
