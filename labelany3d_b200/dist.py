"""Images shard across the GPUs of one node; one all-gather of the packed box records at the end.

The reference scales by running independent processes over disjoint image ranges
(``--start_index/--end_index/--gpu_idx``, ``src/batch_scripts/whole.py:25-27`` of the reference) and
merging JSON files afterwards.  Here: one process per GPU (``torch.distributed``, NCCL over
NVLink / NVSwitch), rank ``r`` owns a contiguous block of images, nothing is exchanged on the data
path, and the ``[B/G, I, 64]`` record tensors are all-gathered once so every rank holds all boxes.
The per-image RNG seed (``seed + global image index``) makes the result independent of ``G``.
"""

from __future__ import annotations

import ctypes

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
    """``BoxFitter`` for this rank's block of a ``B_total``-image batch plus the gather of the records.

    ``collective="p2p"`` (default on CUDA with more than one rank): the gathered ``[world*per, I, 64]``
    buffer of every rank lives in symmetric (peer-mapped) memory and the fit kernel of rank ``r``
    writes its records directly into slot ``r`` of EVERY rank's buffer over NVLink
    (``la3d_fit_boxes_p2p``); a flag barrier over peer memory (``la3d_peer_barrier``) then tells each
    rank that all slots have landed.  No separate collective pass, no NCCL launch on the data path.
    The gathered buffer is double-buffered: the tensor a call returns stays valid until the
    next-but-one call (consume it on the same stream).
    ``collective="nccl"``: one ``all_gather_into_tensor`` after the fit (the plain form).
    """

    def __init__(self, B_total, I, H, W, device=None, out_dtype=torch.float64, group=None, collective=None):
        from .ops import BoxFitter
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B_total = int(B_total)
        self.start, self.stop, self.per = shard_range(self.B_total, self.rank, self.world)
        self.n_local = self.stop - self.start
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.out_dtype = out_dtype
        self.I = int(I)
        self.local = BoxFitter(max(self.n_local, 1), I, H, W, device=device, out_dtype=out_dtype)
        if collective is None:
            collective = "p2p" if self.world > 1 else "none"
        if collective not in ("p2p", "nccl", "none"):
            raise ValueError(f"collective must be 'p2p', 'nccl' or 'none', got {collective!r}")
        if self.world == 1:
            collective = "none"
        self.collective = collective
        if collective == "p2p":
            self._init_p2p()
        else:
            self.gathered = torch.empty((self.world * self.per, I, REC), dtype=out_dtype, device=device)
            self.slot = torch.full((self.per, I, REC), float("nan"), dtype=out_dtype, device=device)
            self.slot[..., O_STATUS] = -1

    # -- peer memory ---------------------------------------------------------------------------
    def _init_p2p(self):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if self.world > 8:
            raise ValueError("the peer-memory gather supports at most 8 ranks (one NVLink node)")
        grp = self.group if self.group is not None else dist.group.WORLD
        shape = (self.world * self.per, self.I, REC)
        self._bufs, self._peer_slots = [], []
        esize = torch.empty((), dtype=self.out_dtype).element_size()
        slot_bytes = self.per * self.I * REC * esize
        for _ in range(2):                                   # double buffer (see the class docstring)
            t = symm.empty(shape, dtype=self.out_dtype, device=self.device)
            t.fill_(float("nan"))
            t[..., O_STATUS] = -1
            hdl = symm.rendezvous(t, grp)
            self._bufs.append((t, hdl))
            self._peer_slots.append([int(p) + self.rank * slot_bytes for p in hdl.buffer_ptrs])
        f = symm.empty((64,), dtype=torch.int32, device=self.device)
        f.zero_()
        fh = symm.rendezvous(f, grp)
        self._flags, self._flag_hdl = f, fh
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in fh.buffer_ptrs])
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._side = torch.cuda.Stream(device=self.device)
        self._ev_fit = torch.cuda.Event()
        self._ev_bar = [torch.cuda.Event(), torch.cuda.Event()]
        self._epoch = 0
        self._lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                       # every rank's fills have landed before the first step

    def wait_gathered(self):
        """Make the current stream wait until the records of the last call have landed on every rank."""
        if self.collective == "p2p" and self._epoch:
            torch.cuda.current_stream(self.device).wait_event(self._ev_bar[self._epoch & 1])

    def check_barrier_status(self):
        """Raises if a peer failed to arrive at a barrier (synchronises the device)."""
        if self.collective == "p2p" and int(self._status.item()) != 0:
            raise RuntimeError("la3d_peer_barrier: a peer did not arrive within the timeout")

    def __call__(self, depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, events=None, wait=True):
        """Inputs are this rank's block (``[n_local, ...]``).  Returns ``[B_total, I, 64]`` on every rank.
        ``wait=False`` (p2p only): do not make the current stream wait for the peer barrier; the result is
        complete once :meth:`wait_gathered` (or a device synchronisation) has passed."""
        if self.collective == "p2p":
            from . import _lib
            self._epoch += 1
            e = self._epoch
            buf, _ = self._bufs[e & 1]
            cur = torch.cuda.current_stream(self.device)
            if self.n_local:
                # the fit may write the peers' buffer e&1 once every rank is past step e-1 (previous barrier)
                self.local(depth, K, masks, ground, method, yaw_steps, seed, image_offset=self.start,
                           peers=self._peer_slots[e & 1], wait_before_fit=self._ev_bar[(e - 1) & 1] if e > 1 else None)
            elif e > 1:
                cur.wait_event(self._ev_bar[(e - 1) & 1])
            self._ev_fit.record(cur)
            self._side.wait_event(self._ev_fit)
            with torch.cuda.device(self.device):
                rc = self._lib.la3d_peer_barrier(self._flag_ptrs, self.rank, self.world, e & 0xFFFFFFFF,
                                                 self._status.data_ptr(), self._side.cuda_stream)
            _lib.check(rc, "la3d_peer_barrier")
            self._ev_bar[e & 1].record(self._side)
            if wait:
                cur.wait_event(self._ev_bar[e & 1])
            return buf[:self.B_total]
        if self.n_local:
            self.local(depth, K, masks, ground, method, yaw_steps, seed, image_offset=self.start,
                       out=self.slot[:self.n_local], events=events)
        if self.world == 1:
            return self.slot[:self.B_total]
        dist.all_gather_into_tensor(self.gathered, self.slot, group=self.group)
        return self.gathered[:self.B_total]
