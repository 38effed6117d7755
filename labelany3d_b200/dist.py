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


class PeerGather:
    """The gathered record buffers of this rank in peer-mapped (symmetric) memory plus the flag rows: what a
    box kernel needs to write its records into every rank's buffer and to synchronise with the peers itself
    (``la3d_sink`` of include/la3d.h).  torch's symmetric memory is used for allocation and the handle
    exchange only.  Two gathered buffers alternate: the tensor a step returns stays valid until the
    next-but-one step (consume it on the stream that issues the steps)."""

    def __init__(self, per, I, world, rank, device, out_dtype, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if world > _lib.MAX_PEERS:
            raise ValueError("the peer-memory gather supports at most 8 ranks (one NVLink node)")
        self._lib_mod = _lib
        self._lib = _lib.load()
        self.world, self.rank, self.per, self.I = int(world), int(rank), int(per), int(I)
        self.device, self.out_dtype = torch.device(device), out_dtype
        grp = group if group is not None else dist.group.WORLD
        shape = (self.world * self.per, self.I, REC)
        esize = torch.empty((), dtype=out_dtype).element_size()
        slot_bytes = self.per * self.I * REC * esize
        self._bufs, self._peer_slots = [], []
        for _ in range(2):
            t = symm.empty(shape, dtype=out_dtype, device=self.device)
            t.fill_(float("nan"))
            t[..., O_STATUS] = -1
            hdl = symm.rendezvous(t, grp)
            self._bufs.append((t, hdl))
            self._peer_slots.append([int(p) + self.rank * slot_bytes for p in hdl.buffer_ptrs])
        f = symm.empty((64,), dtype=torch.int32, device=self.device)
        f.zero_()
        fh = symm.rendezvous(f, grp)
        self._flags, self._flag_hdl = f, fh
        self._flag_list = [int(p) for p in fh.buffer_ptrs]
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*self._flag_list)
        self._counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        # sticky error word in pinned host memory: readable without a device synchronisation, even after a trap
        self._status = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                            # every rank's fills have landed before the first step

    def next_sink(self):
        """Advance the epoch; returns ``(sink, gathered_buffer)`` of the new step."""
        self.epoch += 1
        e = self.epoch
        sink = self._lib_mod.make_sink(self._peer_slots[e & 1], self.out_dtype == torch.float64, self._flag_list,
                                       self._counter.data_ptr(), self._status.data_ptr(), e, self.rank)
        return sink, self._bufs[e & 1][0]

    def signal(self):
        """For a rank without images in this step: publish the epoch without a fit."""
        with torch.cuda.device(self.device):
            rc = self._lib.la3d_peer_signal(self._flag_ptrs, self.rank, self.world, self.epoch & 0xFFFFFFFF,
                                            torch.cuda.current_stream().cuda_stream)
        self._lib_mod.check(rc, "la3d_peer_signal")

    def wait(self):
        """The current stream publishes this rank's epoch (the box kernel leaves that to the next launch on the stream)
        and waits until the records of the last step have landed from every rank."""
        if self.epoch:
            with torch.cuda.device(self.device):
                rc = self._lib.la3d_peer_barrier(self._flag_ptrs, self.rank, self.world, self.epoch & 0xFFFFFFFF,
                                                 self._status.data_ptr(), torch.cuda.current_stream().cuda_stream)
            self._lib_mod.check(rc, "la3d_peer_barrier")

    def check(self):
        """Raises if a peer failed to arrive within the timeout (no device synchronisation: the word is in
        pinned host memory and is written before the kernel traps)."""
        if int(self._status[0]) != 0:
            raise RuntimeError("la3d peer synchronisation: a peer did not arrive within the timeout "
                               "(LA3D_PEER_TIMEOUT_MS); the gathered records are not valid")


class ShardedBoxFitter:
    """``BoxFitter`` (or ``RleBoxFitter`` / the all-pixels fit) for this rank's block of a ``B_total``-image
    batch plus the gather of the records.

    ``collective="p2p"`` (default on CUDA with more than one rank): the gathered ``[world*per, I, 64]``
    buffer of every rank lives in symmetric (peer-mapped) memory and the fit kernel of rank ``r``
    writes its records directly into slot ``r`` of EVERY rank's buffer over NVLink
    (``la3d_fit_boxes_to``).  The cross-GPU synchronisation rides in the step's own launches (flag rows in peer
    memory: the fit kernel waits before its first peer store, the next step's scan publishes the epoch), so a
    step adds no launch, no side stream and no NCCL call to the data path; a consumer publishes and waits with
    :meth:`wait_gathered` (``wait=True`` does it in the call).  The gathered buffer is double-buffered: the tensor a call
    returns stays valid until the next-but-one call (consume it on the same stream).
    ``collective="nccl"``: one ``all_gather_into_tensor`` after the fit (the plain form).
    ``source``: ``"masks"`` (byte masks, the default), ``"rle"`` (COCO run-length annotations:
    ``total_runs`` / ``max_runs`` size the plan; call with ``(depth, K, (run_counts, run_offsets), ground, ...)``)
    or ``"all"`` (every masked pixel instead of the random 500).
    """

    def __init__(self, B_total, I, H, W, device=None, out_dtype=torch.float64, group=None, collective=None,
                 source="masks", total_runs=0, max_runs=0):
        from . import ops
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B_total = int(B_total)
        self.start, self.stop, self.per = shard_range(self.B_total, self.rank, self.world)
        self.n_local = self.stop - self.start
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.out_dtype = out_dtype
        self.I, self.H, self.W = int(I), int(H), int(W)
        if source not in ("masks", "rle", "all"):
            raise ValueError(f"source must be 'masks', 'rle' or 'all', got {source!r}")
        self.source = source
        nb = max(self.n_local, 1)
        if source == "rle":
            self.local = ops.RleBoxFitter(nb, I, H, W, total_runs, max_runs, device=device, out_dtype=out_dtype)
        elif source == "all":
            self.local = None
            self._ws = torch.empty(int(ops._lib.load().la3d_fit_workspace_bytes(nb, I, H, W)), dtype=torch.uint8,
                                   device=self.device)
        else:
            self.local = ops.BoxFitter(nb, I, H, W, device=device, out_dtype=out_dtype)
        if collective is None:
            collective = "p2p" if self.world > 1 else "none"
        if collective not in ("p2p", "nccl", "none"):
            raise ValueError(f"collective must be 'p2p', 'nccl' or 'none', got {collective!r}")
        if self.world == 1:
            collective = "none"
        self.collective = collective
        if collective == "p2p":
            self.peers = PeerGather(self.per, I, self.world, self.rank, self.device, out_dtype, group)
        else:
            self.peers = None
            self.gathered = torch.empty((self.world * self.per, I, REC), dtype=out_dtype, device=device)
            self.slot = torch.full((self.per, I, REC), float("nan"), dtype=out_dtype, device=device)
            self.slot[..., O_STATUS] = -1

    def wait_gathered(self):
        """Make the current stream wait until the records of the last call have landed on every rank."""
        if self.peers is not None:
            self.peers.wait()

    def check_barrier_status(self):
        """Raises if a peer failed to arrive at a synchronisation point within the timeout."""
        if self.peers is not None:
            self.peers.check()

    def _fit_local(self, depth, K, masks, ground, method, yaw_steps, seed, out=None, sink=None, events=None):
        if self.source == "rle":
            run_counts, run_offsets = masks
            return self.local(depth, K, run_counts, run_offsets, ground, method, yaw_steps, seed, image_offset=self.start,
                              out=out, sink=sink)
        if self.source == "all":
            from . import ops
            if sink is None:
                sink = ops._lib.make_sink([out.data_ptr()], self.out_dtype == torch.float64)
            return ops.fit_boxes_all(depth, K, masks, ground, self.out_dtype, method, yaw_steps, sink=sink, workspace=self._ws)
        return self.local(depth, K, masks, ground, method, yaw_steps, seed, image_offset=self.start, out=out, sink=sink,
                          events=events)

    def __call__(self, depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, events=None, wait=True):
        """Inputs are this rank's block (``[n_local, ...]``).  Returns ``[B_total, I, 64]`` on every rank.
        ``wait=False`` (p2p only): do not make the current stream wait for the peers' records; the result is
        complete once :meth:`wait_gathered` has been enqueued before its consumer."""
        if self.peers is not None:
            self.peers.check()                               # a lost peer is fatal, never stale records
            sink, buf = self.peers.next_sink()
            if self.n_local:
                self._fit_local(depth, K, masks, ground, method, yaw_steps, seed, sink=sink)
            else:
                self.peers.signal()
            if wait:
                self.peers.wait()
            return buf[:self.B_total]
        if self.n_local:
            self._fit_local(depth, K, masks, ground, method, yaw_steps, seed, out=self.slot[:self.n_local], events=events)
        if self.world == 1:
            return self.slot[:self.B_total]
        dist.all_gather_into_tensor(self.gathered, self.slot, group=self.group)
        return self.gathered[:self.B_total]
