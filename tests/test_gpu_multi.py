"""Two ranks on two GPUs of one node: the peer-memory gather (records written into every rank's
buffer by the fit kernel, cross-GPU flags handled inside the same kernel) equals the NCCL all-gather and
the single-GPU result - for byte masks, run-length annotations and the all-pixels fit.
Skipped on boxes with fewer than two GPUs (the world_size-2 host logic is covered on CPU by
tests/test_dist_gloo.py); the one-GPU tests below cover the flag protocol itself."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from labelany3d_b200 import coco_rle, ops, synth
    from labelany3d_b200 import dist as la_dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B_total, I, H, W = 7, 3, 96, 128                       # 7 images over 2 ranks: the last rank pads one slot
    depth, K, masks, ground = synth.make_inputs(B_total, H, W, I, seed=5, device=dev, area=(0.05, 0.3))
    start, stop, per = la_dist.shard_range(B_total, rank, world)
    sl = slice(start, stop)
    res = {}
    for coll in ("p2p", "nccl"):
        fit = la_dist.ShardedBoxFitter(B_total, I, H, W, device=dev, collective=coll)
        for step in range(6):                              # several steps: epochs, double buffering, deferred waits
            rec = fit(depth[sl], K[sl], masks[sl], ground[sl], "sweep", 36, seed=11 + step, wait=(step % 2 == 1))
        fit.wait_gathered()
        torch.cuda.synchronize()
        fit.check_barrier_status()
        res[coll] = rec.cpu().numpy().copy()
    single = ops.fit_boxes(depth, K, masks, ground, "sweep", 36, seed=16).cpu().numpy()
    # run-length input, sharded: the same slot scheme
    planes = masks[sl].cpu().numpy().reshape(-1, H, W)
    counts, offsets, max_runs = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in planes])
    d_counts = torch.as_tensor(counts.view(np.int32), device=dev)
    d_off = torch.as_tensor(offsets, device=dev)
    fit = la_dist.ShardedBoxFitter(B_total, I, H, W, device=dev, collective="p2p", source="rle",
                                   total_runs=max(int(counts.size), 1), max_runs=max_runs)
    for step in range(3):
        rec = fit(depth[sl], K[sl], (d_counts, d_off), ground[sl], "sweep", 36, seed=14 + step)
    torch.cuda.synchronize()
    fit.check_barrier_status()
    res["rle"] = rec.cpu().numpy().copy()
    # every masked pixel, sharded
    fit = la_dist.ShardedBoxFitter(B_total, I, H, W, device=dev, collective="p2p", source="all")
    for step in range(3):
        rec = fit(depth[sl], K[sl], masks[sl], ground[sl], "pca")
    torch.cuda.synchronize()
    res["all"] = rec.cpu().numpy().copy()
    single_all = ops.fit_boxes_all(depth, K, masks, ground).cpu().numpy()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack([res["p2p"], res["nccl"], single, res["rle"], res["all"], single_all]))
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_gather_matches_nccl_and_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        p2p, nccl, single, rle, dense, single_dense = np.load(tmp_path / f"rank{rank}.npy")
        np.testing.assert_array_equal(p2p, single)
        np.testing.assert_array_equal(nccl, single)
        np.testing.assert_array_equal(rle, single)
        np.testing.assert_array_equal(dense, single_dense)


_ONE_GPU = r"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, {root!r})
from labelany3d_b200 import _lib, ops, synth
lib = _lib.load()
mode = sys.argv[1]
dev = torch.device("cuda", 0)
B, I, H, W = 3, 2, 64, 96
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=9, device=dev, area=(0.05, 0.3))
want = ops.fit_boxes(depth, K, masks, ground, "pca", seed=3, out_dtype=torch.float32)
# a "world" of two ranks played by one GPU: rank 0 is this process, rank 1's flag row is a second buffer
flags = [torch.zeros(8, dtype=torch.int32, device=dev) for _ in range(2)]
bufs = [torch.full((2 * B, I, 64), float("nan"), dtype=torch.float32, device=dev) for _ in range(2)]
counter = torch.zeros(1, dtype=torch.int32, device=dev)
status = torch.zeros(1, dtype=torch.int32).pin_memory()
fitter = ops.BoxFitter(B, I, H, W, device=dev, out_dtype=torch.float32)
def sink(epoch):
    return _lib.make_sink([b.data_ptr() for b in bufs], False, [f.data_ptr() for f in flags], counter.data_ptr(),
                          status.data_ptr(), epoch, 0)
arr = (ctypes.c_void_p * 2)(*[f.data_ptr() for f in flags])
if mode == "ok":
    fitter(depth, K, masks, ground, "pca", seed=3, sink=sink(1))            # epoch 1 needs nothing from the peers
    torch.cuda.synchronize()
    # the box kernel does not publish its own epoch: the next launch on the stream does
    assert flags[0][0].item() == 0 and flags[1][0].item() == 0, flags
    for b in bufs:
        assert torch.equal(b[:B].view(torch.int32), want.view(torch.int32))
    flags[0][1] = 1                                                           # "rank 1" publishes epoch 1
    fitter(depth, K, masks, ground, "pca", seed=3, sink=sink(2))            # its scan publishes OUR epoch 1 first
    torch.cuda.synchronize()
    assert flags[0][0].item() == 1 and flags[1][0].item() == 1, flags
    lib.la3d_peer_signal(arr, 1, 2, 2, None)                                  # rank 1 reaches epoch 2 ...
    lib.la3d_peer_barrier(arr, 0, 2, 2, status.data_ptr(), None)              # ... we publish ours and wait for it
    torch.cuda.synchronize()
    assert flags[0].tolist()[:2] == [2, 2] and flags[1][0].item() == 2 and int(status[0]) == 0
    # the step-wise entry publishes the previous epoch with a launch of its own
    bits, cc = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I, seed=3)
    counts, ranks = ops.sample_ranks(cc, B, I, H, W, prep=prep)
    flags[0][1] = 2
    s3 = sink(3)
    rc = lib.la3d_fit_scanned_to(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(), ranks.data_ptr(), B, I, H, W,
                                 0, 0, ctypes.byref(s3), None)
    torch.cuda.synchronize()
    assert rc == 0 and flags[0][0].item() == 2
    for b in bufs:
        assert torch.equal(b[:B].view(torch.int32), want.view(torch.int32))
    print("OK")
else:
    lib.la3d_set_peer_timeout_ms(200)
    fitter(depth, K, masks, ground, "pca", seed=3, sink=sink(1))
    torch.cuda.synchronize()
    try:
        fitter(depth, K, masks, ground, "pca", seed=3, sink=sink(2))        # rank 1 never published epoch 1
        torch.cuda.synchronize()
    except Exception as exc:
        print("RAISED", type(exc).__name__, "status", int(status[0]))
        sys.exit(0)
    print("NO ERROR status", int(status[0]))
"""


def _run(mode):
    return subprocess.run([sys.executable, "-c", _ONE_GPU.format(root=ROOT), mode], capture_output=True, text=True, timeout=300)


def test_sink_flag_protocol_on_one_gpu():
    proc = _run("ok")
    assert proc.returncode == 0 and "OK" in proc.stdout, proc.stdout + proc.stderr


def test_lost_peer_is_fatal_not_stale():
    """A peer that never publishes its epoch: the kernel sets the sticky status word and traps, the host sees
    a CUDA error (in a subprocess: the trap poisons the context)."""
    proc = _run("timeout")
    assert "RAISED" in proc.stdout and "status 1" in proc.stdout, proc.stdout + proc.stderr
