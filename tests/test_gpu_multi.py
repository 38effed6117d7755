"""Two ranks on two GPUs of one node: the peer-memory gather (records written into every rank's
buffer by the fit kernel + flag barrier) equals the NCCL all-gather and the single-GPU result.
Skipped on boxes with fewer than two GPUs (the world_size-2 host logic is covered on CPU by
tests/test_dist_gloo.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from labelany3d_b200 import dist as la_dist
    from labelany3d_b200 import ops, synth
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
        for step in range(5):                              # several steps: epochs, double buffering, deferred barriers
            rec = fit(depth[sl], K[sl], masks[sl], ground[sl], "sweep", 36, seed=11 + step, wait=(step % 2 == 0))
        fit.wait_gathered()
        torch.cuda.synchronize()
        fit.check_barrier_status()
        res[coll] = rec.cpu().numpy().copy()
    single = ops.fit_boxes(depth, K, masks, ground, "sweep", 36, seed=15).cpu().numpy()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack([res["p2p"], res["nccl"], single]))
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_gather_matches_nccl_and_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        p2p, nccl, single = np.load(tmp_path / f"rank{rank}.npy")
        np.testing.assert_array_equal(p2p, single)
        np.testing.assert_array_equal(nccl, single)
