"""Per-step timeline of a short multi-GPU timed region (torchrun): where a 20-step region's constant goes.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/step_trace.py [IMAGES_PER_GPU]

Same loop as bench.py's device_leg (fence, rendezvous, K steps, wait_gathered), with one CUDA event after every
step; prints per rank the elapsed time of every step and of the closing wait.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from labelany3d_b200 import dist as la_dist
    from labelany3d_b200 import synth
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    K_STEPS, WARM = 20, 5
    cx = bench.Ctx(types.SimpleNamespace())
    torch.cuda.set_device(cx.dev)
    dist.init_process_group("nccl", device_id=cx.dev)
    dist.barrier()
    cx._rendezvous = torch.zeros(1, device=cx.dev)
    w = bench.SHAPE
    G = per * cx.world
    start, stop, _ = la_dist.shard_range(G, cx.rank, cx.world)
    depth, K, masks, ground = synth.make_inputs(per, w["H"], w["W"], w["I"], seed=bench.SEED + 1000 * cx.rank, device=cx.dev)
    fitter, _ = bench.make_fitter(cx, G, w["I"], w["H"], w["W"], "p2p")

    def step():
        return fitter(depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234, wait=False)

    for rep in range(3):
        for _ in range(WARM):
            step()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K_STEPS + 2)]
        cx.fence()
        dist.all_reduce(cx._rendezvous)
        ev[0].record()
        for k in range(K_STEPS):
            step()
            ev[k + 1].record()
        fitter.wait_gathered()
        ev[K_STEPS + 1].record()
        cx.fence()
        d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(K_STEPS + 1)]
        total = ev[0].elapsed_time(ev[-1]) * 1e3
        for r in range(cx.world):
            if r == cx.rank:
                print(f"rep {rep} rank {r}: total {total:.0f} us = {total / K_STEPS:.1f}/step; steps " +
                      " ".join(f"{x:.0f}" for x in d[:-1]) + f"; closing wait {d[-1]:.0f}", flush=True)
            dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
