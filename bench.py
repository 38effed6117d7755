"""Benchmark of the box-fitting hot path (BASELINE.json: 3D boxes/sec on COCO-shape depth + masks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling strong|weak] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input.  Default (`--scaling strong`,
BASELINE.json's target "strong scaling 1 -> 8 GPUs"): a FIXED global batch of 2048 images of configs[1]'s
shape (640x480 depth, 8 instance masks per image, 36-step yaw sweep) - configs[2]'s sharded batch; rank r
of N fits images shard_range(2048, r, N), so one GPU fits all 2048 (8 x configs[1]'s batch of 256) and each of
8 GPUs fits 256 (= configs[1] exactly); the packed records are gathered on every rank each step.
`--scaling weak`: 256 images per GPU at every N (configs[1] per GPU).

Prints ONE JSON line (rank 0).  `value` = boxes/s with inputs resident in HBM, timed with CUDA events over
exactly K steps (max over ranks); `e2e` = the same metric through the public API with inputs in pinned HOST
memory, the H2D copies and the D2H read of the records inside the timed region; `roofline` = the mask-scan
kernel (the pass that reads every mask byte, the dominant HBM stream of the step) against the measured copy
peak; `roofline_lift` = the depth-lift kernel (la3d_depth_lift, 16 B / pixel) at configs[1] and configs[3];
`cpu_baseline` = the NumPy port of the reference path timed on this box's host cores.  Further legs: the
reference-default `pca` method, configs[2] / configs[4] as strong-scaling legs at every N, configs[1] / configs[3]
at their one-GPU size, the run-length-annotation input (device-resident and end to end, sharded like the
headline), the all-pixels fit.  At N > 1 the gathered records are verified bitwise against an NCCL all-gather of
every rank's local records (`gather_verified`).

`--impl reference` times the reference's CPU algorithm (the oracle port: the reference itself is Python and
cannot travel to the GPU box) on all host cores for the same metric and config.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = dict(H=480, W=640, I=8, method="sweep", yaw_steps=36)      # BASELINE.json configs[1] per image
STRONG_BATCH = 2048                                                 # configs[2]'s global batch, fixed at every N
WEAK_BATCH = 256                                                    # configs[1]'s batch, per GPU
SEED = 1234 + 2
METRIC = "3D boxes/sec on COCO-shape 640x480 depth+masks"
UNIT = "boxes/s"
HOST_CEILING_NOTE = ("byte masks cross PCIe (~50 GB/s per GPU); all GPUs of the box hang off one NUMA node, whose host "
                     "side saturated at ~177 GB/s aggregate in round 1 (SCALE_r01)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="images of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="headline only: skip the extra legs (rle, configs, lift, all-pixels)")
    ap.add_argument("--e2e-copy-depth", action="store_true",
                    help="e2e: copy the depth maps to the device each step instead of letting the fit gather its "
                         "500 values per box from pinned host memory")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: records written into every rank's gathered buffer by the fit kernel over peer memory "
                         "(p2p) or one NCCL all-gather after the fit (nccl)")
    return ap.parse_args()


def global_batch(args):
    return STRONG_BATCH if args.scaling == "strong" else WEAK_BATCH * max(args.gpus, 1)


def config_dict(args, collective="p2p"):
    w, n = SHAPE, max(args.gpus, 1)
    G = global_batch(args)
    how = {"p2p": ", records written by the fit kernel into every rank's gathered buffer over NVLink peer memory, the "
                  "cross-GPU flags handled inside the same kernel, each step",
           "nccl": ", NCCL all-gather of the packed records each step"}.get(collective, "")
    what = (f"fixed global batch of {G} images (BASELINE configs[2]'s sharded batch; = {G // WEAK_BATCH} x configs[1]'s batch), "
            if args.scaling == "strong" else f"{WEAK_BATCH} images per GPU (BASELINE configs[1] per GPU), ")
    return {"workload": what + f"BASELINE configs[1] shape: {w['W']}x{w['H']} depth, {w['I']} instances/image, "
                               f"{w['yaw_steps']}-step yaw sweep",
            "global_images": G, "images_per_gpu": (G + n - 1) // n, "instances_per_image": w["I"],
            "height": w["H"], "width": w["W"], "method": w["method"], "yaw_steps": w["yaw_steps"], "subsample": 500,
            "parallelism": f"images sharded over {n} GPU(s)" + (how if n > 1 else ""),
            "l2": f"inputs ({(G + n - 1) // n * w['H'] * w['W'] * (w['I'] + 4) / 1e6:.0f} MB/GPU/step) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference path on host cores
# ----------------------------------------------------------------------------------------------
_CPU_INPUTS = None


def _cpu_inputs(n_images):
    global _CPU_INPUTS
    if _CPU_INPUTS is None or _CPU_INPUTS[0].shape[0] != n_images:
        from labelany3d_b200 import synth
        w = SHAPE
        d, K, m, g = synth.make_inputs(n_images, w["H"], w["W"], w["I"], seed=SEED, device="cpu")
        _CPU_INPUTS = (d.numpy(), K.numpy(), m.numpy(), g.numpy())
    return _CPU_INPUTS


def _cpu_one_image(b):
    from oracle import la3d_oracle as orc
    d, K, m, g = _CPU_INPUTS
    w = SHAPE
    rec = orc.fit_boxes(d[b:b + 1], K[b:b + 1], m[b:b + 1], g[b:b + 1], w["method"], w["yaw_steps"], seed=1234,
                        image_offset=b, impl="library")
    return float(rec[0, 0, 0])


def cpu_baseline_single(n_images, G):
    """One process, one thread - the reference's real operating mode."""
    _cpu_inputs(n_images)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _cpu_one_image(0)                      # imports, scikit-learn warm-up
    t0 = time.perf_counter()
    for b in range(n_images):
        _cpu_one_image(b)
    dt = time.perf_counter() - t0
    return {"value": n_images * SHAPE["I"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_images} of the {G} images of the workload, oracle port "
                      f"(NumPy + scikit-learn PCA like the reference), one process, {dt:.1f} s"}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm on all host cores, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"               # parallelism comes from one worker per core
    cores = os.cpu_count() or 1
    G = global_batch(args)
    n_images = min(G, max(cores, 8))
    _cpu_inputs(n_images)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_one_image, range(n_images), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one_image, range(n_images), chunksize=1)
        dt = time.perf_counter() - t0
    value = args.steps * n_images * SHAPE["I"] / dt
    sample = (f"each step = a bounded sample of {n_images} of the {G} images of the workload ({n_images * SHAPE['I']} boxes per "
              f"step; the GPU arm's step is all {G} images) through the oracle port of the reference path (NumPy + "
              f"scikit-learn/SciPy as the reference uses them), {cores} worker processes; boxes/s is per box, so the two "
              f"arms compare directly")
    cfg = config_dict(args)
    cfg["reference_images_per_step"] = n_images
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks during the timed region (NVML poll thread; nvidia-smi as a fallback)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag, self.thread = index, [], set(), False, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join()
            self.thread = None

    def summary(self):
        if self.nv is None or not self.samples:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                sm, mx = (float(x) for x in out.strip().split(","))
                return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": [], "samples": 1, "how": "nvidia-smi after the run"}
            except Exception:  # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": "unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "how": "NVML polled every 2 ms during the timed regions"}


# ----------------------------------------------------------------------------------------------
class Ctx:
    """What every leg needs: rank / world, device, the fence and the max-over-ranks reduction."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local_rank)
        self.clocks = None
        self._rendezvous = None

    def fence(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, fn, steps, warmup, after=None, clocks=False):
        """W untimed calls, then exactly `steps` calls between two CUDA events, fenced on both sides (barrier +
        device synchronisation); returns the max over ranks of the elapsed milliseconds."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if clocks and self.clocks:
            self.clocks.start()          # before the fence: the poll thread is up when the ranks leave it
        self.fence()
        if self.world > 1:
            # the ranks' host threads leave the fence tens of microseconds apart; a collective enqueued right before
            # the start event makes every rank's STREAM wait for the slowest host, so the K steps are timed from a
            # common device-side start (with the driver's K = 20 the host skew was ~10 us per step otherwise)
            self.dist.all_reduce(self._rendezvous)
        a.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        b.record()
        self.fence()
        if clocks and self.clocks:
            self.clocks.stop()
        return self.max_ms(a.elapsed_time(b))


def make_fitter(cx, G, I, H, W, collective, source="masks", total_runs=0, max_runs=0):
    """The public API object of a leg: BoxFitter on one GPU, ShardedBoxFitter (peer-memory gather) otherwise."""
    import torch
    from labelany3d_b200 import dist as la_dist
    from labelany3d_b200 import ops
    if cx.world == 1:
        if source == "rle":
            return ops.RleBoxFitter(G, I, H, W, total_runs, max_runs, device=cx.dev, out_dtype=torch.float32), "none"
        return ops.BoxFitter(G, I, H, W, device=cx.dev, out_dtype=torch.float32), "none"
    try:
        fitter = la_dist.ShardedBoxFitter(G, I, H, W, device=cx.dev, out_dtype=torch.float32, collective=collective,
                                          source=source, total_runs=total_runs, max_runs=max_runs)
        ok = True
    except Exception as exc:  # noqa: BLE001  (no peer mapping on this box: say so and use the NCCL gather)
        print(f"[bench] rank {cx.rank}: peer-memory gather unavailable ({exc!r}); falling back to nccl", file=sys.stderr)
        fitter, ok = None, False
    if not cx.all_true(ok):
        collective = "nccl"
        fitter = la_dist.ShardedBoxFitter(G, I, H, W, device=cx.dev, out_dtype=torch.float32, collective="nccl",
                                          source=source, total_runs=total_runs, max_runs=max_runs)
    return fitter, collective


def device_leg(cx, G, I, H, W, method, yaw_steps, steps, warmup, seed, collective="p2p", verify=False, keep=False):
    """boxes/s of one configuration with device-resident inputs: rank r fits shard_range(G, r, world)."""
    import torch
    from labelany3d_b200 import dist as la_dist
    from labelany3d_b200 import ops, synth
    start, stop, per = la_dist.shard_range(G, cx.rank, cx.world)
    n = stop - start
    depth, K, masks, ground = synth.make_inputs(max(n, 1), H, W, I, seed=seed + 1000 * cx.rank, device=cx.dev)
    fitter, collective = make_fitter(cx, G, I, H, W, collective)

    def step():
        if cx.world > 1:
            return fitter(depth[:n], K[:n], masks[:n], ground[:n], method, yaw_steps, seed=1234, wait=False)
        return fitter(depth, K, masks, ground, method, yaw_steps, seed=1234)

    ms = cx.timed(step, steps, warmup, after=(fitter.wait_gathered if cx.world > 1 else None), clocks=keep)
    out = {"value": G * I * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "global_images": G, "images_per_gpu": per, "instances_per_image": I, "height": H, "width": W,
           "method": method, "yaw_steps": yaw_steps, "collective": collective}
    if cx.world > 1:
        fitter.check_barrier_status()
        if verify:
            # the gathered buffer of the LAST step against an NCCL all-gather of every rank's own records, bitwise
            gathered = fitter(depth[:n], K[:n], masks[:n], ground[:n], method, yaw_steps, seed=1234, wait=True).clone()
            local = ops.BoxFitter(max(n, 1), I, H, W, device=cx.dev, out_dtype=torch.float32)(
                depth, K, masks, ground, method, yaw_steps, seed=1234, image_offset=start)
            want = la_dist.all_gather_records(local[:n].contiguous(), total=G)
            same = bool(torch.equal(gathered.view(torch.int32), want.view(torch.int32)))
            ok_status = bool((gathered[..., 41] == 0).all().item())
            out["gather_verified"] = cx.all_true(same)
            out["all_boxes_ok"] = cx.all_true(ok_status)
            fitter.check_barrier_status()
    if keep:
        return out, (depth, K, masks, ground, fitter, collective, start, n)
    del depth, K, masks, ground, fitter
    torch.cuda.empty_cache()
    return out


def rle_leg(cx, depth, K, masks, ground, G, start, n, collective, steps, warmup):
    """boxes/s of the workload when the masks arrive as run-length annotations: device-resident runs
    (`value`) and pinned host runs copied in every step with the records copied back (`e2e`), sharded like
    the headline (the fit kernel writes into every rank's gathered buffer)."""
    import torch
    from labelany3d_b200 import coco_rle, ops
    w = SHAPE
    I, H, W = w["I"], w["H"], w["W"]
    parts, offs, max_runs, base = [], [torch.zeros(1, dtype=torch.int64, device=cx.dev)], 0, 0
    for b0 in range(0, n, 64):                         # encode on the device, 64 images at a time
        c, o, mx = coco_rle.runs_from_masks_device(masks[b0:min(b0 + 64, n)].reshape(-1, H, W))
        parts.append(c)
        offs.append(o[1:] + base)
        base += int(o[-1].item())
        max_runs = max(max_runs, mx)
    d_counts, d_off = torch.cat(parts), torch.cat(offs)
    del parts, offs
    total_runs = cx.max_ms(float(d_counts.numel()))    # every rank sizes its plan for the largest shard
    fitter, collective = make_fitter(cx, G, I, H, W, collective, source="rle", total_runs=int(total_runs),
                                     max_runs=int(cx.max_ms(float(max_runs))))
    want = ops.BoxFitter(n, I, H, W, device=cx.dev, out_dtype=torch.float32)(
        depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234, image_offset=start)

    def step(dep=depth, kk=K, cc=d_counts, oo=d_off, gg=ground, wait=False):
        if cx.world > 1:
            return fitter(dep, kk, (cc, oo), gg, w["method"], w["yaw_steps"], seed=1234, wait=wait)
        return fitter(dep, kk, cc, oo, gg, w["method"], w["yaw_steps"], seed=1234)

    rec = step(wait=True)
    local = rec[start:start + n]
    rle_status = fitter.local.rle_status if cx.world > 1 else fitter.rle_status
    same = cx.all_true(bool(torch.equal(local.view(torch.int32), want.view(torch.int32))) and not bool(rle_status.any()))
    del want
    ms = cx.timed(step, steps, warmup, after=(fitter.wait_gathered if cx.world > 1 else None))
    # end to end: runs, intrinsics and ground normals from pinned host memory every step, the depth values the fit
    # needs gathered in place from pinned host memory, records back to pinned host memory
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)  # noqa: E731
    h_counts, h_off, hd, hK, hg = pin(d_counts), pin(d_off), pin(depth), pin(K), pin(ground)
    dK, dg = torch.empty_like(K), torch.empty_like(ground)
    host_rec = torch.empty((G, I, 64), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        d_counts.copy_(h_counts, non_blocking=True)
        d_off.copy_(h_off, non_blocking=True)
        dK.copy_(hK, non_blocking=True)
        dg.copy_(hg, non_blocking=True)
        host_rec.copy_(step(hd, dK, d_counts, d_off, dg, wait=True), non_blocking=True)

    e2e_steps = max(3, min(steps, 10))
    e2e_ms = cx.timed(e2e_step, e2e_steps, 2)
    if cx.world > 1:
        fitter.check_barrier_status()
    # the in-place depth reads cross the bus as 128-byte line requests (tools/gather_ceiling.py): at most one per sample
    h2d = h_counts.numel() * 4 + h_off.numel() * 8 + hK.numel() * 8 + hg.numel() * 8 + n * I * 500 * 128
    return {"value": G * I * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "runs_per_step_per_gpu": int(h_counts.numel()),
            "records_identical_to_byte_mask_path": same, "collective": collective,
            "e2e": {"value": G * I * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_rec.numel() * 4, "steps": e2e_steps,
                    "depth": "read in place from pinned host memory by the fit kernel, samples sorted by address so that "
                             "lanes of a warp share 128-byte PCIe line requests (ceiling on this box: 343 M requests/s, "
                             "profiles/r2_ae_gather_ceiling.json); h2d_bytes counts one line per sample (upper bound)"},
            "how": "masks of the same workload as COCO run-length annotations (column-major runs): la3d_fit_boxes_rle = "
                   "decode to bit planes (one CTA per plane, preparation CTAs in the same launch) -> subsample ranks -> fit"}


def lift_leg(cx, peak, steps):
    """The depth-lift kernel alone (la3d_depth_lift, float32 points: 4 B read + 12 B written per pixel) at
    configs[1] and configs[3] (1536 x 1536, the HBM-bound lift config): CUDA events around K launches."""
    import torch
    from labelany3d_b200 import ops, synth
    out = {}
    traffic = _traffic()
    for name, (B, H, W) in {"cfg2_256x640x480": (256, 480, 640), "cfg4_128x1536x1536": (128, 1536, 1536)}.items():
        depth, K, _, _ = synth.make_inputs(2, H, W, 1, seed=SEED, device=cx.dev)
        depth = depth[:1].expand(B, H, W).contiguous()
        depth += torch.rand((B, 1, 1), device=cx.dev)
        K = K[:1].expand(B, 3, 3).contiguous()
        pts = ops.depth_lift(depth, K, out_dtype=torch.float32)                     # warm-up + allocation
        lib, st = ops._lib.load(), torch.cuda.current_stream().cuda_stream

        def launch():
            ops._lib.check(lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, pts.data_ptr(), 0, st),
                           "la3d_depth_lift")

        ms = cx.timed(launch, steps, 3) / steps
        nbytes = B * H * W * 16
        ach = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"kernel": "lift_bulk_kernel", "bound": "hbm", "ms_per_launch": ms, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "algorithmic_bytes_per_launch": nbytes, "traffic": traffic.get("lift_bulk_kernel@" + name),
                     "pixels_per_s": B * H * W / (ms * 1e-3), "l2": f"{nbytes / 1e6:.0f} MB per launch exceed the 126 MB L2"}
        del depth, K, pts
        torch.cuda.empty_cache()
    return out


def _traffic():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from labelany3d_b200 import ops
    import __graft_entry__
    cx = Ctx(args)
    world, rank = cx.world, cx.rank
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(cx.local_rank)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=cx.dev)
        dist.barrier()
        cx._rendezvous = torch.zeros(1, device=cx.dev)
    else:
        __graft_entry__.build()

    w = SHAPE
    I, H, W = w["I"], w["H"], w["W"]
    G = global_batch(args)
    cx.clocks = ClockSampler(cx.local_rank) if rank == 0 else None
    warm = max(args.warmup, 3)

    # ---- headline: K steps through the public API, device-resident inputs, gather verified afterwards
    head, (depth, K, masks, ground, fitter, collective, start, n) = device_leg(
        cx, G, I, H, W, w["method"], w["yaw_steps"], args.steps, warm, SEED, args.collective, verify=True, keep=True)
    boxes_per_step = G * I

    # ---- durations of the step's own three launches on this rank's block: the library records four CUDA events
    # on the launching stream around them (la3d_debug_step_events): the scan WITH its preparation CTAs and its
    # sparse bit-plane stores, the sampler, the fit - the same kernels the headline pass just ran
    single = fitter if world == 1 else fitter.local
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step_events(ev):
        single(depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234, image_offset=start, events=ev)

    for _ in range(3):
        step_events(evs[0])
    cx.fence()
    if cx.clocks:
        cx.clocks.start()
    for k in range(args.steps):
        step_events(evs[k])
    cx.fence()
    if cx.clocks:
        cx.clocks.stop()
    k_scan = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    k_samp = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    k_fit = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)

    # ---- end to end through the public API: pinned host buffers -> boxes back on the host
    e2e = None
    if not args.no_e2e:
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)  # noqa: E731
        hd, hK, hm, hg = pin(depth[:n]), pin(K[:n]), pin(masks[:n]), pin(ground[:n])
        host_rec = torch.empty((G, I, 64), dtype=torch.float32, pin_memory=True)
        host_fit = ops.HostBoxFitter(fitter, n, I, H, W, device=cx.dev, depth_in_place=not args.e2e_copy_depth)
        h2d = host_fit.h2d_bytes(n, I, H, W)
        d2h = host_rec.numel() * host_rec.element_size()
        e2e_steps = max(3, min(args.steps, 10))
        e2e_ms = cx.timed(lambda: host_fit(hd, hK, hm, hg, host_rec, w["method"], w["yaw_steps"], seed=1234),
                          e2e_steps, 2, clocks=True)
        e2e = {"value": boxes_per_step * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "ms_per_step": e2e_ms / e2e_steps, "host_ceiling": HOST_CEILING_NOTE,
               "how": "ops.HostBoxFitter: masks / K / ground copied on a copy stream (two buffer sets, the copy of step "
                      "k+1 overlaps the kernels of step k), records copied back each step; depth "
                      + ("copied to the device each step" if args.e2e_copy_depth else
                         "left in pinned host memory, 500 values per box read in place over PCIe by the fit kernel, sorted "
                         "by address (counted as one 128-byte line request per sample in h2d_bytes_per_step: an upper bound)")}
        del hd, hK, hm, hg, host_rec, host_fit

    legs = {}
    if not args.no_legs:
        def guarded(name, fn):
            try:
                legs[name] = fn()
            except Exception as exc:  # noqa: BLE001  (an extra leg must not take the contract line down)
                legs[name] = {"error": repr(exc)}
            cx.fence()

        # the same workload with the masks given as COCO run-length annotations (the format the reference's loader
        # reads, src/util.py:361-370): decoded on the device into bit planes, no byte masks anywhere; sharded
        guarded("rle_input", lambda: rle_leg(cx, depth[:n], K[:n], masks[:n], ground[:n], G, start, n, collective,
                                             args.steps, warm))
        if world == 1:
            # the deterministic form: every masked pixel instead of the reference's random 500 (pca heading)
            def dense():
                ms = cx.timed(lambda: ops.fit_boxes_all(depth, K, masks, ground, out_dtype=torch.float32), max(args.steps // 3, 3), 2)
                steps = max(args.steps // 3, 3)
                return {"value": boxes_per_step * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                        "points_per_step": int(masks.sum().item()), "method": "pca",
                        "how": "la3d_fit_boxes_all: mask scan (preparation CTAs in its grid) -> one CTA per box reduces the "
                               "moments and extents over ALL its masked pixels (no subsample, no random draw)"}
            guarded("all_pixels", dense)
    del depth, K, masks, ground, fitter, single
    torch.cuda.empty_cache()

    if not args.no_legs:
        short = max(args.steps // 3, 5)
        # the reference's default method on the headline workload, then BASELINE's other sharded configs as
        # strong-scaling legs (fixed global batch at every N)
        guarded("pca", lambda: device_leg(cx, G, I, H, W, "pca", 0, args.steps, warm, SEED, args.collective))
        guarded("cfg3_2048x10_pca", lambda: device_leg(cx, 2048, 10, 480, 640, "pca", 0, short, 3, 1234 + 3, args.collective))
        guarded("cfg5_1024x32_sweep360", lambda: device_leg(cx, 1024, 32, 480, 640, "sweep", 360, short, 3, 1234 + 5, args.collective))
        if world == 1:
            # BASELINE's one-GPU configs at exactly their size
            guarded("cfg2_256x8_sweep36", lambda: device_leg(cx, 256, 8, 480, 640, "sweep", 36, args.steps, warm, SEED))
            guarded("cfg2_256x8_pca", lambda: device_leg(cx, 256, 8, 480, 640, "pca", 0, args.steps, warm, SEED))
            guarded("cfg4_128x20_1536_pca", lambda: device_leg(cx, 128, 20, 1536, 1536, "pca", 0, short, 3, 1234 + 4))

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    else:
        peak, peak_src = 6650.0, ""
    lift = None
    if world == 1 and not args.no_legs:
        try:
            lift = lift_leg(cx, peak, max(args.steps, 10))
        except Exception as exc:  # noqa: BLE001
            lift = {"error": repr(exc)}

    if rank == 0:
        scan_bytes = n * I * H * W                        # algorithmic: every mask byte of this rank's block once
        launches_per_step = 3                             # mask_scan_kernel (+ the prep CTAs in its grid), sample_kernel, fit_kernel
        achieved = scan_bytes / (k_scan * 1e-3) / 1e9
        traffic = _traffic()
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, collective),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "gather_verified": head.get("gather_verified"), "all_boxes_ok": head.get("all_boxes_ok"),
            "kernels_ms": {"mask_scan": k_scan, "sample_ranks": k_samp, "fit_boxes": k_fit,
                           "images": n,
                           "how": "second timed pass of K steps over rank 0's block through the same public call; the library "
                                  "records CUDA events on the launching stream around the step's own three launches "
                                  "(la3d_debug_step_events): the scan with the preparation CTAs in its grid, the sampler, the fit"},
            "roofline": {"kernel": "mask_scan_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic.get(f"mask_scan_kernel@{n}x{I}x{H}x{W}"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": scan_bytes,
                         "step_frac": scan_bytes / (head["ms_per_step"] * 1e-3) / 1e9 / peak},
            "roofline_lift": lift,
            "clocks": cx.clocks.summary() if cx.clocks else None,
            "legs": legs,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample, G)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
