"""Benchmark of the box-fitting hot path (BASELINE.json: 3D boxes/sec on COCO-shape depth + masks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: BASELINE.json configs[1]
(256 images of 640x480, 8 instance masks each, 36-step yaw sweep) PER GPU (weak scaling; at N > 1
every rank fits its own block and the packed records are all-gathered each step).

Prints ONE JSON line (rank 0).  `value` = boxes/s with inputs resident in HBM, timed with CUDA
events over exactly K steps (max over ranks); `e2e` = the same metric through the public API with
inputs in pinned HOST memory, the H2D copies and the D2H read of the records inside the timed
region; `roofline` = the mask-scan kernel (the pass that reads every mask byte, the dominant HBM
stream of the step) against the measured copy peak; `cpu_baseline` = the NumPy port of the
reference path timed on this box's host cores.

`--impl reference` times the reference's CPU algorithm (the oracle port: the reference itself is
Python and cannot travel to the GPU box) on all host cores for the same metric and config.
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

WORKLOAD = dict(B=256, H=480, W=640, I=8, method="sweep", yaw_steps=36)   # BASELINE.json configs[1]
SEED = 1234 + 2
METRIC = "3D boxes/sec on COCO-shape 640x480 depth+masks"
UNIT = "boxes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="images of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rle", action="store_true", help="skip the run-length-input leg (N = 1 only)")
    ap.add_argument("--e2e-copy-depth", action="store_true",
                    help="e2e: copy the depth maps to the device each step instead of letting the fit kernel gather its "
                         "500 values per box from pinned host memory")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: records written into every rank's gathered buffer by the fit kernel over peer memory "
                         "(p2p) or one NCCL all-gather after the fit (nccl)")
    return ap.parse_args()


def config_dict(n_gpus, collective="p2p"):
    w = WORKLOAD
    how = {"p2p": ", records written by the fit kernel into every rank's gathered buffer over NVLink peer memory + flag barrier, each step",
           "nccl": ", NCCL all-gather of the packed records each step"}[collective]
    return {"workload": f"BASELINE configs[1]: batch={w['B']} images/GPU, {w['W']}x{w['H']} depth, {w['I']} instances/image, "
                        f"{w['yaw_steps']}-step yaw sweep",
            "images_per_gpu": w["B"], "global_images": w["B"] * n_gpus, "instances_per_image": w["I"],
            "height": w["H"], "width": w["W"], "method": w["method"], "yaw_steps": w["yaw_steps"], "subsample": 500,
            "parallelism": f"images sharded over {n_gpus} GPU(s)" + (how if n_gpus > 1 else ""),
            "l2": "inputs (944 MB/GPU/step) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference path on host cores
# ----------------------------------------------------------------------------------------------
_CPU_INPUTS = None


def _cpu_inputs(n_images):
    global _CPU_INPUTS
    if _CPU_INPUTS is None or _CPU_INPUTS[0].shape[0] != n_images:
        from labelany3d_b200 import synth
        w = WORKLOAD
        d, K, m, g = synth.make_inputs(n_images, w["H"], w["W"], w["I"], seed=SEED, device="cpu")
        _CPU_INPUTS = (d.numpy(), K.numpy(), m.numpy(), g.numpy())
    return _CPU_INPUTS


def _cpu_one_image(b):
    from oracle import la3d_oracle as orc
    d, K, m, g = _CPU_INPUTS
    w = WORKLOAD
    rec = orc.fit_boxes(d[b:b + 1], K[b:b + 1], m[b:b + 1], g[b:b + 1], w["method"], w["yaw_steps"], seed=1234,
                        image_offset=b, impl="library")
    return float(rec[0, 0, 0])


def cpu_baseline_single(n_images):
    """One process, one thread - the reference's real operating mode."""
    _cpu_inputs(n_images)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _cpu_one_image(0)                      # imports, scikit-learn warm-up
    t0 = time.perf_counter()
    for b in range(n_images):
        _cpu_one_image(b)
    dt = time.perf_counter() - t0
    return {"value": n_images * WORKLOAD["I"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_images} of the {WORKLOAD['B']} images of the workload, oracle port "
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
    n_images = min(WORKLOAD["B"], max(cores, 8))
    _cpu_inputs(n_images)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_one_image, range(n_images), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one_image, range(n_images), chunksize=1)
        dt = time.perf_counter() - t0
    value = args.steps * n_images * WORKLOAD["I"] / dt
    sample = (f"each step = {n_images} of the {WORKLOAD['B']} images of the workload through the oracle port of the "
              f"reference path (NumPy + scikit-learn/SciPy as the reference uses them), {cores} worker processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
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
def rle_leg(args, ops, depth, K, masks, ground, dev, w, B, I, H, W, fence):
    """boxes/s of the workload when the masks arrive as run-length annotations: device-resident runs
    (`value`) and pinned host runs copied in every step with the records copied back (`e2e`)."""
    import numpy as np
    import torch
    from labelany3d_b200 import coco_rle
    host_masks = masks.cpu().numpy().reshape(B * I, H, W)
    counts, offsets, max_runs = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in host_masks])
    del host_masks
    h_counts = torch.from_numpy(counts.view(np.int32)).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    d_counts, d_off = h_counts.to(dev), h_off.to(dev)
    fitter = ops.RleBoxFitter(B, I, H, W, d_counts.numel(), max_runs, device=dev, out_dtype=torch.float32)
    want = ops.BoxFitter(B, I, H, W, device=dev, out_dtype=torch.float32)(depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234)

    def step():
        return fitter(depth, K, d_counts, d_off, ground, w["method"], w["yaw_steps"], seed=1234)

    for _ in range(3):
        rec = step()
    same = bool(torch.equal(rec.view(torch.int32), want.view(torch.int32))) and not bool(fitter.rle_status.any())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    fence()
    ms = a.elapsed_time(b) / args.steps
    # end to end: runs, intrinsics and ground normals from pinned host memory every step, depth gathered in
    # place from pinned host memory by the fit kernel, records back to pinned host memory
    hd, hK, hg = (t.cpu().pin_memory() for t in (depth, K, ground))
    dK, dg = torch.empty_like(K), torch.empty_like(ground)
    host_rec = torch.empty((B, I, 64), dtype=torch.float32).pin_memory()

    def e2e_step():
        d_counts.copy_(h_counts, non_blocking=True)
        d_off.copy_(h_off, non_blocking=True)
        dK.copy_(hK, non_blocking=True)
        dg.copy_(hg, non_blocking=True)
        host_rec.copy_(fitter(hd, dK, d_counts, d_off, dg, w["method"], w["yaw_steps"], seed=1234), non_blocking=True)

    n = max(3, min(args.steps, 20))
    for _ in range(2):
        e2e_step()
    fence()
    a.record()
    for _ in range(n):
        e2e_step()
    b.record()
    fence()
    e2e_ms = a.elapsed_time(b) / n
    h2d = (h_counts.numel() * 4 + h_off.numel() * 8 + hK.numel() * 8 + hg.numel() * 8 + B * I * 500 * 32)
    return {"value": B * I / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "runs_per_step": int(h_counts.numel()),
            "records_identical_to_byte_mask_path": same,
            "e2e": {"value": B * I / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": host_rec.numel() * 4},
            "how": "masks of the same workload as COCO run-length annotations (column-major runs): la3d_fit_boxes_rle = "
                   "decode to bit planes (one CTA per plane, preparation CTAs in the same launch) -> subsample ranks -> fit"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from labelany3d_b200 import dist as la_dist
    from labelany3d_b200 import ops, synth
    import __graft_entry__
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    else:
        __graft_entry__.build()

    w = WORKLOAD
    B, I, H, W = w["B"], w["I"], w["H"], w["W"]
    # every rank owns B images of a B*world batch; inputs are generated where they live
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=SEED + 1000 * rank, device=dev)
    fitter = None
    if world > 1:
        try:
            fitter = la_dist.ShardedBoxFitter(B * world, I, H, W, device=dev, out_dtype=torch.float32, collective=args.collective)
            ok = torch.ones(1, device=dev)
        except Exception as exc:  # noqa: BLE001  (no peer mapping on this box: say so and use the NCCL gather)
            print(f"[bench] rank {rank}: peer-memory gather unavailable ({exc!r}); falling back to --collective nccl", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            args.collective = "nccl"
            fitter = la_dist.ShardedBoxFitter(B * world, I, H, W, device=dev, out_dtype=torch.float32, collective="nccl")
    single = ops.BoxFitter(B, I, H, W, device=dev, out_dtype=torch.float32)
    boxes_per_step = B * I * world

    def step(events=None):
        if fitter is not None and events is None:
            # wait=False: the peer barrier of step k runs on a side stream and gates only the fit of step k+1;
            # the timed region ends with wait_gathered() + a device synchronisation, so every gather is inside it
            return fitter(depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234, wait=False)
        # per-kernel timing (events) is a local matter: this rank's block without the gather
        return single(depth, K, masks, ground, w["method"], w["yaw_steps"], seed=1234, image_offset=B * rank, events=events)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region A: exactly K steps through the public API (device-resident inputs)
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    if clocks:
        clocks.start()
    t_beg.record()
    for k in range(args.steps):
        step()
    if fitter is not None:
        fitter.wait_gathered()
    t_end.record()
    fence()
    if clocks:
        clocks.stop()
    elapsed = torch.tensor([t_beg.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    ms_total = float(elapsed.item())

    # ---- timed region B: the same K steps with the four kernels of a step issued one after the
    # other on the current stream and CUDA events between them: per-kernel durations for the
    # roofline (in region A the mask-independent preparation rides in the scan's launch)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for _ in range(3):
        step(evs[0])
    fence()
    if clocks:
        clocks.start()
    for k in range(args.steps):
        step(evs[k])
    fence()
    if clocks:
        clocks.stop()
    k_prep = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    k_scan = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    k_samp = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)
    k_fit = statistics.mean(e[3].elapsed_time(e[4]) for e in evs)

    # ---- end to end through the public API: pinned host buffers -> boxes back on the host
    e2e = None
    if not args.no_e2e:
        hd, hK, hm, hg = (t.cpu().pin_memory() for t in (depth, K, masks, ground))
        host_rec = torch.empty((B * world, I, 64), dtype=torch.float32).pin_memory()
        host_fit = ops.HostBoxFitter(fitter or single, B, I, H, W, device=dev, depth_in_place=not args.e2e_copy_depth)
        h2d = host_fit.h2d_bytes(B, I, H, W)
        d2h = host_rec.numel() * host_rec.element_size()

        def e2e_step():
            # public host-facing call: pinned host buffers in, records back in pinned host memory
            host_fit(hd, hK, hm, hg, host_rec, w["method"], w["yaw_steps"], seed=1234)

        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(2):
            e2e_step()
        fence()
        if clocks:
            clocks.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b.record()
        fence()
        if clocks:
            clocks.stop()
        e2e_ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": boxes_per_step * e2e_steps / (float(e2e_ms.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "ms_per_step": float(e2e_ms.item()) / e2e_steps,
               "how": "ops.HostBoxFitter: masks / K / ground copied on a copy stream (two buffer sets, the copy of step "
                      "k+1 overlaps the kernels of step k), records copied back each step; depth "
                      + ("copied to the device each step" if args.e2e_copy_depth else
                         "left in pinned host memory, 500 values per box gathered over PCIe by the fit kernel "
                         "(counted as 32-byte sectors in h2d_bytes_per_step)")}

    # ---- the same workload with the masks given as COCO run-length annotations (the format the reference's
    # loader reads, src/util.py:361-370): decoded on the device into bit planes, no byte masks anywhere.
    # An additional figure, not the headline: `value` / `e2e` above keep the byte-mask input of BASELINE.json.
    rle = None
    if world == 1 and not args.no_rle:
        try:
            rle = rle_leg(args, ops, depth, K, masks, ground, dev, w, B, I, H, W, fence)
        except Exception as exc:  # noqa: BLE001  (an extra leg must not take the contract line down)
            rle = {"error": repr(exc)}

    # ---- the deterministic form: every masked pixel instead of the reference's random 500 (la3d_fit_boxes_all;
    # pca heading).  Reads the mask bytes once plus, twice, the depth under the masks.  An additional figure.
    dense = None
    if world == 1 and not args.no_rle:
        try:
            for _ in range(3):
                ops.fit_boxes_all(depth, K, masks, ground, out_dtype=torch.float32)
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fence()
            a_ev.record()
            for _ in range(args.steps):
                ops.fit_boxes_all(depth, K, masks, ground, out_dtype=torch.float32)
            b_ev.record()
            fence()
            dense_ms = a_ev.elapsed_time(b_ev) / args.steps
            dense = {"value": boxes_per_step / (dense_ms * 1e-3), "unit": UNIT, "ms_per_step": dense_ms,
                     "points_per_step": int(masks.sum().item()), "method": "pca",
                     "how": "la3d_fit_boxes_all: mask scan (preparation CTAs in its grid) -> one CTA per box reduces the "
                            "moments and extents over ALL its masked pixels (no subsample, no random draw)"}
        except Exception as exc:  # noqa: BLE001
            dense = {"error": repr(exc)}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        scan_bytes = B * I * H * W                       # algorithmic: every mask byte once
        launches_per_step = 3                            # mask_scan_kernel (+ the prep CTAs in its grid), sample_kernel, fit_kernel
        achieved = scan_bytes / (k_scan * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": boxes_per_step * args.steps / (ms_total * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, args.collective),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "kernels_ms": {"fit_prepare": k_prep, "mask_scan": k_scan, "sample_ranks": k_samp, "fit_boxes": k_fit,
                           "how": "second timed pass of K steps, the four kernels serialised on one stream with CUDA events "
                                  "between them (in the headline pass fit_prepare is B extra CTAs inside the scan's launch)"},
            "roofline": {"kernel": "mask_scan_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": scan_bytes},
            "clocks": clocks.summary() if clocks else None,
            "rle_input": rle,
            "all_pixels": dense,
        }
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(traffic_path):
            t = json.load(open(traffic_path)).get("mask_scan_kernel<1, 1, 1>")     # the scan with the prep CTAs in its grid
            line["roofline"]["traffic"] = t[0] if isinstance(t, list) else t
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
