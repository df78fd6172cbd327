#!/usr/bin/env python
"""bench.py -- images/sec of the post-backbone path (decode + NMS + crop + PRN + keypoint decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2] [--prn-mode bf16]

A "step" is one pass of the hot path over one batch of synthetic post-backbone tensors of BASELINE.json's
configs[1] (640x640, batch 8, 9 anchors / location, score thr 0.3, NMS IoU 0.5, <= 25 boxes per image) PER GPU.
With N > 1 (torchrun, one process per GPU) every rank runs its own batch: images shard across ranks with no
collective on the data path, so scaling is "weak" and `value` is the aggregate over ranks
divided by the slowest rank's device time.

Legs (all in one run, one JSON line on rank 0):
  value      inputs already resident in HBM, outputs left in HBM; K steps fed round-robin to --lanes Detector handles
             (default 3, each on its own stream: neighbouring batches overlap), CUDA events around the K steps; the same
             K steps back to back on ONE handle and stream are reported beside it (`single_lane`).
             Inputs rotate over a ring of distinct batches whose total size exceeds the 126 MB L2 (no flush needed).
  e2e        the same steps through the reference-facing call with HOST buffers: pinned host -> device copies of
             the step's inputs, the path, and device -> host copies of all seven outputs (heatmaps included, as
             Detector.__call__ fetches them) inside the timed region; wall clock bracketed by synchronisation.
  roofline   per-kernel CUDA-event times of K more steps (mpn_set_profiling); the dominant kernel's algorithmic
             bytes / its average duration against MEASURED_PEAKS.json.
  cpu_baseline   the CPU oracle port (oracle/, BLAS for the two dense layers) on the host cores, bounded sample.
--impl reference runs only the CPU leg as the reference arm (TensorFlow 1.15 cannot be installed: DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU legs (rank 0 only) are meant to use all host cores, and the
# OpenMP runtimes read the variable when they are first loaded -- so fix it before numpy / torch / the oracle are imported.
if int(os.environ.get("RANK", "0")) == 0:
    try:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    except AttributeError:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
# ONE JSON line on stdout: NCCL prints its version banner to stdout whenever NCCL_DEBUG is set (VERSION, WARN and INFO all
# do), and any native library may print.  So NCCL_DEBUG is dropped unless asked for, and main() points file descriptor 1
# at stderr for the life of the process; emit() writes the result line to the real stdout.
if "MPN_NCCL_DEBUG" in os.environ:
    os.environ["NCCL_DEBUG"] = os.environ["MPN_NCCL_DEBUG"]
else:
    os.environ.pop("NCCL_DEBUG", None)
_REAL_STDOUT = None


def claim_stdout():
    """Called by main() only (importing this module must not touch the importer's descriptors)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
METRIC = "images/sec post-backbone (decode+NMS+PRN)"
UNIT = "images/s"
D = 56 * 36 * 17
HIDDEN = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--prn-mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank runs the workload's batch; strong: the workload's batch is split over the ranks "
                         "(BASELINE configs[3]: 1024x1024, 64 images over 2/4/8 GPUs)")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="the K-step timed block is repeated until this much device time has been measured")
    ap.add_argument("--lanes", type=int, default=0,
                    help="Detector handles / CUDA streams fed round-robin in the device-resident leg (DetectorLanes); "
                         "0 = what was measured best: 2 for calls of latency-bound kernels (c1, c2; c2 with 1 / 2 / 3 / 4 lanes: "
                         "79.9 / 73.5 / 79.1 / 75.5 us per batch, profiles/r02w_lanes.txt), 3 for crowded / large calls (c3: 840 / "
                         "854 / 812 us with 1 / 2 / 3; c4: 421 / 427 / 401)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        self.marks = []

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        self.marks.append((t0, t1))

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if any(a - 0.05 <= t <= b + 0.15 for a, b in self.marks)] or \
               [r for (_, r) in self.rows]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- workload
def make_ring(wl, n_sets, rank):
    from multiposenet_b200 import synthetic
    return [synthetic.make_inputs(wl, replicate=100 * rank + r) for r in range(n_sets)]


def set_bytes(inp):
    return sum(int(inp[k].nbytes) for k in ("class_logits", "encoded_boxes", "heatmap_logits"))


def algorithmic_bytes(kernel, wl, B, n_persons, n_cand, mode):
    """(compulsory, moved) HBM bytes of one launch.  `compulsory` follows SURVEY.md section 8(d): every input byte once,
    every output byte once, nothing that exists only because of how the path is cut into kernels (None for a kernel that
    is such an extra pass); `moved` is what the kernel as built has to read + write ('Kernels and their
    rooflines')."""
    A, pix = wl.num_anchors, (wl.height // 4) * (wl.width // 4)
    w2 = 2 * D * HIDDEN * (2 if mode == "bf16" else 4) + (HIDDEN + D) * 4          # both weight matrices + biases
    table = {
        "candidates_flat": (4 * A * B + 8 * n_cand, 4 * A * B + 8 * n_cand),
        "sort_nms": (24 * n_cand + B * wl.max_detections * 20, 24 * n_cand + B * wl.max_detections * 40),
        "heatmap": (144 * pix * B, 144 * pix * B),                        # one-pass kernel (per-tap crop path)
        "logit_minmax": (None, 72 * pix * B),                             # extra pass over the logits (L2-resident re-read later)
        "heatmap_norm": (144 * pix * B, (144 + 80) * pix * B),            # + the padded normalised map (workspace)
        "crop": (n_persons * D * 4, n_persons * D * (4 + (2 if mode == "bf16" else 0))),
        # section 8(d): weights once + (unfused) crops in and logits out; the bf16 copy of x is an artefact of the cut
        "prn_fused": (None, None) if n_persons > 256 else (w2 + n_persons * D * 8, w2 + n_persons * D * 10),
        "keypoint_decode": (n_persons * D * 4, n_persons * D * 4),
    }
    return table.get(kernel, (None, None))


# ------------------------------------------------------------------------------------------------- CPU leg
def cpu_port_runner(wl, weights):
    """One batch through the oracle port (CPU restatement of the reference path); BLAS (torch CPU, all host
    threads) for the two dense layers, OpenMP C for the rest."""
    import torch

    import oracle
    torch.set_num_threads(host_cores())
    W1, b1, W2, b2 = [torch.from_numpy(np.ascontiguousarray(a)) for a in weights]

    def prn_fn(x):
        xt = torch.from_numpy(x)
        y1 = torch.relu(torch.addmm(b1, xt, W1))
        return (xt + torch.relu(torch.addmm(b2, y1, W2))).numpy()

    def run(inp):
        return oracle.full_path(inp["class_logits"], inp["encoded_boxes"], inp["heatmap_logits"], wl.height, wl.width,
                                *weights, thr=wl.score_threshold, iou_thr=wl.iou_threshold, max_det=wl.max_detections,
                                multipliers=wl.multipliers, ratios=wl.ratios, prn_fn=prn_fn)
    return run


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_leg(wl, weights, ring, seconds, max_steps=None, warmup=2, run=None):
    run = run or cpu_port_runner(wl, weights)
    for i in range(warmup):
        run(ring[i % len(ring)])
    n, t0 = 0, time.perf_counter()
    times = []
    while True:
        t = time.perf_counter()
        run(ring[n % len(ring)])
        times.append(time.perf_counter() - t)
        n += 1
        if (max_steps is not None and n >= max_steps) or (max_steps is None and time.perf_counter() - t0 >= seconds):
            break
    total = sum(times)
    B = ring[0]["class_logits"].shape[0]
    return {"value": B * n / total, "unit": UNIT, "cores": host_cores(), "kind": "port",
            "sample": f"{n} batches of {B} images ({wl.name}), {total:.1f} s of CPU work; oracle C port + torch CPU "
                      f"BLAS for the PRN (TensorFlow 1.15 not installable: restatement, not TF)",
            "ms_per_step": 1e3 * total / n, "steps": n}


def batch_per_gpu(wl, args):
    """weak: every rank runs the workload's batch; strong: the batch is split over the ranks (contiguous image blocks)."""
    if args.scaling == "strong":
        if wl.batch % args.gpus != 0:
            raise SystemExit(f"--scaling strong: batch {wl.batch} is not divisible by {args.gpus} GPUs")
        return wl.batch // args.gpus
    return wl.batch


def reference_arm(args, wl, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow 1.15 (the reference's only
    backend) cannot be installed here (no network, no cp312 wheel), so this times the oracle port.  One host CPU runs the
    per-GPU batch of the other arm (same config); at N > 1 rank 0 alone runs it."""
    if rank != 0:
        return
    from multiposenet_b200 import synthetic
    weights = synthetic.make_prn_weights()
    B = batch_per_gpu(wl, args)
    ring = [synthetic.make_inputs(wl, replicate=r, batch=B) for r in range(2)]
    steps = max(1, min(args.steps, 400))
    # warm for at least a second (thread pools, page faults, BLAS buffers: 20 cold steps once read 584 instead of ~690 images/s)
    run = cpu_port_runner(wl, weights)
    t0, n_warm = time.perf_counter(), 0
    while n_warm < max(1, min(args.warmup, 5)) or time.perf_counter() - t0 < 1.0:
        run(ring[n_warm % 2])
        n_warm += 1
    res = run_cpu_leg(wl, weights, ring, seconds=0, max_steps=steps, warmup=0, run=run)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "warmup_steps_run": n_warm, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(wl, args, batch_per_gpu=B),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config_dict(wl, args, batch_per_gpu=None):
    """The workload description -- identical for both arms (the driver compares it)."""
    return {"workload": f"BASELINE configs[{wl.config_id - 1}]: {wl.name}", "image": [wl.height, wl.width],
            "batch_per_gpu": wl.batch if batch_per_gpu is None else batch_per_gpu,
            "anchors_per_location": wl.n_loc, "anchors_per_image": wl.num_anchors,
            "score_threshold": wl.score_threshold, "iou_threshold": wl.iou_threshold,
            "max_detections": wl.max_detections, "prn": args.prn_mode,
            "parallelism": f"dp{args.gpus} (image-sharded, no collective)",
            "l2": "inputs rotate over distinct batches whose total size exceeds the 126 MB L2, no flush"}


# ------------------------------------------------------------------------------------------------- GPU legs
def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from multiposenet_b200 import synthetic
    wl = synthetic.WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from multiposenet_b200 import DetectorConfig, DetectorLanes, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # the host-side result gather of the e2e leg (no device collective)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = args.steps, max(args.warmup, 3)
    B = batch_per_gpu(wl, args)
    weights = synthetic.make_prn_weights()
    n_lanes = args.lanes if args.lanes > 0 else (2 if wl.batch * wl.max_detections <= 256 else 3)
    lanes = DetectorLanes(weights, DetectorConfig(
        max_batch=B, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
        score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold, scale_multipliers=wl.multipliers,
        aspect_ratios=wl.ratios, prn_mode=args.prn_mode, prn_modes_allocated=(args.prn_mode,), device=local_rank),
        lanes=n_lanes)
    det = lanes.detectors[0]            # the single-lane, e2e and profiling legs run on this handle

    # weak: rank r runs its own batches; strong: rank r runs images [r B, (r + 1) B) of the workload's batches
    def make_set(r):
        if args.scaling == "strong":
            full = synthetic.make_inputs(wl, replicate=r)
            return parallel.shard_inputs(full, rank, world)[0]
        return synthetic.make_inputs(wl, replicate=100 * rank + r)

    probe = make_set(0)
    n_sets = max(2, -(-int(1.3 * L2_BYTES) // set_bytes(probe)))
    ring = [probe] + [make_set(r) for r in range(1, n_sets)]
    names = ("encoded_boxes", "class_logits", "heatmap_logits")
    dev_ring = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).to(dev) for k in names} for s in ring]
    pin_ring = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).pin_memory() for k in names} for s in ring]

    def dev_step(i):
        s = dev_ring[i % n_sets]
        return det.run_device(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], (wl.height, wl.width))

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- leg 1: device-resident ------------------------------------------------------------------------------
    # on side streams: the library replays one CUDA graph per distinct (inputs, outputs) description there (the
    # legacy default stream cannot be captured)
    side = torch.cuda.Stream(device=dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed_blocks(step, prologue=None, epilogue=None):
        """EXACTLY K steps between two CUDA events on `side`, barrier + synchronize on both sides, max over ranks.  The
        block is repeated until --min-seconds of device time have been measured (a 20-step block of this path lasts
        1.6 ms); the median block is the reported one, the fastest is kept beside it."""
        blocks, launches, first = [], 0, Wm
        while True:
            barrier()
            launches0 = lanes.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            with torch.cuda.stream(side):
                e0.record()
                if prologue:
                    prologue(e0)
                for i in range(K):
                    out = step(first + i)
                if epilogue:
                    epilogue()
                e1.record()
            barrier()
            t1 = time.time()
            if sampler:
                sampler.mark(t0, t1)
            launches = int(lanes.launch_count() - launches0)
            blocks.append(max_over_ranks(e0.elapsed_time(e1)))
            first += K
            # every rank must take the same decision: the block times are already the max over ranks
            if sum(blocks) >= 1e3 * args.min_seconds or len(blocks) >= 200:
                break
        return statistics.median(blocks), min(blocks), len(blocks), launches, out

    # (a) one handle, one stream: K calls back to back
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for i in range(max(Wm, n_sets)):
            out = dev_step(i)
    single_ms, single_min_ms, single_blocks, n_launches, out = timed_blocks(dev_step)
    persons = int(out["person_offsets"][-1].item())

    # (b) the K steps fed round-robin to the lanes; the events sit on `side`, every lane starts after e0 and `side`
    # waits for every lane before e1
    def lane_step(i):
        s = dev_ring[i % n_sets]
        return lanes.submit(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], (wl.height, wl.width))

    dev_ms, dev_min_ms, dev_blocks = single_ms, single_min_ms, single_blocks
    if n_lanes > 1:
        with torch.cuda.stream(side):
            lanes.fork()
            for i in range(max(Wm, n_sets * n_lanes)):      # every lane sees every input set once (graph capture)
                lane_step(i)
            lanes.join()
        dev_ms, dev_min_ms, dev_blocks, n_launches, _ = timed_blocks(lane_step, prologue=lanes.fork, epilogue=lanes.join)
    if world > 1:
        t = torch.tensor([n_launches], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n_launches = int(t.item())

    # ---- leg 2: end to end through host buffers --------------------------------------------------------------
    # every step copies its inputs from pinned host memory, runs the path and copies all seven outputs back; up to
    # HOST_DEPTH steps are in flight so that the PCIe copies of neighbouring steps overlap the kernels.  At N > 1 every
    # step's detections and keypoints are then gathered on rank 0 in rank order over the HOST (gloo: fixed-size blocks, no
    # device collective -- north_star's "host-side result gather"); the heatmaps stay with the rank that produced them.
    from multiposenet_b200._lib import HOST_DEPTH

    def consume(bufs):
        _ = int(bufs["num_boxes"][0])                  # the step's result is read on the host
        if world > 1:
            parallel.gather_packed(bufs, dst=0, group=host_group)

    def host_loop(n, first, depth):
        pend = []
        for i in range(n):
            s = pin_ring[(first + i) % n_sets]
            pend.append(det.submit_host(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"],
                                        (wl.height, wl.width), return_heatmaps=True))
            if len(pend) >= depth:
                t, bufs = pend.pop(0)
                det.wait(t)
                consume(bufs)
        while pend:
            t, bufs = pend.pop(0)
            det.wait(t)
            consume(bufs)
        return bufs

    host_loop(Wm, 0, HOST_DEPTH)
    # Passes of K steps each until --min-seconds (at least two); the MEDIAN pass is reported, all of them are in the line:
    # the pinned copies share the host's memory system with whatever else runs on the machine.
    e2e_passes = []
    while len(e2e_passes) < 2 or (sum(e2e_passes) < args.min_seconds and len(e2e_passes) < 50):
        barrier()
        t0 = time.time()
        w0 = time.perf_counter()
        hout = host_loop(K, Wm, HOST_DEPTH)
        torch.cuda.synchronize()
        pass_s = time.perf_counter() - w0
        t1 = time.time()
        if sampler:
            sampler.mark(t0, t1)
        if world > 1:
            t = torch.tensor([pass_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pass_s = float(t.item())
        e2e_passes.append(pass_s)
    e2e_s = statistics.median(e2e_passes)
    del hout
    w0 = time.perf_counter()
    host_loop(min(K, 50), 0, 1)                        # one call at a time: the latency of a single step
    serial_ms = 1e3 * (time.perf_counter() - w0) / min(K, 50)
    h2d, d2h = det.host_traffic()
    in_bytes = set_bytes(ring[0])

    clocks = sampler.stop() if sampler else None

    # ---- leg 3: per-kernel times (profiling events; not part of `value`) ---------------------------------------
    roofline, kernels, stages = None, None, None
    if rank == 0:
        det.set_profiling(True)
        acc, order = {}, []
        n_cand = 0
        for i in range(min(K, 50)):
            dev_step(i)
            for name, ms in det.profile():
                if name not in acc:
                    acc[name] = []
                    order.append(name)
                acc[name].append(ms)
        det.set_profiling(False)
        s = dev_ring[0]
        dstat = det.detect(s["encoded_boxes"], s["class_logits"], (wl.height, wl.width))
        n_cand = int(dstat["n_candidates"].sum().item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        kernels = {}
        mean = {n: statistics.mean(acc[n]) for n in order}
        step_ms = sum(mean.values())
        for name in order:
            ms = mean[name]
            ab, mv = algorithmic_bytes(name, wl, B, persons, n_cand, args.prn_mode)
            kernels[name] = {"ms": round(ms, 5), "share": round(ms / step_ms, 4), "alg_bytes": ab, "moved_bytes": mv,
                             "gbs": None if not ab else round(ab / ms / 1e6, 1),
                             "gbs_moved": None if not mv else round(mv / ms / 1e6, 1),
                             "frac_of_hbm_peak": None if not mv else round(mv / ms / 1e6 / hbm_peak, 3)}
            if name in ("prn_big_fc1", "prn_big_fc2", "prn_bf16_fc1", "prn_bf16_fc2"):      # tensor-bound layers: 2 N D H flop
                kernels[name]["tflops"] = round(2.0 * persons * D * HIDDEN / ms / 1e9, 1)
        # the heatmap stage as a whole (create_pb.py:73-76, 90-94): logits in once, the two outputs out once
        hm = [n for n in ("logit_minmax", "heatmap_norm", "heatmap") if n in mean]
        if hm:
            t = sum(mean[n] for n in hm)
            pix = (wl.height // 4) * (wl.width // 4)
            stages = {"heatmap_stage": {"kernels": hm, "ms": round(t, 5), "alg_bytes": 144 * pix * B,
                                        "gbs": round(144 * pix * B / t / 1e6, 1)}}
        cands = [n for n in order if kernels[n]["alg_bytes"]]
        top = max(cands, key=lambda n: mean[n])
        ab, _ = algorithmic_bytes(top, wl, B, persons, n_cand, args.prn_mode)
        traffic = None
        if args.workload == "c2":          # the committed ncu capture was taken at this workload's sizes
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top)
            except Exception:
                pass
        launch_ms = mean[top]
        timing = "CUDA events between direct launches (profiling pass)"
        if top == "prn_fused":
            # the dominant kernel alone: K replays of a graph that contains only this kernel (every other stage skipped,
            # its inputs -- the crops of the last full step -- stay in place), CUDA events on the launching stream
            # (the skipped stages are the only readers of the rotating inputs, so a few sets are enough: a workload with
            # a long input ring would otherwise spend this pass capturing one new graph per set)
            det.debug_skip(1 | 2 | 8 | 32)
            m_sets = min(n_sets, 7)
            reps = max(min(K, 200), 100)
            with torch.cuda.stream(side):
                for i in range(max(10, m_sets)):
                    dev_step(i % m_sets)
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for i in range(reps):
                    dev_step(i % m_sets)
                p1.record()
            torch.cuda.synchronize()
            det.debug_skip(0)
            launch_ms = p0.elapsed_time(p1) / reps
            timing = f"CUDA events around {reps} graph replays that contain only this kernel (all other stages skipped)"
            # for explanation only: the kernel's own duration INSIDE whole calls, from its globaltimer stamps (first CTA
            # past its prologue -> last logit stored; mpn_debug_fused_trace) -- what a replay adds to that is launch latency
            try:
                det.fused_trace(True)
                spans = []
                with torch.cuda.stream(side):
                    for i in range(3 * m_sets):
                        dev_step(i % m_sets)
                        side.synchronize()
                        tr = det.fused_trace(True).astype("int64")
                        if (tr[:, 10] > 0).any():
                            spans.append((tr[:, 10].max() - tr[:, 0][tr[:, 0] > 0].min()) / 1e6)
                det.fused_trace(False)
                in_kernel_ms = float(sorted(spans)[len(spans) // 2]) if spans else None
            except Exception:
                in_kernel_ms = None
        ach = ab / launch_ms / 1e6
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(ach / hbm_peak, 4), "traffic": traffic, "peak_source": which,
                    "alg_bytes_per_launch": ab, "alg_bytes_formula": "SURVEY 8(d): PRN weights + biases once + persons x 274176 B "
                    "(fp32 crops in, logits out)" if top == "prn_fused" else "DESIGN.md section 4",
                    "avg_launch_ms": round(launch_ms, 5), "timing": timing,
                    "avg_launch_ms_profiling_pass": round(mean[top], 5),
                    "persons_per_batch": persons, "candidates_per_batch": n_cand}
        if top == "prn_fused" and in_kernel_ms:
            roofline["in_kernel_ms_inside_a_call"] = round(in_kernel_ms, 5)
            roofline["frac_in_kernel"] = round(ab / in_kernel_ms / 1e6 / hbm_peak, 4)

    # ---- leg 4: CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = run_cpu_leg(wl, weights, ring, seconds=args.cpu_seconds)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    lanes.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    images = B * K * world
    line = {
        "metric": METRIC, "value": images / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms / K, "lanes": n_lanes,
        "timed_blocks": {"count": dev_blocks, "steps_per_block": K, "reported": "median block",
                         "fastest_block_ms_per_step": dev_min_ms / K, "fastest_block_value": images / (dev_min_ms / 1e3)},
        "single_lane": {"value": images / (single_ms / 1e3), "unit": UNIT, "ms_per_step": single_ms / K,
                        "fastest_block_ms_per_step": single_min_ms / K, "blocks": single_blocks,
                        "note": "the same K steps back to back on one handle and one stream"},
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32 (decode/NMS/crop/softmax) + " + ("bf16 tcgen05, f32 accumulate (PRN)" if args.prn_mode == "bf16" else "f32 (PRN)"),
        "data": "synthetic", "config": config_dict(wl, args, batch_per_gpu=B), "input_ring_sets": n_sets,
        "e2e": {"value": images / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / K, "passes_ms_per_step": [round(1e3 * x / K, 5) for x in e2e_passes],
                "reported": "median pass", "in_flight": HOST_DEPTH, "serial_ms_per_step": serial_ms,
                "input_bytes_per_step": in_bytes,
                "result_gather": None if world == 1 else "every step: detections + keypoints of all ranks gathered on rank 0 "
                                                         "over the host (gloo, fixed-size blocks)",
                "note": "class logits and heatmap logits are copied by DMA; the box codes stay in pinned host memory "
                        "and only the rows of confident anchors are gathered over PCIe by the NMS kernel"},
        "gpu_launches": n_launches,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels, "stages": stages,
    }
    emit(line)


if __name__ == "__main__":
    main()
