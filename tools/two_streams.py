"""Development aid: device-resident throughput with D Detector handles driven round-robin on D streams (calls of
neighbouring batches overlap: one call's front half runs beside another call's tail).  python tools/two_streams.py [workload] [D ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
depths = [int(a) for a in sys.argv[2:]] or [1, 2, 3]
w = synthetic.make_prn_weights()
cfg = DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
                     score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold, scale_multipliers=wl.multipliers,
                     aspect_ratios=wl.ratios, prn_mode="bf16", prn_modes_allocated=("bf16",))
probe = synthetic.make_inputs(wl, replicate=0)
nbytes = sum(int(probe[k].nbytes) for k in ("class_logits", "encoded_boxes", "heatmap_logits"))
n_sets = max(2, -(-int(1.3 * 126 * 2**20) // nbytes))
ring = [probe] + [synthetic.make_inputs(wl, replicate=r) for r in range(1, n_sets)]
dev_ring = [{k: torch.from_numpy(s[k]).cuda() for k in ("encoded_boxes", "class_logits", "heatmap_logits")} for s in ring]
for D in depths:
    dets = [Detector(w, cfg) for _ in range(D)]
    streams = [torch.cuda.Stream() for _ in range(D)]
    def step(i):
        s = dev_ring[i % n_sets]
        with torch.cuda.stream(streams[i % D]):
            return dets[i % D].run_device(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], (wl.height, wl.width))
    for i in range(max(50, n_sets * D)):
        step(i)
    torch.cuda.synchronize()
    K = 600
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()                       # default stream; the side streams were idle at this point
        for i in range(K):
            step(i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        best = us if best is None else min(best, us)
    print(f"{wl.name.split(':')[0]}: {D} handle(s) / stream(s): {best:7.2f} us per step  {wl.batch / best * 1e6:9.0f} images/s", flush=True)
    for d in dets:
        d.close()
