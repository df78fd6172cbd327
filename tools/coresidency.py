"""Development aid: do the kernels of one call run BESIDE the fused PRN kernel of another call?  Handle A replays a graph
that contains only the PRN kernel, handle B one that contains only the stages named on the command line; each is timed
alone and then both together on two streams.  together ~ max(A, B): they share the SMs; together ~ A + B: they do not.
    python tools/coresidency.py [stages ...]     stages: detect heatmap crop decode (default: each in turn)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

BITS = {"detect": 1, "heatmap": 2 | 4, "crop": 8, "prn": 16, "decode": 32}
ALL = 63
wl = synthetic.WORKLOADS["c2"]
w = synthetic.make_prn_weights()
cfg = DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
                     score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold, scale_multipliers=wl.multipliers,
                     aspect_ratios=wl.ratios, prn_mode="bf16", prn_modes_allocated=("bf16",))
inp = synthetic.make_inputs(wl, replicate=0)
d = {k: torch.from_numpy(inp[k]).cuda() for k in ("encoded_boxes", "class_logits", "heatmap_logits")}
A, B = Detector(w, cfg), Detector(w, cfg)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def call(det, st):
    with torch.cuda.stream(st):
        det.run_device(d["encoded_boxes"], d["class_logits"], d["heatmap_logits"], (wl.height, wl.width))


def timed(fn, K=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_event(e0); sb.wait_event(e0)
    for _ in range(K):
        fn()
    cur = torch.cuda.current_stream()
    cur.wait_stream(sa); cur.wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


for det, st in ((A, sa), (B, sb)):      # one full call each: crops, person list, heatmaps exist
    call(det, st)
torch.cuda.synchronize()
A.debug_skip(ALL & ~BITS["prn"])
ta = timed(lambda: call(A, sa))
print(f"PRN alone: {ta:6.2f} us", flush=True)
for name in (sys.argv[1:] or ["detect", "heatmap", "crop", "decode", "detect+heatmap+crop"]):
    keep = 0
    for part in name.split("+"):
        keep |= BITS[part]
    B.debug_skip(ALL & ~keep)
    tb = timed(lambda: call(B, sb))
    tab = timed(lambda: (call(A, sa), call(B, sb)))
    print(f"{name:22s} alone {tb:6.2f} us   beside the PRN {tab:6.2f} us   (sum {ta + tb:6.2f}, max {max(ta, tb):6.2f})", flush=True)
