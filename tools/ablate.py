"""Development aid: marginal cost of every stage inside the replayed graph (skip one stage at a time, time the step).
python tools/ablate.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
w = synthetic.make_prn_weights()
det = Detector(w, DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
                                 score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold,
                                 scale_multipliers=wl.multipliers, prn_mode="bf16", prn_modes_allocated=("bf16",)))
ring = [synthetic.make_inputs(wl, replicate=r) for r in range(7)]
dev = [{k: torch.from_numpy(s[k]).cuda() for k in ("encoded_boxes", "class_logits", "heatmap_logits")} for s in ring]
side = torch.cuda.Stream()
def step_us(mask, K=300):
    det.debug_skip(mask)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for i in range(20):
            s = dev[i % 7]; det.run_device(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            s = dev[i % 7]; det.run_device(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"])
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K
full = step_us(0)
print(f"full step {full:7.1f} us")
for name, bit in (("detect branch", 1), ("heatmap stage", 2), ("crop", 8), ("prn", 16),
                  ("decode", 32), ("all but prn", 47), ("all but front", 56), ("nothing launched", 63)):
    t = step_us(bit)
    print(f"without {name:18s} {t:7.1f} us   (saves {full - t:6.1f})")
det.debug_skip(0)
det.close()
