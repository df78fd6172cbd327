"""Development aid: phase timeline of the sort / NMS kernel (mpn_debug_nms_trace).  python tools/nms_trace.py [workload]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
det = Detector(None, DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
                                    score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold,
                                    scale_multipliers=wl.multipliers, prn_mode="bf16", prn_modes_allocated=("bf16",)))
inp = synthetic.make_inputs(wl)
enc, cls = torch.from_numpy(inp["encoded_boxes"]).cuda(), torch.from_numpy(inp["class_logits"]).cuda()
det.nms_trace(True)
names = ["CTA resident", "candidates complete", "keys sorted", "first chunk decoded", "all chunks resolved", "results published"] + \
        [f"chunk {i // 2} {'tested vs kept' if i % 2 == 0 else 'resolved'}" for i in range(10)]
for rep in range(3):
    for _ in range(20):
        out = det.detect(enc, cls, (wl.height, wl.width))
    torch.cuda.synchronize()
    t = det.nms_trace(True).astype(np.int64)[:wl.batch]
    t0 = t[:, 0].min()
    print(f"--- rep {rep}: {wl.name}; candidates per image {out['n_candidates'].cpu().numpy().tolist()[:8]}, kept {out['num_boxes'].cpu().numpy().tolist()[:8]}")
    for slot, nm in enumerate(names):
        col = t[:, slot]
        col = col[col > 0] - t0
        if col.size:
            print(f"  {nm:22s} min {col.min() / 1e3:8.2f}  median {np.median(col) / 1e3:8.2f}  max {col.max() / 1e3:8.2f} us  ({col.size} CTAs)")
det.nms_trace(False)
det.close()
