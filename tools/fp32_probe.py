import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from multiposenet_b200 import Detector, DetectorConfig, synthetic
wl = synthetic.WORKLOADS["c2"]
inp = synthetic.make_inputs(wl)
w = synthetic.make_prn_weights()
for mb in (10, 25):
    det = Detector(w, DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=mb,
                                     score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold,
                                     scale_multipliers=wl.multipliers, prn_mode="fp32", prn_modes_allocated=("fp32",)))
    dev_in = [torch.from_numpy(inp[k]).cuda() for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(20): out = det.run_device(*dev_in)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(300): out = det.run_device(*dev_in)
        e1.record(side)
    torch.cuda.synchronize()
    print(f"max_boxes {mb}: N = {int(out['person_offsets'][-1])}, {e0.elapsed_time(e1) / 300 * 1e3:.1f} us per call, launches per call {det.launch_count()[0]}")
    det.close()
