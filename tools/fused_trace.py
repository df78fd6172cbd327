"""Development aid: phase timeline of the single-kernel PRN (mpn_debug_fused_trace).  python tools/fused_trace.py [N]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

# python tools/fused_trace.py 78        the PRN stage alone on 78 synthetic crops
# python tools/fused_trace.py c2        whole calls of a workload (the kernel as mpn_run launches it; MPN_FUSE_CROP=1: with
#                                        crop_and_resize inside it)
arg = sys.argv[1] if len(sys.argv) > 1 else "78"
w = synthetic.make_prn_weights()
if arg.isdigit():
    n = int(arg)
    det = Detector(w, DetectorConfig(max_batch=8, max_boxes=32, prn_mode="bf16", prn_modes_allocated=("bf16",)))
    x = torch.from_numpy(synthetic.make_crops(n)).cuda()
    call = lambda: det.prn(x, "bf16", inplace=True)
else:
    wl = synthetic.WORKLOADS[arg]
    inp = synthetic.make_inputs(wl)
    det = Detector(w, DetectorConfig(max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
                                     score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold,
                                     scale_multipliers=wl.multipliers, prn_mode="bf16", prn_modes_allocated=("bf16",)))
    dev_in = [torch.from_numpy(inp[k]).cuda() for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    n = int(det.run_device(*dev_in)["person_offsets"][-1])
    call = lambda: det.run_device(*dev_in)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for _ in range(3):
    call()
det.fused_trace(True)
# wave B slots stay empty in the one-wave build of the kernel (MPN_FC1_WAVES, csrc/prn_fused.cu)
names = {7: "pdl wait passed", 0: "prologue", 1: "fc1 loads issued", 2: "fc1 A mma issued", 14: "fc1 B mma issued",
         3: "fc1 A acc complete", 4: "partials A stored", 5: "barrier A1 passed", 8: "y1 A stored",
         13: "fc1 B acc / fp32 crops", 11: "partials B / crop chunk 0", 12: "barrier B1 / last chunk", 15: "y1 B stored",
         6: "producer saw last y1", 9: "fc2 acc complete", 10: "logits stored"}
for rep in range(3):
    # many launches back to back so that the SM clock is at its loaded value; the trace holds the last launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(300):
        if i == 100:
            e0.record()
        call()
    e1.record(); torch.cuda.synchronize()
    t = det.fused_trace(True).astype(np.int64)
    t0 = t[:, 0].min()
    print(f"--- rep {rep}: N={n}, {e0.elapsed_time(e1) * 1e3 / 200:.1f} us per call (PRN stage or whole call)")
    for slot in (0, 7, 2, 3, 4, 5, 8, 1, 14, 13, 11, 12, 15, 6, 9, 10):
        col = t[:, slot]
        col = col[col > 0] - t0
        if col.size:
            print(f"  {names[slot]:24s} min {col.min() / 1e3:7.2f}  median {np.median(col) / 1e3:7.2f}  max {col.max() / 1e3:7.2f} us  ({col.size} CTAs)")
if os.environ.get("MPN_TRACE_PER_CTA"):
    # per-CTA view of the last launch: when did every CTA's fc1 accumulators complete / its partial sums leave / its fc2
    # accumulators complete (us after the first prologue stamp), in blockIdx order -- is the spread systematic?
    t = det.fused_trace(True).astype(np.int64)
    t0 = t[:, 0].min()
    smid = t[:, 13] if t.shape[1] > 13 else None
    print("cta hq z   fc1_done  partials_out  fc2_done")
    for cta in range(t.shape[0]):
        row = [(t[cta, k] - t0) / 1e3 if t[cta, k] > 0 else float("nan") for k in (3, 4, 9)]
        print(f"{cta:3d} {cta % 4:2d} {cta // 4:2d}  {row[0]:8.2f}  {row[1]:8.2f}  {row[2]:8.2f}")
det.fused_trace(False)
det.close()
