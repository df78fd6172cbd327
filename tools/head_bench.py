"""Development aid: SURVEY 8(f) row 2 -- fused heatmap head vs (library 1x1 conv + transpose) followed by mpn_heatmaps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

B, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 160, 160)
det = Detector(None, DetectorConfig(max_batch=B, max_height=4 * h, max_width=4 * w, prn_modes_allocated=("bf16",)))
x = torch.relu(torch.randn(B, 64, h, w, device="cuda"))
wt = torch.randn(64, 18, device="cuda") * 0.4
bias = torch.cat([torch.full((17,), -4.595, device="cuda"), torch.zeros(1, device="cuda")])
def unfused():
    lg = (torch.einsum("bchw,ck->bhwk", x, wt) + bias).contiguous()       # cuBLAS / cuDNN class library path
    return det.heatmaps(lg)
def fused():
    return det.heatmap_head(x, wt, bias, want_logits=False)
for name, fn in (("library 1x1 conv + transpose, then heatmap kernel", unfused), ("fused heatmap_head kernel", fused)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} {h}x{w}  {name:52s} {e0.elapsed_time(e1) * 10:8.1f} us")
det.close()
