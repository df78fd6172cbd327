"""Development aid: keypoint decode kernel alone.  python tools/decode_bench.py [N ...]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

ns = [int(a) for a in sys.argv[1:]] or [78, 200, 600, 2801]
det = Detector(synthetic.make_prn_weights(), DetectorConfig(max_batch=1, max_boxes=8, prn_mode="bf16",
                                                            prn_modes_allocated=("bf16",)))
for n in ns:
    x = torch.randn((n, 56, 36, 17), device="cuda")
    for _ in range(5):
        det.keypoint_decode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 100
    e0.record()
    for _ in range(K):
        det.keypoint_decode(x)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / K
    print(f"N={n:6d}  {us:8.1f} us per call  {n * 56 * 36 * 17 * 4 / us / 1e6:7.2f} TB/s")
det.close()
