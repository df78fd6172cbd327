"""BASELINE configs[4]: PRN-only sweep, bf16 tcgen05 vs fp32 SIMT, N persons.  python tools/prn_sweep.py [N ...]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

Ns = [int(a) for a in sys.argv[1:]] or [16, 78, 256, 1000, 4000, 10000]
cap = max(Ns)
w = synthetic.make_prn_weights()
det = Detector(w, DetectorConfig(max_batch=(cap + 127) // 128, max_boxes=128, prn_mode="bf16"))
rng = np.random.default_rng(0)
for n in Ns:
    x = torch.from_numpy(synthetic.make_crops(min(n, 2000))).cuda()
    if n > 2000:
        x = x.repeat((n + 1999) // 2000, 1, 1, 1)[:n].contiguous()
    for mode in ("bf16", "fp32"):
        if mode == "fp32" and n > 4000:
            continue
        # bf16: in place (logits overwrite the crops, as in the full path); the crops are restored outside the timed region
        inplace = mode == "bf16"
        work = x.clone() if inplace else x
        for _ in range(3):
            if inplace:
                work.copy_(x)
            out = det.prn(work, mode, inplace=inplace)
        torch.cuda.synchronize()
        reps = 20 if n <= 1000 else 5
        ms = 0.0
        for _ in range(reps):
            if inplace:
                work.copy_(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = det.prn(work, mode, inplace=inplace)
            e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1) / reps
        del work
        tf = n * 140378112 / ms / 1e9
        print(f"N={n:6d} {mode}: {ms * 1e3:10.1f} us  {n / ms * 1e3:12.0f} persons/s  {tf:8.1f} TFLOP/s  "
              f"(weights-only stream {(281 if mode == 'fp32' else 140.4) / ms / 1e3:6.2f} TB/s)")
    del x, out
det.close()
