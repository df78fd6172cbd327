"""Development aid: per-kernel times of the large-batch PRN (prn_big.cu).  python tools/fc2_bench.py [N ...]
MPN_FC2_PAIRS=0 selects the single-CTA fc2 kernel."""
import hashlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiposenet_b200 import Detector, DetectorConfig, synthetic

INPLACE = os.environ.get("INPLACE", "0") == "1"
Ns = [int(a) for a in sys.argv[1:]] or [600, 2801, 10000]
cap = max(Ns)
det = Detector(synthetic.make_prn_weights(), DetectorConfig(max_batch=(cap + 127) // 128, max_boxes=128, prn_mode="bf16",
                                                            prn_modes_allocated=("bf16",)))
base = torch.from_numpy(synthetic.make_crops(1000)).cuda()
for n in Ns:
    x = base.repeat((n + 999) // 1000, 1, 1, 1)[:n].contiguous()
    for _ in range(3):
        out = det.prn(x.clone() if INPLACE else x, "bf16", inplace=INPLACE)
    torch.cuda.synchronize()
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:12]
    det.set_profiling(True)
    acc = {}
    for _ in range(5):
        xi = x.clone() if INPLACE else x
        det.prn(xi, "bf16", inplace=INPLACE)
        for name, ms in det.profile():
            acc.setdefault(name, []).append(ms)
    det.set_profiling(False)
    parts = "  ".join(f"{k} {1e3 * np.mean(v):8.1f} us" for k, v in acc.items())
    fc2 = np.mean(acc.get("prn_big_fc2", [float("nan")]))
    print(f"N={n:6d} pairs={os.environ.get('MPN_FC2_PAIRS', '1')} inplace={int(INPLACE)}  {parts}  | fc2 {n * 2 * 34272 * 1024 / fc2 / 1e9:7.1f} TFLOP/s  out {digest}")
    del x, out
det.close()
