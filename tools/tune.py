"""Development aid: device-resident step time of mpn_run (BASELINE configs[1]) under MPN_TUNE_* / MPN_NO_* environment
settings, one fresh process per setting, with a digest of the outputs so that a setting that changes results is visible.

    python tools/tune.py [--workload c2] [--steps 600] "" "MPN_NO_PDL=1" "MPN_TUNE_L2_AHEAD=6,MPN_TUNE_W2_MODE=2" ...
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(workload, steps):
    sys.path.insert(0, ROOT)
    import torch
    from multiposenet_b200 import Detector, DetectorConfig, synthetic
    wl = synthetic.WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    det = Detector(synthetic.make_prn_weights(), DetectorConfig(
        max_batch=wl.batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
        score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold, scale_multipliers=wl.multipliers,
        aspect_ratios=wl.ratios, prn_mode="bf16", prn_modes_allocated=("bf16",), device=0))
    probe = synthetic.make_inputs(wl, replicate=0)
    nbytes = sum(int(probe[k].nbytes) for k in ("class_logits", "encoded_boxes", "heatmap_logits"))
    n_sets = max(2, -(-int(1.3 * 126 * 2**20) // nbytes))
    ring = [probe] + [synthetic.make_inputs(wl, replicate=r) for r in range(1, n_sets)]
    names = ("encoded_boxes", "class_logits", "heatmap_logits")
    dev_ring = [{k: torch.from_numpy(s[k]).to(dev) for k in names} for s in ring]

    def step(i):
        s = dev_ring[i % n_sets]
        return det.run_device(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], (wl.height, wl.width))

    side = torch.cuda.Stream(device=dev)
    digest = hashlib.sha256()
    with torch.cuda.stream(side):
        for i in range(n_sets):
            out = step(i)
            side.synchronize()
            for k in ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions", "person_offsets"):
                digest.update(out[k].cpu().numpy().tobytes())
        for i in range(50):
            step(i)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / steps
        best = us if best is None else min(best, us)
    print(json.dumps({"us_per_step": round(best, 2), "digest": digest.hexdigest()[:16],
                      "persons": int(out["person_offsets"][-1].item())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("settings", nargs="*")
    a = ap.parse_args()
    if a.child:
        child(a.workload, a.steps)
        return
    for spec in (a.settings or [""]):
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=", 1)
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--workload", a.workload, "--steps",
                            str(a.steps)], env=env, capture_output=True, text=True, timeout=600)
        last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        print(f"{spec or '(defaults)':60s} rc={r.returncode} {last}", flush=True)
        if r.returncode != 0:
            print(r.stderr[-1500:], flush=True)


if __name__ == "__main__":
    main()
