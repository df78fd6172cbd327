"""Generates tests/golden/metrics.npz by running the REFERENCE's own metrics.py (evaluate_detector and the Evaluator's
accumulation methods).  metrics.py imports tensorflow at module level only for get_metric_ops; a stub module lets the
import succeed, no TensorFlow function is called.  Run in the build container only (needs /root/reference):

    python tests/golden/make_metrics_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/metrics.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.npz")


def main():
    tf = types.ModuleType("tensorflow"); compat = types.ModuleType("tensorflow.compat"); v1 = types.ModuleType("tensorflow.compat.v1")
    tf.compat = compat; compat.v1 = v1
    sys.modules.update({"tensorflow": tf, "tensorflow.compat": compat, "tensorflow.compat.v1": v1})
    spec = importlib.util.spec_from_file_location("ref_metrics", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.Generator(np.random.PCG64(20241018))
    cases = {}
    k = 0
    for n_images, noise, extra, ties in [(1, 0.02, 0, False), (3, 0.05, 2, False), (6, 0.15, 4, True), (4, 0.4, 6, False),
                                         (2, 0.0, 0, True), (5, 0.08, 3, False)]:
        ev = mod.Evaluator()
        gts, dets, scs = [], [], []
        for i in range(n_images):
            P = int(rng.integers(0, 7))
            h = rng.uniform(0.1, 0.5, P); w = rng.uniform(0.05, 0.3, P)
            cy = rng.uniform(0.2, 0.8, P); cx = rng.uniform(0.2, 0.8, P)
            gt = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32).reshape(-1, 4)
            keep = rng.random(P) > 0.2
            det = gt[keep] + rng.normal(0, noise, (int(keep.sum()), 4)).astype(np.float32)
            dup = det[: min(len(det), 2)] + np.float32(0.01)                      # duplicates of already matched boxes
            fp = rng.uniform(0, 1, (extra, 4)).astype(np.float32); fp[:, 2:] = fp[:, :2] + rng.uniform(0.05, 0.3, (extra, 2)).astype(np.float32)
            det = np.concatenate([det, dup, fp], 0).astype(np.float32)
            sc = rng.uniform(0.3, 1.0, len(det)).astype(np.float32)
            if ties and len(sc) > 2:
                sc[1] = sc[0]
            ev.add_groundtruth(str(i), gt)
            ev.add_detections(str(i), det, sc)
            gts.append(gt); dets.append(det); scs.append(sc)
        ev.evaluate(0.5)
        cases[f"n_images_{k}"] = np.int64(n_images)
        for i in range(n_images):
            cases[f"gt_{k}_{i}"] = gts[i]; cases[f"det_{k}_{i}"] = dets[i]; cases[f"score_{k}_{i}"] = scs[i]
        cases[f"metrics_{k}"] = np.array([float(ev.metrics[m]) for m in
                                          ("AP", "precision", "recall", "mean_iou_for_TP", "best_threshold", "total_FP", "total_FN")],
                                         np.float64)
        k += 1
    cases["n"] = np.int64(k)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, "cases", k)


if __name__ == "__main__":
    main()
