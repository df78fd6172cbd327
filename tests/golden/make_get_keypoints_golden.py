"""Generates tests/golden/get_keypoints.npz by running the REFERENCE's own
inference/utils.py::get_keypoints (the only function on the path that imports
without TensorFlow).  Run in the build container only (needs /root/reference):

    python tests/golden/make_get_keypoints_golden.py
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/inference/utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "get_keypoints.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_inference_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.Generator(np.random.PCG64(20240915))
    cases = {}
    k = 0
    for (h, w) in [(56, 36), (40, 24), (17, 9)]:
        for int_box in (True, False):
            for thr in (0.1, 0.5, 0.9):
                hm = (rng.integers(0, 40, (h, w, 17)).astype(np.float32) / np.float32(64.0))  # coarse values: small npz
                # plant peaks in most channels, leave some below the threshold, duplicate one maximum (first-index rule)
                for c in range(17):
                    if c % 5 == 4:
                        continue
                    y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
                    hm[y, x, c] = 0.95
                    if c % 6 == 0:
                        y2, x2 = int(rng.integers(0, h)), int(rng.integers(0, w))
                        hm[y2, x2, c] = 0.95
                if int_box:
                    y0, x0 = int(rng.integers(0, 200)), int(rng.integers(0, 200))
                    box = np.array([y0, x0, y0 + int(rng.integers(20, 400)), x0 + int(rng.integers(20, 300))], np.int64)
                else:
                    y0, x0 = rng.random() * 200, rng.random() * 200
                    box = np.array([y0, x0, y0 + 20 + rng.random() * 380, x0 + 20 + rng.random() * 280], np.float64)
                out = mod.get_keypoints(hm, box, thr)
                cases[f"hm_{k}"] = hm
                cases[f"box_{k}"] = box
                cases[f"thr_{k}"] = np.float64(thr)
                cases[f"out_{k}"] = out
                k += 1
    cases["n"] = np.int64(k)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, "cases", k)


if __name__ == "__main__":
    main()
