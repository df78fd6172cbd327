"""Deterministic synthetic inputs at the graph cut of the hot path.

What a backbone + heads would hand to the post-processing of the reference
(create_pb.py:73-81): class logits [B,A], box codes [B,A,4] (both in the
anchor order of detector/box_predictor.py:53-90) and keypoint-subnet heatmap
logits [B,H/4,W/4,18] (detector/keypoint_subnet.py:49-58), plus random-init
PRN weights (detector/prn.py:19-22: variance_scaling_initializer defaults).

Pure numpy, host side; no dependency on the CUDA library or on the test oracle.
"""
import itertools
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

STRIDES = (8, 16, 32, 64, 128)
SCALES = (32, 64, 128, 256, 512)
MULT_6 = (1.0, 1.4142)                                   # detector/retinanet.py:41
MULT_9 = (1.0, 2.0 ** (1.0 / 3.0), 2.0 ** (2.0 / 3.0))   # "COCO-style 9 anchors/location" (BASELINE config 2)
RATIOS = (1.0, 2.0, 0.5)
SCALE_FACTORS = (10.0, 10.0, 5.0, 5.0)
PRIOR_LOGIT = -4.59511985013459                          # -log((1-p)/p), p = 0.01 (box_predictor.py:107-114)
D_PRN = 56 * 36 * 17
HIDDEN = 1024


@dataclass
class Workload:
    """One BASELINE.json configuration."""
    name: str
    height: int
    width: int
    batch: int
    multipliers: Tuple[float, ...] = MULT_6
    ratios: Tuple[float, ...] = RATIOS
    persons: Tuple[int, int] = (8, 12)          # planted persons per image, inclusive range
    score_threshold: float = 0.3
    iou_threshold: float = 0.5
    max_detections: int = 25
    config_id: int = 0
    strides: Tuple[int, ...] = STRIDES
    scales: Tuple[int, ...] = SCALES
    extra: dict = field(default_factory=dict)

    @property
    def n_loc(self):
        return len(self.multipliers) * len(self.ratios)

    @property
    def num_anchors(self):
        return sum(-(-self.height // s) * -(-self.width // s) for s in self.strides) * self.n_loc


WORKLOADS = {
    # BASELINE.json configs[0..3]; thresholds: create_pb.py:31-36 for c1, the config text for the others
    "c1": Workload("c1: 512x512 B=1 n_loc=6 ~10 persons", 512, 512, 1, MULT_6, RATIOS, (8, 12), 0.3, 0.6, 25, 1),
    "c2": Workload("c2: 640x640 B=8 n_loc=9 thr 0.3 iou 0.5", 640, 640, 8, MULT_9, RATIOS, (8, 12), 0.3, 0.5, 25, 2),
    "c2_n6": Workload("c2 (n_loc=6): 640x640 B=8", 640, 640, 8, MULT_6, RATIOS, (8, 12), 0.3, 0.5, 25, 2),
    "c3": Workload("c3: crowded 640x640 B=32 100+ persons", 640, 640, 32, MULT_9, RATIOS, (100, 140), 0.3, 0.5, 128, 3),
    "c4": Workload("c4: 1024x1024 B=64 n_loc=6", 1024, 1024, 64, MULT_6, RATIOS, (8, 12), 0.3, 0.5, 25, 4),
    "tiny": Workload("tiny: 256x256 B=2 (tests)", 256, 256, 2, MULT_6, RATIOS, (2, 4), 0.3, 0.6, 25, 9),
}


def seed_for(config_id, replicate=0):
    return 20240000 + 100 * int(config_id) + int(replicate)


def anchors_np(H, W, strides=STRIDES, scales=SCALES, multipliers=MULT_6, ratios=RATIOS):
    """Anchor boxes [A,4] (ymin,xmin,ymax,xmax, normalised) in the reference's order
    (detector/anchor_generator.py:53-116).  Used only to PLACE synthetic detections."""
    pairs = list(itertools.product(multipliers, ratios))
    out = []
    for s, S in zip(strides, scales):
        gh, gw = -(-H // s), -(-W // s)
        oy = 0.5 * (H - (gh - 1) * s)
        ox = 0.5 * (W - (gw - 1) * s)
        cy = (np.arange(gh) * s + oy)[:, None, None]
        cx = (np.arange(gw) * s + ox)[None, :, None]
        ah = np.array([m * S / np.sqrt(r) for m, r in pairs])[None, None, :]
        aw = np.array([m * S * np.sqrt(r) for m, r in pairs])[None, None, :]
        b = np.stack(np.broadcast_arrays((cy - 0.5 * ah) / H, (cx - 0.5 * aw) / W, (cy + 0.5 * ah) / H,
                                         (cx + 0.5 * aw) / W), -1)
        out.append(b.reshape(-1, 4))
    return np.concatenate(out, 0).astype(np.float32)


def _iou_matrix(a, b):
    ymin = np.maximum(a[:, None, 0], b[None, :, 0]); xmin = np.maximum(a[:, None, 1], b[None, :, 1])
    ymax = np.minimum(a[:, None, 2], b[None, :, 2]); xmax = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(ymax - ymin, 0, None) * np.clip(xmax - xmin, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-12)


def _encode(boxes, anc):
    """detector/utils/box_utils.py:79-109 (float64 here; only produces plausible codes)."""
    eps = 1e-8
    ha = anc[:, 2] - anc[:, 0] + eps; wa = anc[:, 3] - anc[:, 1] + eps
    cya = anc[:, 0] + 0.5 * (anc[:, 2] - anc[:, 0]); cxa = anc[:, 1] + 0.5 * (anc[:, 3] - anc[:, 1])
    h = boxes[:, 2] - boxes[:, 0] + eps; w = boxes[:, 3] - boxes[:, 1] + eps
    cy = boxes[:, 0] + 0.5 * (boxes[:, 2] - boxes[:, 0]); cx = boxes[:, 1] + 0.5 * (boxes[:, 3] - boxes[:, 1])
    return np.stack([10.0 * (cy - cya) / ha, 10.0 * (cx - cxa) / wa, 5.0 * np.log(h / ha), 5.0 * np.log(w / wa)], 1)


def _plant_boxes(rng, n):
    h = rng.uniform(0.15, 0.6, n); w = h * rng.uniform(0.3, 0.6, n)
    cy = rng.uniform(0.0, 1.0, n); cx = rng.uniform(0.0, 1.0, n)
    b = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], 1)
    return np.clip(b, 0.0, 1.0)


def make_inputs(wl: Workload, replicate=0, batch=None, seed=None):
    """-> dict(class_logits [B,A] f32, encoded_boxes [B,A,4] f32, heatmap_logits [B,h,w,18] f32, gt_boxes list)."""
    B = wl.batch if batch is None else int(batch)
    H, W = wl.height, wl.width
    rng = np.random.Generator(np.random.PCG64(seed_for(wl.config_id, replicate) if seed is None else seed))
    anc = anchors_np(H, W, wl.strides, wl.scales, wl.multipliers, wl.ratios).astype(np.float64)
    A = anc.shape[0]
    h4, w4 = H // 4, W // 4
    thr = float(wl.score_threshold)
    logit_thr = np.log(thr / (1.0 - thr))

    cls = rng.normal(PRIOR_LOGIT, 0.5, (B, A)).astype(np.float32)
    enc = rng.normal(0.0, 0.5, (B, A, 4)).astype(np.float32)
    hml = rng.normal(PRIOR_LOGIT, 0.3, (B, h4, w4, 18)).astype(np.float32)
    hml[..., 17] = rng.normal(0.0, 1.0, (B, h4, w4)).astype(np.float32)
    gts = []
    for b in range(B):
        P = int(rng.integers(wl.persons[0], wl.persons[1] + 1))
        gt = _plant_boxes(rng, P)
        gts.append(gt.astype(np.float32))
        iou = _iou_matrix(anc, gt)                              # [A,P]
        best = iou.argmax(1)
        matched = np.nonzero(iou.max(1) >= 0.5)[0]
        cls[b, matched] = rng.normal(1.5, 0.7, matched.size).astype(np.float32)
        codes = _encode(gt[best[matched]], anc[matched]) + rng.normal(0.0, 0.15, (matched.size, 4))
        enc[b, matched] = codes.astype(np.float32)
        # guard band around the score threshold and tie-free confident scores:
        # confident logits are spread at least 1e-4 apart, none within 1e-3 of the threshold logit
        row = cls[b].astype(np.float64)
        near = np.abs(row - logit_thr) < 1e-3
        row[near] = logit_thr + np.where(row[near] >= logit_thr, 2e-3, -2e-3)
        conf = np.nonzero(row > logit_thr - 0.05)[0]
        order = conf[np.argsort(row[conf], kind="stable")]
        vals = row[order]
        for i in range(1, vals.size):
            if vals[i] < vals[i - 1] + 1e-4:
                vals[i] = vals[i - 1] + 1e-4
        row[order] = vals
        near = np.abs(row - logit_thr) < 1e-3
        row[near] += 2.5e-3
        cls[b] = row.astype(np.float32)
        # keypoint bumps: 17 keypoints uniformly inside every planted box
        area_px = (gt[:, 2] - gt[:, 0]) * H * (gt[:, 3] - gt[:, 1]) * W
        sig = np.clip(0.007 * np.sqrt(area_px), 1.0, 4.0)       # heatmap_creation.py:22-37
        for p in range(P):
            ky = rng.uniform(gt[p, 0], gt[p, 2], 17) * (h4 - 1)
            kx = rng.uniform(gt[p, 1], gt[p, 3], 17) * (w4 - 1)
            rad = int(np.ceil(3.0 * sig[p]))
            for c in range(17):
                y0, x0 = int(round(ky[c])), int(round(kx[c]))
                ya, yb = max(y0 - rad, 0), min(y0 + rad + 1, h4)
                xa, xb = max(x0 - rad, 0), min(x0 + rad + 1, w4)
                yy, xx = np.mgrid[ya:yb, xa:xb]
                g = np.exp(-0.5 * ((yy - ky[c]) ** 2 + (xx - kx[c]) ** 2) / sig[p] ** 2)
                bump = (PRIOR_LOGIT + (3.0 - PRIOR_LOGIT) * g).astype(np.float32)
                hml[b, ya:yb, xa:xb, c] = np.maximum(hml[b, ya:yb, xa:xb, c], bump)
    enc = np.clip(enc, -4.0, 4.0).astype(np.float32)
    return {"class_logits": cls, "encoded_boxes": enc, "heatmap_logits": hml, "gt_boxes": gts,
            "image_hw": (H, W)}


def truncated_normal(rng, shape, std):
    out = rng.standard_normal(shape, dtype=np.float32)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()), dtype=np.float32)
        bad = np.abs(out) > 2.0
    out *= np.float32(std)
    return out


def make_prn_weights(seed=20240077, d=D_PRN, hidden=HIDDEN, bias_std=0.0):
    """tf.variance_scaling_initializer() defaults (prn.py:19): scale 1, fan_in, truncated normal,
    stddev = sqrt(1/fan_in)/0.87962566; biases zero (slim default).  `bias_std` > 0 perturbs the
    biases so that parity tests also exercise them."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W1 = truncated_normal(rng, (d, hidden), np.sqrt(1.0 / d) / 0.87962566103423978)
    W2 = truncated_normal(rng, (hidden, d), np.sqrt(1.0 / hidden) / 0.87962566103423978)
    b1 = (rng.standard_normal(hidden) * bias_std).astype(np.float32)
    b2 = (rng.standard_normal(d) * bias_std).astype(np.float32)
    return W1, b1, W2, b2


def make_crops(n, seed=20240500, density=0.03, d=D_PRN):
    """Config 5 (PRN-only sweep) inputs: sparse U(0,1) crops [n, 56, 36, 17]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.random((n, d), dtype=np.float32)
    x *= (rng.random((n, d), dtype=np.float32) < density)
    return x.reshape(n, 56, 36, 17)
