"""Readable numpy restatement of the hot path -- TEST INFRASTRUCTURE ONLY.

A second, independently written statement of SURVEY.md Appendix B, used by
tests/ to cross-check the C oracle (oracle/mpn_oracle.c) on random inputs.  All
arithmetic is numpy float32 (IEEE, one rounding per operation); the only shared
primitive is the exp recipe of oracle/exact_math.h (numpy has no fmaf).

Reference lines are cited per function (paths relative to the reference).
"""
import itertools
import math

import numpy as np

from . import expf as _expf

f32 = np.float32


def sigmoid(x):
    """tf.sigmoid (retinanet.py:73, create_pb.py:74): 1 / (1 + exp(-x))."""
    x = np.asarray(x, f32)
    return (f32(1) / (f32(1) + _expf(-x).reshape(x.shape))).astype(f32)


def anchors(H, W, strides, scales, multipliers, ratios):
    """anchor_generator.py:53-114 and tile_anchors :140-165."""
    fH, fW = f32(H), f32(W)
    pairs = list(itertools.product(multipliers, ratios))                      # :70
    ratio_sqrts = np.sqrt(np.array([r for _, r in pairs], f32))               # :141
    out = []
    for stride, S in zip(strides, scales):
        s = f32(stride)
        gh = int(np.ceil(fH / s)); gw = int(np.ceil(fW / s))                  # :59-60
        sc = np.array([m * S for m, _ in pairs], f32)                         # :75
        heights = sc / ratio_sqrts; widths = sc * ratio_sqrts                 # :142-143
        oy = f32(0.5) * (fH - (f32(gh) - f32(1)) * s)                         # :92
        ox = f32(0.5) * (fW - (f32(gw) - f32(1)) * s)                         # :93
        yc = np.arange(gh).astype(f32) * s + oy                               # :148
        xc = np.arange(gw).astype(f32) * s + ox                               # :149
        xg, yg = np.meshgrid(xc, yc)                                          # :150
        centers = np.stack([yg, xg], 2)[:, :, None, :]                        # :153-155
        sizes = np.stack([heights, widths], 1)[None, None]                    # :158-160
        half = f32(0.5) * sizes
        boxes = np.concatenate([centers - half, centers + half], 3)           # :163
        out.append(boxes.reshape(-1, 4).astype(f32))                          # :165
    a = np.concatenate(out, 0)                                                # :107
    return (a / np.array([fH, fW, fH, fW], f32)).astype(f32)                  # :110-114


def decode(codes, anc, sf=(10.0, 10.0, 5.0, 5.0)):
    """box_utils.py:112-139 (+ to_center_coordinates :63-76) and the clip of nms.py:36."""
    codes = np.asarray(codes, f32); anc = np.asarray(anc, f32)
    ha = anc[:, 2] - anc[:, 0]; wa = anc[:, 3] - anc[:, 1]
    cya = anc[:, 0] + f32(0.5) * ha; cxa = anc[:, 1] + f32(0.5) * wa
    ty = codes[:, 0] / f32(sf[0]); tx = codes[:, 1] / f32(sf[1])
    th = codes[:, 2] / f32(sf[2]); tw = codes[:, 3] / f32(sf[3])
    h = _expf(th) * ha; w = _expf(tw) * wa
    cy = ty * ha + cya; cx = tx * wa + cxa
    b = np.stack([cy - f32(0.5) * h, cx - f32(0.5) * w, cy + f32(0.5) * h, cx + f32(0.5) * w], 1)
    return np.clip(b, f32(0), f32(1)).astype(f32)


def iou(p, q):
    """TF 1.15 NonMaxSuppressionV3 overlap (third-party, restated)."""
    p = np.asarray(p, f32); q = np.asarray(q, f32)
    y0i, x0i, y1i, x1i = min(p[0], p[2]), min(p[1], p[3]), max(p[0], p[2]), max(p[1], p[3])
    y0j, x0j, y1j, x1j = min(q[0], q[2]), min(q[1], q[3]), max(q[0], q[2]), max(q[1], q[3])
    ai = f32(y1i - y0i) * f32(x1i - x0i); aj = f32(y1j - y0j) * f32(x1j - x0j)
    if ai <= 0 or aj <= 0:
        return f32(0)
    ih = max(f32(min(y1i, y1j) - max(y0i, y0j)), f32(0)); iw = max(f32(min(x1i, x1j) - max(x0i, x0j)), f32(0))
    inter = f32(ih * iw)
    return f32(inter / f32(f32(ai + aj) - inter))


def detect_image(cls, enc, anc, thr, iou_thr, max_det, sf=(10.0, 10.0, 5.0, 5.0)):
    """retinanet.py:73 + nms.py:27-53 for one image -> (boxes [max_det,4], scores [max_det], n, sel_anchor)."""
    scores = sigmoid(cls)
    conf = np.nonzero(scores >= f32(thr))[0]                                  # nms.py:30-33 (ascending index)
    boxes = decode(np.asarray(enc, f32)[conf], np.asarray(anc, f32)[conf], sf)  # nms.py:35-36
    sc = scores[conf]
    order = [i for i in range(len(conf)) if sc[i] > f32(thr)]                 # TF NMS: strict >
    order.sort(key=lambda i: (-float(sc[i]), int(conf[i])))                   # score desc, index asc
    sel = []
    for i in order:
        if len(sel) >= max_det:
            break
        if all(not (iou(boxes[i], boxes[j]) > f32(iou_thr)) for j in sel):
            sel.append(i)
    ob = np.zeros((max_det, 4), f32); os_ = np.zeros((max_det,), f32); oa = -np.ones((max_det,), np.int32)
    for k, i in enumerate(sel):                                               # nms.py:43-52
        ob[k] = boxes[i]; os_[k] = sc[i]; oa[k] = conf[i]
    return ob, os_, len(sel), oa


def normalise(kh):
    """create_pb.py:90-94 on [B,h,w,c]."""
    kh = np.asarray(kh, f32)
    M = kh.max(axis=(1, 2), keepdims=True); m = kh.min(axis=(1, 2), keepdims=True)
    mask = (M > f32(0.2)).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (((kh - m) / (M - m)) * mask).astype(f32)


def crop_and_resize(img, box, crop_h, crop_w):
    """TF 1.15 CropAndResize, bilinear, extrapolation 0 (create_pb.py:106-109).  img [h,w,c] -> [crop_h,crop_w,c]."""
    img = np.asarray(img, f32)
    h, w, c = img.shape
    y1, x1, y2, x2 = [f32(v) for v in box]
    hs = f32(f32(f32(y2 - y1) * f32(h - 1)) / f32(crop_h - 1)) if crop_h > 1 else f32(0)
    ws = f32(f32(f32(x2 - x1) * f32(w - 1)) / f32(crop_w - 1)) if crop_w > 1 else f32(0)
    out = np.zeros((crop_h, crop_w, c), f32)
    for cy in range(crop_h):
        iy = f32(f32(y1 * f32(h - 1)) + f32(f32(cy) * hs)) if crop_h > 1 else f32(f32(f32(0.5) * f32(y1 + y2)) * f32(h - 1))
        if iy < 0 or iy > h - 1:
            continue
        t = int(math.floor(iy)); b = int(math.ceil(iy)); ly = f32(iy - f32(t))
        for cx in range(crop_w):
            ix = f32(f32(x1 * f32(w - 1)) + f32(f32(cx) * ws)) if crop_w > 1 else f32(f32(f32(0.5) * f32(x1 + x2)) * f32(w - 1))
            if ix < 0 or ix > w - 1:
                continue
            l = int(math.floor(ix)); r = int(math.ceil(ix)); lx = f32(ix - f32(l))
            top = img[t, l] + (img[t, r] - img[t, l]) * lx
            bot = img[b, l] + (img[b, r] - img[b, l]) * lx
            out[cy, cx] = top + (bot - top) * ly
    return out


def prn(x, W1, b1, W2, b2):
    """prn.py:15-25 in float64 (rounded to fp32 at the layer outputs)."""
    x2 = np.asarray(x, f32).reshape(x.shape[0], -1)
    y1 = np.maximum((x2.astype(np.float64) @ W1.astype(np.float64) + b1).astype(f32), f32(0))
    y2 = np.maximum((y1.astype(np.float64) @ W2.astype(np.float64) + b2).astype(f32), f32(0))
    return (x2 + y2).reshape(x.shape)


def keypoint_decode(logits, crop_h, crop_w):
    """create_pb.py:115-142 on [N,crop_h,crop_w,c] -> scores [N,c], positions [N,c,2], argmax [N,c]."""
    N, _, _, c = logits.shape
    L = np.asarray(logits, f32).reshape(N, crop_h * crop_w, c)
    lmax = L.max(axis=1, keepdims=True)
    E = _expf((L - lmax).astype(f32)).reshape(L.shape)
    S = np.zeros((N, c), f32)
    for p in range(L.shape[1]):                   # sequential fp32 sum, the oracle's canonical order
        S = (S + E[:, p, :]).astype(f32)
    P = (E * (f32(1) / S)[:, None, :]).astype(f32)
    arg = P.argmax(axis=1).astype(np.int32)       # first index on ties, as tf.argmax
    scores = P.max(axis=1)
    pos = np.stack([(arg // crop_w).astype(f32) / f32(crop_h), (arg % crop_w).astype(f32) / f32(crop_w)], 2)
    return scores.astype(f32), pos.astype(f32), arg


def get_keypoints(heatmaps, box, threshold):
    """inference/utils.py:29-52, vectorised over the channels: rows (x, y, visible)."""
    hm = np.asarray(heatmaps)
    h, w, nk = hm.shape
    flat = hm.reshape(h * w, nk)
    peak = flat.argmax(axis=0)                       # first maximum per channel (:47)
    visible = flat.max(axis=0) > threshold           # strict (:46)
    box_h, box_w = box[2] - box[0], box[3] - box[1]  # :40-41
    out = np.zeros((nk, 3), np.int32)
    for j in np.nonzero(visible)[0]:
        py, px = divmod(int(peak[j]), w)
        yy = min(max(int(py * box_h / h), 0), box_h)     # int() truncates, then clip (:48)
        xx = min(max(int(px * box_w / w), 0), box_w)     # :49
        out[j] = (int(xx), int(yy), 1)                   # :50 (x first)
    return out
