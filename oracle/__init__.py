"""CPU oracle of the MultiPoseNet post-backbone path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers over oracle/libmpn_oracle.so (oracle/mpn_oracle.c).  Only
tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this package; the product package
(multiposenet_b200) never does.

PINNING: see the header of mpn_oracle.c (reference-executed goldens pin everything but the last ulp of the TensorFlow
library kernels, which stays "parity unpinned").
"""
import ctypes as C
import itertools

import numpy as np

from .build import ensure_oracle

_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def _fp(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_i32p) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(ensure_oracle())
        _lib.orc_iou.restype = C.c_float
        _lib.orc_num_anchors.restype = C.c_int
        _lib.orc_nms.restype = C.c_int
        _lib.orc_detector_filter.restype = C.c_int
        _lib.orc_version.restype = C.c_int
        _lib.orc_sigmoid_monotone_violations.restype = C.c_int64
    return _lib


# defaults of the reference (detector/retinanet.py:38-43, detector/constants.py:19, create_pb.py:19)
STRIDES = (8, 16, 32, 64, 128)
SCALES = (32, 64, 128, 256, 512)
MULTIPLIERS = (1.0, 1.4142)
RATIOS = (1.0, 2.0, 0.5)
SCALE_FACTORS = (10.0, 10.0, 5.0, 5.0)
CROP_SIZE = (56, 36)
NUM_KEYPOINTS = 17


def expf(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().orc_expf_array(_fp(x), _fp(y), C.c_int64(x.size))
    return y


def sigmoidf(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().orc_sigmoidf_array(_fp(x), _fp(y), C.c_int64(x.size))
    return y


def sigmoid_monotone_violations(key_begin=0, count=1 << 32):
    """Neighbouring floats x < x' with sigmoid(x) > sigmoid(x') among `count` pairs in increasing order (0 = monotone)."""
    return int(lib().orc_sigmoid_monotone_violations(C.c_uint32(key_begin), C.c_uint64(count)))


def round_bf16(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().orc_round_bf16_array(_fp(x), _fp(y), C.c_int64(x.size))
    return y


def num_anchors(H, W, strides=STRIDES, n_loc=6):
    s = np.asarray(strides, dtype=np.int32)
    return int(lib().orc_num_anchors(int(H), int(W), len(s), _ip(s), int(n_loc)))


def anchors(H, W, strides=STRIDES, scales=SCALES, multipliers=MULTIPLIERS, ratios=RATIOS):
    """detector/anchor_generator.py:40-116 -> float32 [A,4] normalised (ymin,xmin,ymax,xmax)."""
    s = np.asarray(strides, dtype=np.int32)
    sc = np.asarray(scales, dtype=np.float64)
    m = np.asarray(multipliers, dtype=np.float64)
    r = np.asarray(ratios, dtype=np.float64)
    A = num_anchors(H, W, strides, len(m) * len(r))
    out = np.empty((A, 4), dtype=np.float32)
    lib().orc_anchors(int(H), int(W), len(s), _ip(s), _dp(sc), len(m), _dp(m), len(r), _dp(r), _fp(out))
    return out


def decode(codes, anchors_, scale_factors=SCALE_FACTORS):
    """box_utils.py:112-139 followed by the clip of nms.py:36."""
    codes, anchors_ = _f32(codes), _f32(anchors_)
    sf = np.asarray(scale_factors, dtype=np.float32)
    out = np.empty_like(codes)
    lib().orc_decode(_fp(codes), _fp(anchors_), C.c_int64(codes.shape[0]), _fp(sf), _fp(out))
    return out


def iou(p, q):
    p, q = _f32(p), _f32(q)
    return float(lib().orc_iou(_fp(p), _fp(q)))


def nms(boxes, scores, score_thr, iou_thr, max_out):
    """The TF NonMaxSuppressionV3 op alone (nms.py:38-41) -> selected indices (int32)."""
    boxes, scores = _f32(boxes), _f32(scores)
    sel = np.empty(max(int(max_out), 1), dtype=np.int32)
    k = lib().orc_nms(_fp(boxes), _fp(scores), int(scores.shape[0]), C.c_float(score_thr), C.c_float(iou_thr),
                      int(max_out), _ip(sel))
    return sel[:k].copy()


def detect(cls, enc, anchors_, thr, iou_thr, max_det, scale_factors=SCALE_FACTORS):
    """retinanet.py:73 + nms.py:6-61 for a batch.  cls [B,A], enc [B,A,4] ->
    dict(boxes [B,max_det,4], scores [B,max_det], num_boxes [B], sel_anchor [B,max_det], n_conf [B])."""
    cls, enc, anchors_ = _f32(cls), _f32(enc), _f32(anchors_)
    B, A = cls.shape
    sf = np.asarray(scale_factors, dtype=np.float32)
    boxes = np.zeros((B, max_det, 4), np.float32)
    scores = np.zeros((B, max_det), np.float32)
    num = np.zeros((B,), np.int32)
    sel = np.zeros((B, max_det), np.int32)
    nconf = np.zeros((B,), np.int32)
    lib().orc_detect_batch(_fp(cls), _fp(enc), _fp(anchors_), int(B), int(A), _fp(sf), C.c_float(thr),
                           C.c_float(iou_thr), int(max_det), _fp(boxes), _fp(scores), _ip(num), _ip(sel), _ip(nconf))
    return {"boxes": boxes, "scores": scores, "num_boxes": num, "sel_anchor": sel, "n_conf": nconf}


def detect_image_margin(cls, enc, anchors_, thr, iou_thr, max_det, scale_factors=SCALE_FACTORS):
    """Single image; additionally returns min |IoU - iou_thr| over evaluated pairs (guard-band report)."""
    cls, enc, anchors_ = _f32(cls), _f32(enc), _f32(anchors_)
    A = cls.shape[0]
    sf = np.asarray(scale_factors, dtype=np.float32)
    boxes = np.zeros((max_det, 4), np.float32)
    scores = np.zeros((max_det,), np.float32)
    num = C.c_int32(0)
    sel = np.zeros((max_det,), np.int32)
    nconf = C.c_int32(0)
    margin = C.c_float(0)
    lib().orc_detect_image(_fp(cls), _fp(enc), _fp(anchors_), int(A), _fp(sf), C.c_float(thr), C.c_float(iou_thr),
                           int(max_det), _fp(boxes), _fp(scores), C.byref(num), _ip(sel), C.byref(nconf),
                           C.byref(margin))
    return boxes, scores, num.value, sel, nconf.value, margin.value


def heatmaps(hml, nk=NUM_KEYPOINTS):
    """create_pb.py:73-76, 90, 92.  hml [B,h,w,nk+1] -> kh [B,h,w,nk], seg [B,h,w], mn [B,nk], mx [B,nk]."""
    hml = _f32(hml)
    B, h, w, ch = hml.shape
    kh = np.empty((B, h, w, nk), np.float32)
    seg = np.empty((B, h, w), np.float32)
    mn = np.empty((B, nk), np.float32)
    mx = np.empty((B, nk), np.float32)
    lib().orc_heatmaps_batch(_fp(hml), int(B), int(h), int(w), int(ch), int(nk), _fp(kh), _fp(seg), _fp(mn), _fp(mx))
    return kh, seg, mn, mx


def heatmap_head(features, weight, bias):
    """SURVEY section 8(f) row 2 -- detector/keypoint_subnet.py:49-58: tf.layers.conv2d(x, 18, kernel_size=1) with bias on
    the NCHW feature map, then the NCHW -> NHWC transpose (:56).  features [B,64,h,w], weight [64,18] (the [1,1,64,18]
    HWIO kernel), bias [18] -> heatmap logits [B,h,w,18] float32 (accumulated in float64: the order of TF's own
    convolution kernel is unspecified, parity is held to 1e-4)."""
    x = np.asarray(features, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    out = np.einsum("bchw,ck->bhwk", x, w) + np.asarray(bias, dtype=np.float64)
    return out.astype(np.float32)


def crop_and_resize(img, boxes, box_ind, crop_size=CROP_SIZE, mn=None, mx=None):
    """tf.image.crop_and_resize (create_pb.py:106-109); with mn/mx the taps are min-max
    normalised first (create_pb.py:90-94).  img [B,h,w,c] -> [N,ch,cw,c]."""
    img, boxes = _f32(img), _f32(boxes).reshape(-1, 4)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    B, h, w, c = img.shape
    N = boxes.shape[0]
    if mn is not None:
        mn, mx = _f32(mn), _f32(mx)
    out = np.empty((N, crop_size[0], crop_size[1], c), np.float32)
    lib().orc_crop_batch(_fp(img), int(B), int(h), int(w), int(c), _fp(mn), _fp(mx), _fp(boxes), _ip(box_ind), int(N),
                         int(crop_size[0]), int(crop_size[1]), _fp(out))
    return out


def prn(x, W1, b1, W2, b2, mode=0, return_hidden=False):
    """detector/prn.py:15-25.  x [N,h,w,c] or [N,D].  mode 0: fp32 operands / fp64 accumulate;
    mode 1: bf16-rounded operands (x, W1, y1, W2) / fp64 accumulate."""
    shape = x.shape
    x = _f32(x).reshape(shape[0], -1)
    N, D = x.shape
    W1, b1, W2, b2 = _f32(W1), _f32(b1), _f32(W2), _f32(b2)
    hidden = W1.shape[1]
    assert W1.shape == (D, hidden) and W2.shape == (hidden, D)
    out = np.empty((N, D), np.float32)
    y1 = np.empty((N, hidden), np.float32) if return_hidden else None
    lib().orc_prn(_fp(x), int(N), int(D), int(hidden), _fp(W1), _fp(b1), _fp(W2), _fp(b2), int(mode), _fp(out),
                  _fp(y1))
    out = out.reshape(shape)
    return (out, y1) if return_hidden else out


def keypoint_decode(logits, crop_size=CROP_SIZE, nk=NUM_KEYPOINTS):
    """create_pb.py:115-142.  logits [N,ch,cw,nk] -> scores [N,nk], positions [N,nk,2], argmax [N,nk], gap [N,nk]."""
    logits = _f32(logits)
    N = logits.shape[0]
    scores = np.empty((N, nk), np.float32)
    pos = np.empty((N, nk, 2), np.float32)
    arg = np.empty((N, nk), np.int32)
    gap = np.empty((N, nk), np.float32)
    lib().orc_keypoint_decode(_fp(logits), int(N), int(crop_size[0]), int(crop_size[1]), int(nk), _fp(scores), _fp(pos),
                              _ip(arg), _fp(gap))
    return scores, pos, arg, gap


def get_keypoints(heatmaps_, box, threshold):
    """inference/utils.py:29-52 -> int32 [17,3] rows (x, y, visible)."""
    hm = _f32(heatmaps_)
    h, w, nk = hm.shape
    b = np.asarray(box, dtype=np.float64)
    out = np.zeros((nk, 3), np.int32)
    lib().orc_get_keypoints(_fp(hm), int(h), int(w), int(nk), _dp(b), C.c_double(threshold), _ip(out))
    return out


def full_path(cls, enc, hml, H, W, W1, b1, W2, b2, thr=0.3, iou_thr=0.6, max_det=25, prn_mode=0,
              strides=STRIDES, scales=SCALES, multipliers=MULTIPLIERS, ratios=RATIOS,
              scale_factors=SCALE_FACTORS, crop_size=CROP_SIZE, prn_fn=None):
    """create_pb.py:73-152 downstream of the networks: the seven outputs (+ diagnostics).
    `prn_fn(x2d) -> logits2d` may replace the fp64-accumulating PRN (bench cpu_baseline uses BLAS)."""
    anc = anchors(H, W, strides, scales, multipliers, ratios)
    det = detect(cls, enc, anc, thr, iou_thr, max_det, scale_factors)
    kh, seg, mn, mx = heatmaps(hml)
    B = cls.shape[0]
    pb, pi = [], []
    for b in range(B):
        n = int(det["num_boxes"][b])
        pb.append(det["boxes"][b, :n])
        pi.append(np.full((n,), b, np.int32))
    pboxes = np.concatenate(pb, 0) if pb else np.zeros((0, 4), np.float32)
    pind = np.concatenate(pi, 0) if pi else np.zeros((0,), np.int32)
    crops = crop_and_resize(kh, pboxes, pind, crop_size, mn, mx)
    N = crops.shape[0]
    if prn_fn is not None:
        logits = prn_fn(crops.reshape(N, -1)).reshape(crops.shape).astype(np.float32)
    else:
        logits = prn(crops, W1, b1, W2, b2, prn_mode)
    ks, kp, arg, gap = keypoint_decode(logits, crop_size)
    out = dict(det)
    out.update({"keypoint_heatmaps": kh, "segmentation_masks": seg, "keypoint_scores": ks, "keypoint_positions": kp,
                "keypoint_argmax": arg, "keypoint_gap": gap, "crops": crops, "logits": logits, "person_boxes": pboxes,
                "person_image": pind, "hm_min": mn, "hm_max": mx, "anchors": anc})
    return out


def anchor_pairs(multipliers=MULTIPLIERS, ratios=RATIOS):
    return list(itertools.product(multipliers, ratios))
