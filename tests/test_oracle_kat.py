"""Pins the CPU oracle: hand-derived known answers of SURVEY.md section 8(c), the golden vectors produced by the
reference's own get_keypoints, and cross-checks against independent restatements (numpy spec, torchvision NMS,
torch grid_sample).  No GPU needed."""
import math
import os

import numpy as np
import pytest

import oracle
from oracle import spec_np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


# ---------------------------------------------------------------------------- exp / sigmoid recipe
def test_exp_recipe_accuracy_and_edges():
    x = np.concatenate([np.linspace(-87, 88, 400001), np.linspace(-1, 1, 200001), [0.0, -0.0]]).astype(f32)
    e = oracle.expf(x)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(e / ref - 1.0)) < 2e-7
    assert oracle.expf(f32([0.0]))[0] == 1.0
    assert oracle.expf(f32([-87.5]))[0] == 0.0 and oracle.expf(f32([-1e30]))[0] == 0.0
    assert np.isinf(oracle.expf(f32([88.5]))[0]) and np.isnan(oracle.expf(f32([np.nan]))[0])
    s = oracle.sigmoidf(f32([-200, 0, 200]))
    assert s[0] == 0.0 and s[1] == 0.5 and s[2] == 1.0
    xs = np.linspace(-20, 20, 100001).astype(f32)
    assert np.max(np.abs(oracle.sigmoidf(xs) - 1 / (1 + np.exp(-xs.astype(np.float64)))) /
                  (1 / (1 + np.exp(-xs.astype(np.float64))))) < 3e-7


def test_sigmoid_recipe_is_monotone_over_all_floats():
    """Exhaustive: every pair of neighbouring finite floats.  The product takes min / max of the heatmap LOGITS and applies
    the sigmoid to the two extremes (create_pb.py:90,92 ask for min / max of the activations): identical bits iff the
    recipe never decreases."""
    assert oracle.sigmoid_monotone_violations(0, 1 << 32) == 0
    e = oracle.sigmoidf(f32([-104.0, -103.0, -88.0, -87.5, -87.0, 0.0, 16.0, 17.0, 88.0, 200.0]))
    assert (np.diff(e) >= 0).all()


def test_bf16_rounding():
    x = f32([1.0, 1.00390625, 1.005859375, -2.5, 3.0e38, 1e-40, np.inf])
    y = oracle.round_bf16(x)
    import torch
    ref = torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    r = np.random.default_rng(0).standard_normal(100000).astype(f32)
    assert np.array_equal(oracle.round_bf16(r), torch.tensor(r).to(torch.bfloat16).to(torch.float32).numpy())


# ---------------------------------------------------------------------------- (1)(2) anchors
def test_anchor_counts():
    assert oracle.num_anchors(512, 512) == 32736
    assert oracle.num_anchors(640, 640) == 51150
    assert oracle.num_anchors(1024, 1024) == 130944
    assert oracle.num_anchors(640, 640, n_loc=9) == 76725
    assert [math.ceil(640 / s) ** 2 for s in oracle.STRIDES] == [6400, 1600, 400, 100, 25]


def test_anchor_values_640():
    a = oracle.anchors(640, 640)
    want = f32([[-0.01875, -0.01875, 0.03125, 0.03125],
                [-0.011427669, -0.02910534, 0.02392767, 0.04160534],
                [-0.02910534, -0.011427669, 0.04160534, 0.02392767],
                [-0.029104998, -0.029104998, 0.041605, 0.041605],
                [-0.01874976, -0.04374952, 0.031249762, 0.056249518],
                [-0.04374952, -0.018749759, 0.05624952, 0.031249758],
                [-0.01875, -0.00625, 0.03125, 0.04375]])
    np.testing.assert_allclose(a[:7], want, rtol=2e-7, atol=0)
    np.testing.assert_allclose(a[-1], f32([0.100007676, 0.5000039, 1.6999924, 1.2999961]), rtol=2e-7)


@pytest.mark.parametrize("hw", [(128, 128), (256, 384), (640, 640)])
@pytest.mark.parametrize("mults", [(1.0, 1.4142), (1.0, 2 ** (1 / 3), 2 ** (2 / 3))])
def test_anchors_c_equals_numpy_spec(hw, mults):
    a = oracle.anchors(hw[0], hw[1], multipliers=mults)
    b = spec_np.anchors(hw[0], hw[1], oracle.STRIDES, oracle.SCALES, mults, oracle.RATIOS)
    assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---------------------------------------------------------------------------- (3) decode
def test_decode_identity_and_roundtrip():
    anc = oracle.anchors(256, 256)
    zero = np.zeros_like(anc)
    np.testing.assert_allclose(oracle.decode(zero, anc), np.clip(anc, 0, 1), atol=1e-7)
    rng = np.random.default_rng(1)
    n = 2000
    idx = rng.integers(0, anc.shape[0], n)
    a = anc[idx].astype(np.float64)
    cy, cx = rng.uniform(0.3, 0.7, n), rng.uniform(0.3, 0.7, n)
    h, w = rng.uniform(0.05, 0.3, n), rng.uniform(0.05, 0.3, n)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
    ha, wa = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    cya, cxa = a[:, 0] + ha / 2, a[:, 1] + wa / 2
    codes = np.stack([10 * (cy - cya) / ha, 10 * (cx - cxa) / wa, 5 * np.log(h / ha), 5 * np.log(w / wa)], 1)
    out = oracle.decode(codes.astype(f32), anc[idx])
    np.testing.assert_allclose(out, b, atol=3e-6)


def test_decode_c_equals_numpy_spec():
    rng = np.random.default_rng(2)
    anc = oracle.anchors(256, 256)
    codes = rng.normal(0, 1.5, anc.shape).astype(f32)
    a = oracle.decode(codes, anc)
    b = spec_np.decode(codes, anc)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---------------------------------------------------------------------------- (4) NMS
def test_nms_known_answers():
    box = f32([[0.1, 0.1, 0.5, 0.5], [0.1, 0.1, 0.5, 0.5]])
    assert list(oracle.nms(box, f32([0.9, 0.8]), 0.5, 0.5, 10)) == [0]
    # IoU exactly 0.5: a = [0,0,1,1] (area 1), b = [0,0,0.5,1] (area .5): inter .5, union 1 -> 0.5; strict > keeps both
    box = f32([[0, 0, 1, 1], [0, 0, 0.5, 1]])
    assert oracle.iou(box[0], box[1]) == 0.5
    assert list(oracle.nms(box, f32([0.9, 0.8]), 0.3, 0.5, 10)) == [0, 1]
    assert list(oracle.nms(box, f32([0.9, 0.8]), 0.3, 0.49999, 10)) == [0]
    # score exactly == threshold is never selected (strict >)
    assert list(oracle.nms(box, f32([0.5, 0.6]), 0.5, 0.9, 10)) == [1]
    # zero-area box: never suppressed, never suppresses
    box = f32([[0.2, 0.2, 0.2, 0.6], [0.2, 0.2, 0.2, 0.6], [0.1, 0.1, 0.7, 0.7]])
    assert list(oracle.nms(box, f32([0.9, 0.8, 0.7]), 0.1, 0.01, 10)) == [0, 1, 2]
    # flipped corners are re-ordered
    assert oracle.iou(f32([0.5, 0.5, 0.1, 0.1]), f32([0.1, 0.1, 0.5, 0.5])) == 1.0
    # max_output_size
    box = f32([[0, 0, .1, .1], [.2, .2, .3, .3], [.4, .4, .5, .5]])
    assert list(oracle.nms(box, f32([0.5, 0.9, 0.7]), 0.1, 0.5, 2)) == [1, 2]
    # ties: lower index first (matched tie-break)
    assert list(oracle.nms(box, f32([0.7, 0.7, 0.7]), 0.1, 0.5, 3)) == [0, 1, 2]


def test_nms_matches_torchvision_on_tie_free_inputs():
    torchvision = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = int(rng.integers(5, 400))
        cy, cx = rng.uniform(0.2, 0.8, n), rng.uniform(0.2, 0.8, n)
        h, w = rng.uniform(0.05, 0.4, n), rng.uniform(0.05, 0.4, n)
        boxes = np.clip(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1), 0, 1).astype(f32)
        scores = rng.permutation(n).astype(f32) / f32(n) * f32(0.9) + f32(0.05)
        thr, iou_thr = 0.3, float(rng.choice([0.3, 0.5, 0.6]))
        sel = oracle.nms(boxes, scores, thr, iou_thr, n)
        keep = np.nonzero(scores > thr)[0]
        xyxy = torch.tensor(boxes[keep][:, [1, 0, 3, 2]])
        tv = torchvision.ops.nms(xyxy, torch.tensor(scores[keep]), iou_thr).numpy()
        assert list(sel) == list(keep[tv])


def test_detect_c_equals_numpy_spec():
    from multiposenet_b200 import synthetic
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    anc = oracle.anchors(wl.height, wl.width, multipliers=wl.multipliers)
    det = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, 0.3, 0.6, 25)
    for b in range(wl.batch):
        ob, os_, n, oa = spec_np.detect_image(inp["class_logits"][b], inp["encoded_boxes"][b], anc, 0.3, 0.6, 25)
        assert n == det["num_boxes"][b] and n >= 1
        assert np.array_equal(oa, det["sel_anchor"][b])
        assert np.array_equal(ob.view(np.uint32), det["boxes"][b].view(np.uint32))
        assert np.array_equal(os_.view(np.uint32), det["scores"][b].view(np.uint32))
    # padding rows are zero (nms.py:47-52)
    n0 = det["num_boxes"][0]
    assert not det["boxes"][0, n0:].any() and not det["scores"][0, n0:].any()


# ---------------------------------------------------------------------------- heatmaps
def test_heatmaps_and_normalise():
    rng = np.random.default_rng(4)
    hml = rng.normal(-3, 1.5, (2, 32, 32, 18)).astype(f32)
    hml[1, :, :, 5] = -6.0          # weak channel: max sigmoid <= 0.2 -> masked to zero, and M == m -> NaN * 0 = NaN
    hml[1, 3, 4, 6] = 4.0
    kh, seg, mn, mx = oracle.heatmaps(hml)
    assert np.array_equal(seg, hml[..., 17])
    assert np.array_equal(kh, spec_np.sigmoid(hml[..., :17]))
    assert np.array_equal(mn, kh.min(axis=(1, 2))) and np.array_equal(mx, kh.max(axis=(1, 2)))
    nh = spec_np.normalise(kh)
    assert np.isnan(nh[1, :, :, 5]).all()          # 0/0 propagates exactly as the reference graph would
    box = f32([[0, 0, 1, 1]])
    full = oracle.crop_and_resize(kh, box, np.array([1], np.int32), (32, 32), mn, mx)[0]
    ok = ~np.isnan(nh[1])
    assert np.array_equal(full[ok], nh[1][ok])


# ---------------------------------------------------------------------------- (5) crop_and_resize
def test_crop_identity_and_decimation():
    rng = np.random.default_rng(5)
    img = rng.random((1, 56, 36, 17), dtype=f32)
    out = oracle.crop_and_resize(img, f32([[0, 0, 1, 1]]), np.zeros(1, np.int32))
    assert np.array_equal(out[0], img[0])
    img2 = rng.random((1, 2 * 55 + 1, 2 * 35 + 1, 17), dtype=f32)
    out2 = oracle.crop_and_resize(img2, f32([[0, 0, 1, 1]]), np.zeros(1, np.int32))
    assert np.array_equal(out2[0], img2[0, ::2, ::2])


def test_crop_extrapolation_and_spec():
    rng = np.random.default_rng(6)
    img = rng.random((2, 40, 24, 17), dtype=f32)
    boxes = f32([[-0.2, 0.1, 0.7, 1.3], [0.3, 0.2, 0.9, 0.8], [0.9, 0.9, 0.1, 0.1], [0.5, 0.5, 0.5, 0.5]])
    ind = np.array([0, 1, 1, 0], np.int32)
    out = oracle.crop_and_resize(img, boxes, ind, (14, 9))
    for n in range(4):
        ref = spec_np.crop_and_resize(img[ind[n]], boxes[n], 14, 9)
        assert np.array_equal(out[n].view(np.uint32), ref.view(np.uint32))
    assert (out[0][0] == 0).all()          # first rows sample above the image -> extrapolation value 0
    assert (out[0][:, -1] == 0).all()


def test_crop_matches_grid_sample_inside_image():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(7)
    img = rng.random((1, 40, 24, 17), dtype=f32)
    box = f32([0.1, 0.2, 0.8, 0.9])
    out = oracle.crop_and_resize(img, box[None], np.zeros(1, np.int32), (56, 36))[0]
    ys = box[0] * 39 + np.arange(56) * ((box[2] - box[0]) * 39 / 55)
    xs = box[1] * 23 + np.arange(36) * ((box[3] - box[1]) * 23 / 35)
    gy, gx = np.meshgrid(ys / 39 * 2 - 1, xs / 23 * 2 - 1, indexing="ij")
    grid = torch.tensor(np.stack([gx, gy], -1)[None], dtype=torch.float32)
    ref = F.grid_sample(torch.tensor(img).permute(0, 3, 1, 2), grid, mode="bilinear", align_corners=True)
    np.testing.assert_allclose(out, ref[0].permute(1, 2, 0).numpy(), atol=2e-5)


# ---------------------------------------------------------------------------- (6) PRN
def test_prn_known_answers():
    rng = np.random.default_rng(8)
    D, Hd, N = 17 * 6 * 4, 32, 5
    x = rng.random((N, 6, 4, 17), dtype=f32)
    W1 = rng.normal(0, 0.1, (D, Hd)).astype(f32); W2 = rng.normal(0, 0.1, (Hd, D)).astype(f32)
    b1 = rng.normal(0, 0.1, Hd).astype(f32); b2 = rng.normal(0, 0.1, D).astype(f32)
    z = oracle.prn(x, np.zeros_like(W1), np.zeros_like(b1), np.zeros_like(W2), np.zeros_like(b2))
    assert np.array_equal(z, x)                                         # zero weights -> identity (prn.py:24)
    z = oracle.prn(x, W1, b1, W2, np.full_like(b2, -1e6))
    assert np.array_equal(z, x)                                         # ReLU clamps -> identity
    out = oracle.prn(x, W1, b1, W2, b2)
    np.testing.assert_allclose(out, spec_np.prn(x, W1, b1, W2, b2), rtol=1e-6, atol=1e-6)
    out_b = oracle.prn(x, W1, b1, W2, b2, mode=1)
    np.testing.assert_allclose(out_b, out, rtol=3e-2, atol=3e-2)
    rb = oracle.round_bf16
    y1 = np.maximum(rb(x.reshape(N, -1)).astype(np.float64) @ rb(W1).astype(np.float64) + b1, 0).astype(f32)
    y2 = np.maximum(rb(y1).astype(np.float64) @ rb(W2).astype(np.float64) + b2, 0).astype(f32)
    np.testing.assert_allclose(out_b.reshape(N, -1), x.reshape(N, -1) + y2, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------- (7) keypoint decode
def test_keypoint_decode_known_answers():
    L = np.zeros((3, 56, 36, 17), f32)
    L[1] = 2.5
    L[2, 10, 7, :] = 30.0
    L[2, 20, 30, 3] = 30.0                      # exact tie in channel 3 -> first index wins
    L[2, 5, 5, 4] = 40.0
    s, p, a, gap = oracle.keypoint_decode(L)
    assert (a[0] == 0).all() and (a[1] == 0).all()
    np.testing.assert_allclose(s[0], 1.0 / 2016, rtol=1e-6)
    assert (p[0] == 0).all()
    assert a[2, 0] == 10 * 36 + 7 and a[2, 3] == 10 * 36 + 7 and a[2, 4] == 5 * 36 + 5
    assert p[2, 0, 0] == f32(10) / f32(56) and p[2, 0, 1] == f32(7) / f32(36)
    assert s[2, 0] > 0.999 and abs(s[2, 3] - 0.5) < 1e-6
    s2, p2, a2 = spec_np.keypoint_decode(L, 56, 36)
    assert np.array_equal(a, a2) and np.array_equal(p, p2) and np.array_equal(s, s2)
    # exp(d) rounds to 1.0f for |d| < 2^-25: the FIRST such position wins even though its logit is lower
    L = np.zeros((1, 56, 36, 17), f32)
    L[0, 30, 30, :] = f32(0.01)
    L[0, 2, 2, :] = np.nextafter(f32(0.01), f32(0))
    s, p, a, gap = oracle.keypoint_decode(L)
    assert (a[0] == 2 * 36 + 2).all()
    _, _, a2 = spec_np.keypoint_decode(L, 56, 36)
    assert np.array_equal(a, a2)


def test_keypoint_decode_random_vs_spec():
    rng = np.random.default_rng(9)
    L = rng.normal(0, 1, (4, 56, 36, 17)).astype(f32)
    s, p, a, gap = oracle.keypoint_decode(L)
    s2, p2, a2 = spec_np.keypoint_decode(L, 56, 36)
    assert np.array_equal(a, a2) and np.array_equal(p, p2) and np.array_equal(s, s2)
    assert (gap > 0).all()


def test_heatmap_head_known_answers():
    """keypoint_subnet.py:49-58: a 1x1 convolution is a per-pixel matrix product; checked against torch's conv2d."""
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 64, 8, 16)).astype(f32)
    w = (rng.standard_normal((64, 18)) * 0.1).astype(f32)
    b = rng.standard_normal(18).astype(f32)
    got = oracle.heatmap_head(x, w, b)
    assert got.shape == (2, 8, 16, 18)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w.T.copy()).double()[:, :, None, None],
                                     torch.from_numpy(b).double()).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)
    one = oracle.heatmap_head(np.ones((1, 64, 4, 16), f32), np.full((64, 18), 0.5, f32), np.zeros(18, f32))
    assert np.all(one == 32.0)


# ---------------------------------------------------------------------------- (8) get_keypoints: GOLDEN (reference output)
def test_get_keypoints_golden_vectors_from_reference():
    g = np.load(os.path.join(HERE, "golden", "get_keypoints.npz"))
    n = int(g["n"])
    assert n >= 12
    for k in range(n):
        hm, box, thr, want = g[f"hm_{k}"], g[f"box_{k}"], float(g[f"thr_{k}"]), g[f"out_{k}"]
        got = oracle.get_keypoints(hm, box, thr)
        assert np.array_equal(got, want), f"case {k}"
        got2 = spec_np.get_keypoints(hm, box.tolist(), thr)
        assert np.array_equal(got2, want), f"case {k} (numpy spec)"


def test_get_keypoints_known_answers():
    hm = np.zeros((56, 36, 17), f32)
    hm[20, 11, :] = 0.95
    out = oracle.get_keypoints(hm, (0, 0, 560, 360), 0.9)
    assert (out == np.array([110, 200, 1])).all()
    assert (oracle.get_keypoints(hm, (0, 0, 560, 360), 0.95) == 0).all()      # max <= threshold


# ---------------------------------------------------------------------------- whole path + Detector filter
def test_full_path_shapes_and_filter(prn_weights):
    from multiposenet_b200 import synthetic
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl, batch=1)
    out = oracle.full_path(inp["class_logits"], inp["encoded_boxes"], inp["heatmap_logits"], wl.height, wl.width,
                           *prn_weights, thr=0.3, iou_thr=0.6, max_det=25)
    N = int(out["num_boxes"].sum())
    assert N >= 1 and out["keypoint_scores"].shape == (N, 17) and out["keypoint_positions"].shape == (N, 17, 2)
    assert out["keypoint_heatmaps"].shape == (1, 64, 64, 17) and out["segmentation_masks"].shape == (1, 64, 64)
    assert (np.diff(out["scores"][0][:N]) <= 0).all()
    assert ((out["keypoint_positions"] >= 0) & (out["keypoint_positions"] < 1)).all()
    import ctypes
    keep = np.zeros(N, np.int32)
    k = oracle.lib().orc_detector_filter(out["scores"][0].ctypes.data_as(ctypes.POINTER(ctypes.c_float)), N,
                                         ctypes.c_float(0.5), keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert k == int((out["scores"][0][:N] > 0.5).sum())
