"""Oracle (CPU) and device (GPU) against tests/golden/graph_goldens.npz: vectors produced by EXECUTING THE REFERENCE'S OWN
CODE (detector/anchor_generator.py, detector/utils/box_utils.py, detector/box_predictor.py, detector/retinanet.py,
detector/utils/nms.py, detector/prn.py, create_pb.py:90-142, inference/detector.py:49-59) under the numpy-backed
TensorFlow stand-in of tests/golden/tf_numpy_shim.py (generator: tests/golden/make_graph_goldens.py).

Bars: everything the reference computes with element-wise float32 / integer ops is compared BIT FOR BIT (anchors, decode
around the shared exp recipe, clip, keep sets, padding, min-max normalisation, weak-channel mask, person list, argmax
positions, post-filter); values behind a TensorFlow library kernel whose summation order is its own (softmax denominator,
dense layers) at 1e-6 / 1e-4 relative, the tolerance stated at the assertion.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle
from multiposenet_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "graph_goldens.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, dtype=f32).view(np.uint32)


def assert_bits(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = bits(got) != bits(want) if want.dtype == f32 else got != want
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} differ, first at {np.argwhere(bad)[0]}"


def tiny_inputs(G):
    """The generator's inputs, regenerated (digests checked) together with the NCHW level tensors."""
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl, replicate=7)
    digests = [sha(inp[k]) for k in ("class_logits", "encoded_boxes", "heatmap_logits")]
    assert digests == [str(s) for s in G["pred_inputs_sha"]], "synthetic generator changed: regenerate the goldens"
    B, n_loc = wl.batch, wl.n_loc
    cls_levels, box_levels, off = [], [], 0
    for lvl, s_ in enumerate(wl.strides):      # NCHW: class channel k, box channel k*4 + coord (detector/box_predictor.py:77-85)
        gh, gw = -(-wl.height // s_), -(-wl.width // s_)
        n = gh * gw * n_loc
        cls_levels.append(np.ascontiguousarray(inp["class_logits"][:, off:off + n].reshape(B, gh, gw, n_loc).transpose(0, 3, 1, 2)))
        box_levels.append(np.ascontiguousarray(inp["encoded_boxes"][:, off:off + n].reshape(B, gh, gw, n_loc * 4).transpose(0, 3, 1, 2)))
        assert sha(cls_levels[-1]) + sha(box_levels[-1]) == str(G["pred_levels_sha"][lvl])
        off += n
    return wl, inp, cls_levels, box_levels


def golden_kh(G, inp):
    kh, seg, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
    kh = kh.copy()
    kh[1, :, :, 4] *= f32(0.15)
    assert sha(kh) == str(G["graph_keypoint_heatmaps_sha"])
    return kh, seg


def prn_weights_of_goldens():
    return synthetic.make_prn_weights(bias_std=0.01)


# =========================================================================================== CPU: oracle vs the reference
def test_oracle_anchors_equal_reference_anchor_generator(G):
    """detector/anchor_generator.py:12-166, 7 image sizes, 6 and 9 anchors per location: every bit."""
    for i in range(int(G["anchors_n"])):
        H, W = (int(v) for v in G[f"anchors_hw_{i}"])
        mult = tuple(float(m) for m in G[f"anchors_mult_{i}"])
        a = oracle.anchors(H, W, multipliers=mult)
        assert a.shape[0] == int(G[f"anchors_count_{i}"])
        assert sha(a) == str(G[f"anchors_sha_{i}"]), f"anchors {H}x{W} {len(mult) * 3}/location"
        rows = G[f"anchors_rows_{i}"]
        assert_bits(a if a.shape[0] < 20000 else a[::61], rows, "anchor rows")
        per_level = [-(-H // s) * -(-W // s) * len(mult) * 3 for s in oracle.STRIDES]
        assert per_level == [int(v) for v in G[f"anchors_per_level_{i}"]]


def test_oracle_decode_equals_reference_box_utils(G):
    """detector/utils/box_utils.py:112-139 (+ the clip of detector/utils/nms.py:36, which the oracle's decode includes)."""
    want = np.clip(G["decode_boxes"], f32(0), f32(1))
    assert_bits(oracle.decode(G["decode_codes"], G["decode_anchors"]), want, "decode")
    # tf.exp is a library kernel (the stand-in used the shared recipe): the same boxes with libm's exp agree to 1e-6
    c, a = G["decode_codes"].astype(np.float64), G["decode_anchors"].astype(np.float64)
    ha, wa = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    cy, cx = c[:, 0] / 10 * ha + a[:, 0] + 0.5 * ha, c[:, 1] / 10 * wa + a[:, 1] + 0.5 * wa
    h, w = np.exp(c[:, 2] / 5) * ha, np.exp(c[:, 3] / 5) * wa
    libm = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], 1)
    np.testing.assert_allclose(G["decode_boxes"], libm, rtol=1e-6, atol=1e-6)
    # round trip through the reference's encode (box_utils.py:79-109)
    back = oracle.decode(G["encode_codes"], G["decode_anchors"])
    np.testing.assert_allclose(back, G["encode_boxes"], atol=2e-6)


def test_oracle_detect_equals_reference_get_predictions(G):
    """detector/retinanet.py:56-81 + detector/utils/nms.py:6-61 on the concatenated layout the reference's
    reshape_and_concatenate (detector/box_predictor.py:53-90) produced from NCHW level tensors."""
    wl, inp, _, _ = tiny_inputs(G)
    anc = oracle.anchors(wl.height, wl.width)
    for k in range(int(G["pred_n"])):
        thr, iou_thr, max_det = G[f"pred_params_{k}"]
        got = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, float(thr), float(iou_thr), int(max_det))
        assert_bits(got["num_boxes"], G[f"pred_num_boxes_{k}"], f"num_boxes case {k}")
        assert_bits(got["boxes"], G[f"pred_boxes_{k}"], f"boxes case {k}")
        assert_bits(got["scores"], G[f"pred_scores_{k}"], f"scores case {k}")
    assert int(G["pred_num_boxes_0"].sum()) >= 4


def test_oracle_normalisation_person_list_and_crops_equal_reference_graph(G):
    """create_pb.py:90-94 (normalise + mask), :96-103 (boxes, box_ind), :106-109 (crop_and_resize call)."""
    wl, inp, _, _ = tiny_inputs(G)
    kh, _ = golden_kh(G, inp)
    mn, mx = kh.min((1, 2)), kh.max((1, 2))
    assert_bits(mn, G["graph_min"], "min")
    assert_bits(mx, G["graph_max"], "max")
    h, w = kh.shape[1:3]
    # the oracle normalises taps inside its crop: an identity crop at the map's own size returns the normalised map
    ident = np.array([[0, 0, 1, 1]], f32)
    norm1 = oracle.crop_and_resize(kh, ident, np.array([1], np.int32), (h, w), mn, mx)[0]
    assert_bits(norm1, G["graph_normalised_image1"], "normalised map of image 1")
    assert not norm1[:, :, 4].any()                          # the weak channel (max <= 0.2) is zeroed
    norm = np.stack([oracle.crop_and_resize(kh, ident, np.array([b], np.int32), (h, w), mn, mx)[0] for b in range(2)])
    assert sha(norm) == str(G["graph_normalised_sha"])
    anc = oracle.anchors(wl.height, wl.width)
    det = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, 0.3, 0.6, 25)
    pb = np.concatenate([det["boxes"][b, :det["num_boxes"][b]] for b in range(2)])
    pi = np.concatenate([np.full(det["num_boxes"][b], b, np.int32) for b in range(2)])
    assert_bits(pb, G["graph_person_boxes"], "person boxes")
    assert_bits(pi, G["graph_person_image"], "box_ind")
    crops = oracle.crop_and_resize(kh, pb, pi, (56, 36), mn, mx)
    assert sha(crops) == str(G["graph_crops_sha"])
    assert_bits(crops[:2], G["graph_crops_rows"], "crops")


def test_oracle_prn_and_keypoint_decode_equal_reference_graph(G):
    """detector/prn.py:5-25 (1e-5: the reference's float32 matmul has its own summation order; the oracle accumulates
    in float64) and create_pb.py:115-142 (positions bit for bit, scores 1e-6: softmax denominator order)."""
    W1, b1, W2, b2 = prn_weights_of_goldens()
    got = oracle.prn(G["graph_crops_rows"], W1, b1, W2, b2, mode=0)
    want = G["graph_prn_logits_rows"]
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    lg = G["graph_decode_logits_rows"]
    s, pos, arg, gap = oracle.keypoint_decode(lg)
    assert_bits(pos, G["graph_keypoint_positions"][:3], "keypoint_positions")
    np.testing.assert_allclose(s, G["graph_keypoint_scores"][:3], rtol=1e-6)
    assert arg[0, 5] == 0 and arg[1, 9] == 10 * 36 + 3 and arg[2, 16] == 2015


def test_detector_post_filter_equals_reference_detector_call(G):
    """inference/detector.py:49-59 -- the product's host-side mirror (multiposenet_b200.detector.post_filter)."""
    from multiposenet_b200.detector import post_filter
    raw = {k[len("filter_in_"):]: G[k] for k in G.files if k.startswith("filter_in_")}
    n = int(raw["num_boxes"][0])
    raw["person_offsets"] = np.array([0, n], np.int32)
    out = post_filter(raw, float(G["filter_threshold"]))
    names = [k[len("filter_out_"):] for k in G.files if k.startswith("filter_out_")]
    assert sorted(names) == sorted(out)
    for name in names:
        assert_bits(out[name], G[f"filter_out_{name}"], name)
    assert int(out["num_boxes"]) == n and len(out["scores"]) < n        # num_boxes stays unfiltered (:54-59)
    # the score equal to the threshold is dropped: strict > (:55)
    assert float(G["filter_threshold"]) in [float(v) for v in raw["scores"][0, :n]]
    assert float(G["filter_threshold"]) not in [float(v) for v in out["scores"]]


# =========================================================================================== GPU: device vs the reference
def _cuda(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def det():
    from multiposenet_b200 import Detector, DetectorConfig
    d = Detector(prn_weights_of_goldens(), DetectorConfig(max_batch=2, max_height=1024, max_width=1024, max_boxes=25))
    yield d
    d.close()


@pytest.mark.gpu
def test_device_anchors_equal_reference_anchor_generator(G):
    from multiposenet_b200 import Detector, DetectorConfig
    for mult_len in (2, 3):
        mult = synthetic.MULT_6 if mult_len == 2 else synthetic.MULT_9
        d = Detector(None, DetectorConfig(max_batch=1, max_height=1024, max_width=1024, scale_multipliers=mult))
        try:
            for i in range(int(G["anchors_n"])):
                if len(G[f"anchors_mult_{i}"]) != mult_len:
                    continue
                H, W = (int(v) for v in G[f"anchors_hw_{i}"])
                assert sha(d.anchors(H, W).cpu().numpy()) == str(G[f"anchors_sha_{i}"]), f"device anchors {H}x{W}"
        finally:
            d.close()


@pytest.mark.gpu
def test_device_detect_equals_reference_get_predictions(G, det):
    """Both layouts: the concatenated tensors and the raw NCHW level tensors (reshape_and_concatenate fused away)."""
    wl, inp, cls_levels, box_levels = tiny_inputs(G)
    for k in range(int(G["pred_n"])):
        thr, iou_thr, max_det = G[f"pred_params_{k}"]
        for enc, cls in ((_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"])),
                         ([_cuda(b) for b in box_levels], [_cuda(c) for c in cls_levels])):
            got = det.detect(enc, cls, (wl.height, wl.width), float(thr), float(iou_thr), int(max_det))
            assert_bits(got["num_boxes"].cpu().numpy(), G[f"pred_num_boxes_{k}"], f"num_boxes case {k}")
            assert_bits(got["boxes"].cpu().numpy(), G[f"pred_boxes_{k}"], f"boxes case {k}")
            assert_bits(got["scores"].cpu().numpy(), G[f"pred_scores_{k}"], f"scores case {k}")


@pytest.mark.gpu
def test_device_crops_equal_reference_graph(G, det):
    """create_pb.py:90-94 + :106-109 through BOTH device crop paths: taps normalised on the fly (mpn_crop) and the
    padded normalised map (mpn_crop_padded, the kernel mpn_run uses at these sizes) fed with the reference's map."""
    wl, inp, _, _ = tiny_inputs(G)
    kh, _ = golden_kh(G, inp)
    mm = np.stack([G["graph_min"], G["graph_max"]], -1)
    pb, pi = G["graph_person_boxes"], G["graph_person_image"]
    got = det.crop(_cuda(kh), _cuda(pb), _cuda(pi), _cuda(mm)).cpu().numpy()
    assert sha(got) == str(G["graph_crops_sha"])
    ident = np.array([[0, 0, 1, 1]], f32)
    h, w = kh.shape[1:3]
    norm = np.stack([oracle.crop_and_resize(kh, ident, np.array([b], np.int32), (h, w), G["graph_min"], G["graph_max"])[0]
                     for b in range(2)])
    assert sha(norm) == str(G["graph_normalised_sha"])       # == the reference's own normalised map (CPU test above)
    f, b16 = det.crop_padded(_cuda(norm), _cuda(pb), _cuda(pi))
    assert sha(f.cpu().numpy()) == str(G["graph_crops_sha"])
    import torch
    assert torch.equal(b16, f.to(torch.bfloat16))            # the bf16 copy is the RNE rounding of the fp32 crop


@pytest.mark.gpu
def test_device_prn_and_keypoint_decode_equal_reference_graph(G, det):
    want = G["graph_prn_logits_rows"]
    got = det.prn(_cuda(G["graph_crops_rows"]), "fp32").cpu().numpy()
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()          # north_star: 1e-4 relative (fp32)
    got = det.prn(_cuda(G["graph_crops_rows"]), "bf16").cpu().numpy()
    assert np.abs(got - want).max() <= 1e-2 * np.abs(want).max()          # north_star: 1e-2 (bf16 PRN)
    s, pos, arg = det.keypoint_decode(_cuda(G["graph_decode_logits_rows"]))
    assert_bits(pos.cpu().numpy(), G["graph_keypoint_positions"][:3], "keypoint_positions")
    np.testing.assert_allclose(s.cpu().numpy(), G["graph_keypoint_scores"][:3], rtol=1e-4)
