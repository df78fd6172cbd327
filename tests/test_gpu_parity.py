"""GPU parity: every stage of libmpn_b200.so, called through the C ABI (ctypes, multiposenet_b200.Detector), against the
CPU oracle (oracle/) on the same seeded inputs.

Bars (BASELINE.json north_star): NMS keep sets and keypoint argmax bit-exact, integer / index work bit-exact; boxes,
scores and fp32 PRN outputs within 1e-4 relative (the fp32-elementwise stages are in fact required to be BIT-EXACT here,
because host and device share one exp recipe and no contraction happens on either side); bf16 PRN within 1e-2.
Sizes are chosen so that the oracle finishes in seconds; BASELINE-size cases are covered by size-independent
properties in test_gpu_properties.py.
"""
import numpy as np
import pytest
import torch

import oracle
from multiposenet_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL_FP32 = 1e-4      # north_star: "decoded boxes, scores and PRN outputs within 1e-4 relative (fp32)"
RTOL_BF16 = 1e-2      # north_star: "or 1e-2 (bf16 PRN)"
ARGMAX_GAP = 1e-5     # argmax must be exact where the top-2 logit gap exceeds accumulation noise


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.dtype == np.float32:
        bad = _bits(got) != _bits(want)
    else:
        bad = got != want
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} elements differ; first at {np.argwhere(bad)[0]}: " \
                          f"{got[tuple(np.argwhere(bad)[0])]!r} vs {want[tuple(np.argwhere(bad)[0])]!r}"


@pytest.fixture(scope="module")
def det6(prn_weights):
    """n_loc = 6 (the reference's anchors), capacity 8 x 640 x 640, 128 detections."""
    from multiposenet_b200 import Detector, DetectorConfig
    d = Detector(prn_weights, DetectorConfig(max_batch=8, max_height=640, max_width=640, max_boxes=128))
    yield d
    d.close()


@pytest.fixture(scope="module")
def det9(prn_weights):
    """n_loc = 9 (BASELINE config 2)."""
    from multiposenet_b200 import Detector, DetectorConfig
    d = Detector(prn_weights, DetectorConfig(max_batch=8, max_height=640, max_width=640, max_boxes=128,
                                             scale_multipliers=synthetic.MULT_9))
    yield d
    d.close()


# ----------------------------------------------------------------------------------------------- exp / sigmoid
def test_device_exp_and_sigmoid_are_bit_identical_to_the_oracle(det6):
    rng = np.random.default_rng(1)
    x = np.concatenate([
        rng.normal(0, 3, 1 << 20), rng.uniform(-100, 100, 1 << 18), np.linspace(-0.01, 0.01, 4097),
        np.array([0.0, -0.0, 1.0, -1.0, 87.0, -87.0, 88.0, 88.5, -87.5, -103.0, 1e-30, -1e-30, np.inf, -np.inf])
    ]).astype(np.float32)
    assert_bit_equal(det6.device_exp(_cuda(x)).cpu().numpy(), oracle.expf(x), "exp")
    assert_bit_equal(det6.device_sigmoid(_cuda(x)).cpu().numpy(), oracle.sigmoidf(x), "sigmoid")
    nan = det6.device_exp(_cuda(np.array([np.nan], np.float32))).cpu().numpy()
    assert np.isnan(nan[0])


# ----------------------------------------------------------------------------------------------- anchors
@pytest.mark.parametrize("hw", [(128, 128), (256, 384), (512, 512), (640, 640)])
def test_anchors_bit_exact(det6, det9, hw):
    H, W = hw
    assert_bit_equal(det6.anchors(H, W).cpu().numpy(), oracle.anchors(H, W), "anchors n_loc=6")
    assert_bit_equal(det9.anchors(H, W).cpu().numpy(),
                     oracle.anchors(H, W, multipliers=synthetic.MULT_9), "anchors n_loc=9")
    assert det6.num_anchors(640, 640) == 51150 and det9.num_anchors(640, 640) == 76725


# ----------------------------------------------------------------------------------------------- detect (decode + NMS)
def _detect_case(det, wl, inputs, thr, iou, max_det):
    anc = oracle.anchors(wl.height, wl.width, wl.strides, wl.scales, wl.multipliers, wl.ratios)
    want = oracle.detect(inputs["class_logits"], inputs["encoded_boxes"], anc, thr, iou, max_det)
    got = det.detect(_cuda(inputs["encoded_boxes"]), _cuda(inputs["class_logits"]), (wl.height, wl.width),
                     score_threshold=thr, iou_threshold=iou, max_boxes=max_det)
    got = {k: v.cpu().numpy() for k, v in got.items()}
    assert_bit_equal(got["num_boxes"], want["num_boxes"], "num_boxes")
    for b in range(inputs["class_logits"].shape[0]):
        n = int(want["num_boxes"][b])
        assert_bit_equal(got["sel_anchor"][b, :n], want["sel_anchor"][b, :n], f"keep set image {b}")
        assert (got["sel_anchor"][b, n:] == -1).all()
    assert_bit_equal(got["boxes"], want["boxes"], "boxes")
    assert_bit_equal(got["scores"], want["scores"], "scores")
    return got, want


@pytest.mark.parametrize("key,batch", [("tiny", None), ("c1", None), ("c2_n6", 4)])
def test_detect_bit_exact_n6(det6, key, batch):
    wl = synthetic.WORKLOADS[key]
    inp = synthetic.make_inputs(wl, batch=batch)
    got, want = _detect_case(det6, wl, inp, wl.score_threshold, wl.iou_threshold, wl.max_detections)
    assert want["num_boxes"].min() >= 1
    assert_bit_equal(got["n_candidates"], want["n_conf"], "candidate count")


def test_detect_bit_exact_n9_and_crowded(det9):
    wl = synthetic.WORKLOADS["c2"]
    _detect_case(det9, wl, synthetic.make_inputs(wl, batch=3), 0.3, 0.5, 25)
    wl = synthetic.WORKLOADS["c3"]
    got, want = _detect_case(det9, wl, synthetic.make_inputs(wl, batch=2), 0.3, 0.5, 128)
    assert want["num_boxes"].max() > 60      # NMS-heavy: many kept boxes per image


def test_detect_many_candidates_uses_the_global_sort_path(det6):
    """A very low threshold makes > 8192 candidates per image (the shared-memory sort no longer fits) and the
    kept list saturates at max_detections."""
    wl = synthetic.WORKLOADS["c1"]
    inp = synthetic.make_inputs(wl)
    got, want = _detect_case(det6, wl, inp, 0.008, 0.5, 100)
    assert want["n_conf"][0] > 8192 and want["num_boxes"][0] == 100


def test_detect_edge_cases(det6):
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    A = wl.num_anchors
    # nothing confident -> num_boxes 0, all-zero padding (nms.py:47-52)
    cold = dict(inp, class_logits=np.full((2, A), -9.0, np.float32))
    got, _ = _detect_case(det6, wl, cold, 0.3, 0.6, 25)
    assert (got["num_boxes"] == 0).all() and not got["boxes"].any() and not got["scores"].any()
    # a score exactly equal to the threshold passes nms.py:30 (>=) but is never selected by the op (strict >)
    edge = dict(inp, class_logits=np.full((2, A), -9.0, np.float32))
    edge["class_logits"][:, 7] = 0.0          # sigmoid(0) = 0.5 exactly
    edge["class_logits"][:, 1000] = 2.0
    got, want = _detect_case(det6, wl, edge, 0.5, 0.6, 25)
    assert (got["num_boxes"] == 1).all() and (got["sel_anchor"][:, 0] == 1000).all()
    # iou_threshold 0 and 1, max_detections 1
    _detect_case(det6, wl, inp, 0.3, 0.0, 25)
    _detect_case(det6, wl, inp, 0.3, 1.0, 25)
    _detect_case(det6, wl, inp, 0.3, 0.6, 1)
    # exact score ties: the build's rule is (score desc, anchor index asc)
    tie = dict(inp, class_logits=np.full((2, A), -9.0, np.float32))
    tie["class_logits"][:, [50, 20, 900, 901]] = 1.25
    tie["encoded_boxes"] = np.zeros_like(inp["encoded_boxes"])
    _detect_case(det6, wl, tie, 0.3, 0.6, 25)
    # degenerate boxes: th = tw = -1000 collapses a box to its centre (exp -> +0), a huge ty pushes one outside the image
    # where the clip flattens it; TensorFlow's rule is that a zero-area box neither suppresses nor is suppressed, so the six
    # coincident points of one location are all kept next to the ordinary box that contains them
    flat = dict(inp, class_logits=np.full((2, A), -9.0, np.float32), encoded_boxes=np.zeros_like(inp["encoded_boxes"]))
    flat["class_logits"][:, 600:606] = np.linspace(2.0, 1.0, 6, dtype=np.float32)       # one location, all six shapes
    flat["encoded_boxes"][:, 600:606, 2:] = -1000.0
    flat["class_logits"][:, 612] = 0.5                                                   # the neighbouring location, intact
    flat["class_logits"][:, 30] = 3.0
    flat["encoded_boxes"][:, 30, 0] = 500.0                                              # far below the image: clipped flat
    got, want = _detect_case(det6, wl, flat, 0.3, 0.01, 25)
    assert (want["num_boxes"] == 8).all()
    pts = got["boxes"][0, 1:7]
    assert (pts[:, 0] == pts[:, 2]).all() and (pts[:, 1] == pts[:, 3]).all()


def test_detect_nchw_levels_equal_concatenated_layout(det6):
    """detector/box_predictor.py:53-90 fused away: raw per-level NCHW head outputs give the same result."""
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    B, n_loc = wl.batch, wl.n_loc
    cls_levels, box_levels, off = [], [], 0
    for s in wl.strides:
        gh, gw = -(-wl.height // s), -(-wl.width // s)
        n = gh * gw * n_loc
        c = inp["class_logits"][:, off:off + n].reshape(B, gh, gw, n_loc).transpose(0, 3, 1, 2)
        e = inp["encoded_boxes"][:, off:off + n].reshape(B, gh, gw, n_loc * 4).transpose(0, 3, 1, 2)
        cls_levels.append(_cuda(c))
        box_levels.append(_cuda(e))
        off += n
    a = det6.detect(box_levels, cls_levels, (wl.height, wl.width), 0.3, 0.6, 25)
    b = det6.detect(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), (wl.height, wl.width), 0.3, 0.6, 25)
    for k in a:
        assert_bit_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), k)


def test_full_path_from_raw_nchw_head_outputs(det6):
    """the whole path fed with the raw per-level NCHW class_net / box_net outputs (the
    reshape_and_concatenate of detector/box_predictor.py:53-90 never happens) equals the concatenated-layout result."""
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    B, n_loc = wl.batch, wl.n_loc
    cls_levels, box_levels, off = [], [], 0
    for s_ in wl.strides:
        gh, gw = -(-wl.height // s_), -(-wl.width // s_)
        n = gh * gw * n_loc
        cls_levels.append(_cuda(inp["class_logits"][:, off:off + n].reshape(B, gh, gw, n_loc).transpose(0, 3, 1, 2)))
        box_levels.append(_cuda(inp["encoded_boxes"][:, off:off + n].reshape(B, gh, gw, n_loc * 4).transpose(0, 3, 1, 2)))
        off += n
    hml = _cuda(inp["heatmap_logits"])
    a = det6.run_device(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), hml, prn_mode="bf16")
    torch.cuda.synchronize()
    a = {k: v.cpu().numpy().copy() for k, v in a.items()}
    b = det6.run_device(box_levels, cls_levels, hml, (wl.height, wl.width), prn_mode="bf16")
    torch.cuda.synchronize()
    n = int(a["person_offsets"][-1])
    for k, v in b.items():
        rows = n if k in ("keypoint_scores", "keypoint_positions") else None
        assert_bit_equal(v.cpu().numpy()[:rows], a[k][:rows], k)


# ----------------------------------------------------------------------------------------------- heatmaps
@pytest.mark.parametrize("key", ["tiny", "c1"])
def test_heatmaps_bit_exact(det6, key):
    wl = synthetic.WORKLOADS[key]
    hml = synthetic.make_inputs(wl)["heatmap_logits"]
    kh, seg, mn, mx = oracle.heatmaps(hml)
    gkh, gseg, gmm = det6.heatmaps(_cuda(hml))
    assert_bit_equal(gkh.cpu().numpy(), kh, "keypoint_heatmaps")
    assert_bit_equal(gseg.cpu().numpy(), seg, "segmentation_masks")
    gmm = gmm.cpu().numpy()
    assert_bit_equal(gmm[..., 0], mn, "min")
    assert_bit_equal(gmm[..., 1], mx, "max")


def test_heatmap_head_fused_with_activation(det6):
    """1x1 conv (64 -> 18) + bias + NCHW -> NHWC fused in front of sigmoid / split / min-max.
    Logits within 1e-4 of the float64 oracle; everything downstream bit-identical to running the un-fused heatmap
    stage on the kernel's own logits."""
    rng = np.random.default_rng(12)
    B, h, w = 2, 64, 64
    x = np.maximum(rng.standard_normal((B, 64, h, w)), 0).astype(np.float32)             # after final_bn + ReLU
    wt = (rng.standard_normal((64, 18)) * 0.4).astype(np.float32)
    bias = np.concatenate([np.full(17, -4.59511985, np.float32), np.zeros(1, np.float32)])   # keypoint_subnet.py:41-47
    lg, kh, seg, mm = det6.heatmap_head(_cuda(x), _cuda(wt), _cuda(bias))
    want = oracle.heatmap_head(x, wt, bias)
    np.testing.assert_allclose(lg.cpu().numpy(), want, rtol=RTOL_FP32, atol=RTOL_FP32 * np.abs(want).max())
    kh2, seg2, mm2 = det6.heatmaps(lg)
    assert_bit_equal(kh.cpu().numpy(), kh2.cpu().numpy(), "keypoint_heatmaps")
    assert_bit_equal(seg.cpu().numpy(), seg2.cpu().numpy(), "segmentation_masks")
    assert_bit_equal(mm.cpu().numpy(), mm2.cpu().numpy(), "min / max")
    okh, oseg, omn, omx = oracle.heatmaps(want)
    np.testing.assert_allclose(kh.cpu().numpy(), okh, rtol=RTOL_FP32, atol=1e-7)
    np.testing.assert_allclose(mm.cpu().numpy()[..., 1], omx, rtol=RTOL_FP32)
    # without the logits output (the tensor is never materialised) the other outputs are unchanged
    lg3, kh3, seg3, mm3 = det6.heatmap_head(_cuda(x), _cuda(wt), _cuda(bias), want_logits=False)
    assert lg3 is None
    assert_bit_equal(kh3.cpu().numpy(), kh.cpu().numpy(), "keypoint_heatmaps without logits")
    assert_bit_equal(mm3.cpu().numpy(), mm.cpu().numpy(), "min / max without logits")


# ----------------------------------------------------------------------------------------------- crop_and_resize
def test_crop_and_resize_bit_exact(det6):
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    kh, _, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
    rng = np.random.default_rng(5)
    boxes = np.concatenate([inp["gt_boxes"][0], inp["gt_boxes"][1],
                            np.array([[0, 0, 1, 1], [-0.2, -0.1, 0.5, 0.6], [0.5, 0.5, 1.3, 1.2], [0.3, 0.3, 0.3, 0.3],
                                      [0.9, 0.9, 0.1, 0.1]], np.float32),
                            rng.uniform(0, 1, (6, 4)).astype(np.float32)]).astype(np.float32)
    ind = np.concatenate([np.zeros(len(inp["gt_boxes"][0])), np.ones(len(inp["gt_boxes"][1])),
                          rng.integers(0, 2, 11)]).astype(np.int32)
    mm = np.stack([mn, mx], -1)
    want = oracle.crop_and_resize(kh, boxes, ind, (56, 36), mn, mx)
    got = det6.crop(_cuda(kh), _cuda(boxes), _cuda(ind), _cuda(mm)).cpu().numpy()
    assert_bit_equal(got, want, "normalised crops")
    want = oracle.crop_and_resize(kh, boxes, ind, (56, 36))
    got = det6.crop(_cuda(kh), _cuda(boxes), _cuda(ind)).cpu().numpy()
    assert_bit_equal(got, want, "plain crop_and_resize")
    assert det6.crop(_cuda(kh), _cuda(boxes[:0]), _cuda(ind[:0])).shape == (0, 56, 36, 17)


def test_crop_weak_channel_is_zeroed_and_identity_box(det6):
    rng = np.random.default_rng(6)
    kh = rng.uniform(0.0, 1.0, (1, 56, 36, 17)).astype(np.float32)
    kh[..., 3] *= 0.19                       # max <= 0.2 -> channel masked (create_pb.py:91,94)
    mn, mx = kh.min((1, 2)), kh.max((1, 2))
    box = np.array([[0, 0, 1, 1]], np.float32)
    ind = np.zeros(1, np.int32)
    got = det6.crop(_cuda(kh), _cuda(box), _cuda(ind), _cuda(np.stack([mn, mx], -1))).cpu().numpy()
    assert not got[..., 3].any()
    assert_bit_equal(got, oracle.crop_and_resize(kh, box, ind, (56, 36), mn, mx), "identity box, normalised")
    plain = det6.crop(_cuda(kh), _cuda(box), _cuda(ind)).cpu().numpy()
    assert_bit_equal(plain[0], kh[0], "box [0,0,1,1] on a 56x36 map is the identity")


def test_crop_other_crop_sizes_take_the_general_kernel():
    """The specialised crop kernels are compiled for the reference's 56 x 36 crop (create_pb.py:19); a handle configured
    with another crop size takes the runtime-sized kernel: same op, same bits."""
    from multiposenet_b200 import Detector, DetectorConfig
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    kh, _, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
    rng = np.random.default_rng(8)
    boxes = np.concatenate([inp["gt_boxes"][0], np.array([[0, 0, 1, 1], [-0.2, -0.1, 0.5, 0.6], [0.9, 0.9, 0.1, 0.1]], np.float32),
                            rng.uniform(0, 1, (5, 4)).astype(np.float32)]).astype(np.float32)
    ind = rng.integers(0, 2, len(boxes)).astype(np.int32)
    for size in ((8, 12), (24, 12)):          # crop_h * crop_w a multiple of 96: the PRN layers of the handle accept it
        det = Detector(None, DetectorConfig(max_batch=2, max_height=256, max_width=256, crop_size=size))
        try:
            got = det.crop(_cuda(kh), _cuda(boxes), _cuda(ind), _cuda(np.stack([mn, mx], -1))).cpu().numpy()
            assert got.shape == (len(boxes),) + size + (17,)
            assert_bit_equal(got, oracle.crop_and_resize(kh, boxes, ind, size, mn, mx), f"crop size {size}")
        finally:
            det.close()


# ----------------------------------------------------------------------------------------------- PRN
def _rel_err(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.mark.parametrize("n", [1, 5, 70])
def test_prn_fp32_within_1e4(det6, prn_weights, n):
    x = synthetic.make_crops(n, seed=11 + n)
    want = oracle.prn(x, *prn_weights, mode=0)
    got = det6.prn(_cuda(x), "fp32").cpu().numpy()
    assert got.shape == x.shape
    assert _rel_err(got, want) < RTOL_FP32
    np.testing.assert_allclose(got, want, rtol=RTOL_FP32, atol=RTOL_FP32 * np.abs(want).max())


@pytest.mark.parametrize("n", [1, 16, 17, 78, 80, 81, 200, 240, 250])
def test_prn_fp32_on_tensor_cores_is_fp32_accurate(prn_weights, monkeypatch, n):
    """fp32 mode, <= 240 persons: prn_split3.cu (every number as three bf16 parts, fp32 accumulation in tensor memory, one
    launch per group of 80 persons) must be as good as the SIMT fp32 kernels it replaces there (detector/prn.py:15-25):
    within 1e-5 of the fp64-accumulating oracle relative to the output scale (the north_star bar for fp32 is 1e-4), and
    within 1e-5 of the SIMT result.  250 persons take the SIMT kernels (both are launched, each exits outside its regime)."""
    from multiposenet_b200 import Detector, DetectorConfig
    x = synthetic.make_crops(n, seed=500 + n)
    want = oracle.prn(x, *prn_weights, mode=0)
    got = {}
    for simt in (False, True):
        if simt:
            monkeypatch.setenv("MPN_FP32_SIMT", "1")
        else:
            monkeypatch.delenv("MPN_FP32_SIMT", raising=False)
        det = Detector(prn_weights, DetectorConfig(max_batch=8, max_boxes=32, prn_mode="fp32", prn_modes_allocated=("fp32",)))
        try:
            before = det.launch_count()[1]
            for _ in range(2):
                out = det.prn(_cuda(x), "fp32")
            got[simt] = out.cpu().numpy()
            launches = (det.launch_count()[1] - before) // 2
        finally:
            det.close()
        if not simt:
            assert launches == (2 * ((n + 79) // 80) if n <= 240 else 2 * 3 + 3), launches   # parts + kernel per group (+ SIMT)
        else:
            assert launches == 3, launches
    scale = np.abs(want).max()
    err_tc, err_simt = np.abs(got[False] - want).max() / scale, np.abs(got[True] - want).max() / scale
    print(f"n={n}: max |error| / scale: tensor-core fp32 {err_tc:.2e}, SIMT fp32 {err_simt:.2e}")
    assert err_tc < 1e-5 and err_simt < 1e-5
    assert np.abs(got[False] - got[True]).max() / scale < 1e-5


@pytest.mark.parametrize("n", [1, 5, 130, 256, 300, 600])
def test_prn_bf16_tcgen05_within_1e2_and_close_to_bf16_oracle(det6, prn_weights, n):
    x = synthetic.make_crops(n, seed=31 + n)
    got = det6.prn(_cuda(x), "bf16").cpu().numpy()
    exact = oracle.prn(x, *prn_weights, mode=0)
    assert _rel_err(got, exact) < RTOL_BF16                     # the north_star bar
    emul = oracle.prn(x, *prn_weights, mode=1)                  # same operand rounding, fp64 accumulate
    assert _rel_err(got, emul) < 2e-3                           # what is left is accumulation order + y1 rounding flips


def test_prn_tiled_fallback_kernels(prn_weights, monkeypatch):
    """prn_tcgen05.cu (non-persistent 128 x 128 / 128 x 96 tiles) is the fallback for shapes the persistent large-batch
    kernels do not cover; MPN_NO_BIG_GEMM forces it so that it stays tested."""
    from multiposenet_b200 import Detector, DetectorConfig
    monkeypatch.setenv("MPN_NO_BIG_GEMM", "1")
    det = Detector(prn_weights, DetectorConfig(max_batch=4, max_boxes=128, prn_mode="bf16", prn_modes_allocated=("bf16",)))
    try:
        x = synthetic.make_crops(300, seed=77)
        got = det.prn(_cuda(x), "bf16").cpu().numpy()
        emul = oracle.prn(x, *prn_weights, mode=1)
        assert _rel_err(got, emul) < 2e-3
    finally:
        det.close()


@pytest.mark.parametrize("n", [70, 300, 700])
def test_prn_in_place_equals_out_of_place_bit_for_bit(det6, prn_weights, n):
    """logits == crops: the residual addition is a TMA reduce-add performed in L2 (single-kernel PRN for <= 256 persons,
    large-batch fc2 on CTA pairs above) instead of load / add / store -- the same single fp32 rounding, so the same
    bits; fp32 mode refuses to run in place."""
    x = synthetic.make_crops(n, seed=200 + n)
    want = det6.prn(_cuda(x), "bf16").cpu().numpy()
    buf = _cuda(x)
    got = det6.prn(buf, "bf16", inplace=True)
    assert got.data_ptr() == buf.data_ptr()
    assert_bit_equal(got.cpu().numpy(), want, f"in place n={n}")
    with pytest.raises(Exception):
        det6.prn(_cuda(x), "fp32", inplace=True)


def test_single_cta_fc2_and_no_pdl_paths_give_the_same_bits(prn_weights, monkeypatch):
    """MPN_FC2_PAIRS=0 (large-batch fc2 on single CTAs) and MPN_NO_PDL=1 (no programmatic dependent launch) are kept as
    switches: both must reproduce the default path bit for bit."""
    from multiposenet_b200 import Detector, DetectorConfig
    wl = synthetic.WORKLOADS["c3"]
    inp = synthetic.make_inputs(wl, batch=3)
    cfg = dict(max_batch=3, max_height=640, max_width=640, max_boxes=128, score_threshold=0.3, iou_threshold=0.5,
               scale_multipliers=wl.multipliers, prn_mode="bf16", prn_modes_allocated=("bf16",))
    outs = []
    for env in ({}, {"MPN_FC2_PAIRS": "0", "MPN_NO_PDL": "1"}):
        for k in ("MPN_FC2_PAIRS", "MPN_NO_PDL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        det = Detector(prn_weights, DetectorConfig(**cfg))
        try:
            side = torch.cuda.Stream()
            dev_in = [_cuda(inp[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
            with torch.cuda.stream(side):
                for _ in range(2):                      # second call = graph replay
                    o = det.run_device(*dev_in)
            side.synchronize()
            outs.append({k: v.cpu().numpy().copy() for k, v in o.items()})
        finally:
            det.close()
    n = int(outs[0]["person_offsets"][-1])
    assert n > 256
    for k in outs[0]:
        rows = n if k in ("keypoint_scores", "keypoint_positions") else None
        assert_bit_equal(outs[1][k][:rows], outs[0][k][:rows], k)


def test_prn_known_answers(prn_weights):
    """detector/prn.py:24: zero weights -> output == input; large negative b2 -> ReLU clamps -> output == input."""
    from multiposenet_b200 import Detector, DetectorConfig
    D = 56 * 36 * 17
    x = synthetic.make_crops(3, seed=99)
    det = Detector(None, DetectorConfig(max_batch=1, max_boxes=8))
    try:
        with pytest.raises(RuntimeError):
            det.prn(_cuda(x))                                    # MPN_ERR_NO_WEIGHTS
        det.set_prn_weights(np.zeros((D, 1024), np.float32), np.zeros(1024, np.float32),
                            np.zeros((1024, D), np.float32), np.zeros(D, np.float32))
        for mode in ("fp32", "bf16"):
            assert_bit_equal(det.prn(_cuda(x), mode).cpu().numpy(), x, f"zero weights {mode}")
        W1, b1, W2, _ = prn_weights
        det.set_prn_weights(W1, b1, W2, np.full(D, -1e4, np.float32))
        for mode in ("fp32", "bf16"):
            assert_bit_equal(det.prn(_cuda(x), mode).cpu().numpy(), x, f"clamped {mode}")
    finally:
        det.close()


# ----------------------------------------------------------------------------------------------- keypoint decode
def test_keypoint_decode_argmax_bit_exact(det6, prn_weights):
    x = synthetic.make_crops(9, seed=3)
    logits = oracle.prn(x, *prn_weights, mode=0)
    logits[0, :, :, 5] = 0.25                          # all-equal channel -> position (0,0), score 1/2016
    logits[1, 40, 7, 2] = 30.0                         # dominant peak
    logits[2, :, :, 9] = 0.0
    logits[2, 10, 3, 9] = logits[2, 30, 30, 9] = 1.5   # exact tie -> first index
    s, pos, arg, gap = oracle.keypoint_decode(logits)
    gs, gpos, garg = det6.keypoint_decode(_cuda(logits))
    assert_bit_equal(garg.cpu().numpy(), arg, "argmax")
    assert_bit_equal(gpos.cpu().numpy(), pos, "positions")
    np.testing.assert_allclose(gs.cpu().numpy(), s, rtol=RTOL_FP32)   # 2016-term fp32 denominator, different summation order
    assert arg[0, 5] == 0 and abs(s[0, 5] - 1 / 2016) < 1e-9
    assert arg[1, 2] == 40 * 36 + 7 and arg[2, 9] == 10 * 36 + 3


def test_keypoint_decode_near_ties_take_the_exact_rescan(det6):
    """Two DIFFERENT logits so close that exp(l - max) is 1.0f for both (only possible for small |max|): the reference
    rule (first index of the maximal fp32 probability) then picks the earlier position even when it holds the smaller
    logit.  The one-pass decode flags such channels and rescans them exactly."""
    rng = np.random.default_rng(5)
    for n in (6, 170):                                  # fewer / more persons than resident clusters
        logits = rng.normal(-3.0, 0.5, (n, 56, 36, 17)).astype(np.float32)
        cases = [(0, 0, 0.3, 1), (1, 3, 0.3, 2), (2, 5, 0.1, 1), (3, 7, 0.01, 3), (4, 16, 0.45, 1), (5, 9, 0.02, 40)]
        for person, ch, mx, ulps in cases:
            below = np.float32(mx)
            for _ in range(ulps):
                below = np.nextafter(below, np.float32(0))
            logits[person, 50, 30, ch] = mx             # the maximum, late (last quarter of the positions)
            logits[person, 2, 5, ch] = below            # slightly smaller, early (first quarter, another CTA)
            logits[person, 30, 1, ch] = below           # and in between
        s, pos, arg, gap = oracle.keypoint_decode(logits)
        gs, gpos, garg = det6.keypoint_decode(_cuda(logits))
        assert_bit_equal(garg.cpu().numpy(), arg, f"argmax n={n}")
        assert_bit_equal(gpos.cpu().numpy(), pos, f"positions n={n}")
        np.testing.assert_allclose(gs.cpu().numpy(), s, rtol=RTOL_FP32)
        picked = [int(arg[p, c]) for p, c, _, _ in cases]
        assert 2 * 36 + 5 in picked and 50 * 36 + 30 in picked      # both outcomes occur


def test_keypoint_decode_streaming_kernel_many_persons(det6):
    """>= 160 persons per call: the persistent, bulk-copy fed decode kernel (clusters walk the persons, three slabs in
    flight).  More persons than resident clusters, a count that is not a multiple of anything, planted ties."""
    rng = np.random.default_rng(77)
    n = 523
    logits = rng.normal(0.0, 1.0, (n, 56, 36, 17)).astype(np.float32)
    logits[5, :, :, 3] = -0.5                          # all-equal channel
    logits[200, 55, 35, 16] = 40.0                     # last position, last channel
    logits[522, 0, 0, 0] = logits[522, 17, 9, 0] = 9.0  # exact tie in the last person -> first index
    s, pos, arg, gap = oracle.keypoint_decode(logits)
    gs, gpos, garg = det6.keypoint_decode(_cuda(logits))
    assert_bit_equal(garg.cpu().numpy(), arg, "argmax")
    assert_bit_equal(gpos.cpu().numpy(), pos, "positions")
    np.testing.assert_allclose(gs.cpu().numpy(), s, rtol=RTOL_FP32)
    assert arg[5, 3] == 0 and arg[200, 16] == 2015 and arg[522, 0] == 0


def test_get_keypoints_matches_reference_golden_vectors(det6):
    """tests/golden/get_keypoints.npz holds outputs of the reference's own inference/utils.py::get_keypoints."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "get_keypoints.npz"))
    n = int(g["n"])
    assert n >= 12
    for i in range(n):
        hm, box, thr, want = g[f"hm_{i}"], g[f"box_{i}"], float(g[f"thr_{i}"]), g[f"out_{i}"]
        got = det6.get_keypoints(hm, box, thr)
        assert_bit_equal(got, want.astype(np.int32), f"golden case {i}")
        assert_bit_equal(got, oracle.get_keypoints(hm, box, thr), f"oracle case {i}")


# ----------------------------------------------------------------------------------------------- the whole path
def _full_case(det, wl, inp, prn_weights, mode, host):
    thr, iou, md = wl.score_threshold, wl.iou_threshold, wl.max_detections
    want = oracle.full_path(inp["class_logits"], inp["encoded_boxes"], inp["heatmap_logits"], wl.height, wl.width,
                            *prn_weights, thr=thr, iou_thr=iou, max_det=md, prn_mode=1 if mode == "bf16" else 0,
                            multipliers=wl.multipliers, ratios=wl.ratios)
    if host:
        bufs = det.run_host_async(inp["encoded_boxes"], inp["class_logits"], inp["heatmap_logits"],
                                  score_threshold=thr, iou_threshold=iou, max_boxes=md, prn_mode=mode)
        det.synchronize()
        got = {k: v.numpy().copy() for k, v in bufs.items()}
    else:
        out = det.run_device(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), _cuda(inp["heatmap_logits"]),
                             score_threshold=thr, iou_threshold=iou, max_boxes=md, prn_mode=mode)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
    B = inp["class_logits"].shape[0]
    assert_bit_equal(got["num_boxes"], want["num_boxes"], "num_boxes")
    assert_bit_equal(got["boxes"], want["boxes"], "boxes")
    assert_bit_equal(got["scores"], want["scores"], "scores")
    assert_bit_equal(got["keypoint_heatmaps"], want["keypoint_heatmaps"], "keypoint_heatmaps")
    assert_bit_equal(got["segmentation_masks"], want["segmentation_masks"], "segmentation_masks")
    offs = np.concatenate([[0], np.cumsum(want["num_boxes"])]).astype(np.int32)
    assert_bit_equal(got["person_offsets"], offs, "person_offsets")
    N = int(offs[B])
    assert N >= B
    ks, kp = got["keypoint_scores"][:N], got["keypoint_positions"][:N]
    decided = want["keypoint_gap"] > (ARGMAX_GAP if mode == "fp32" else 2e-2)
    assert decided.mean() > 0.5
    wrong = (kp != want["keypoint_positions"]).any(-1) & decided
    assert not wrong.any(), f"{int(wrong.sum())} keypoint argmax mismatches among {int(decided.sum())} decided"
    # ... and NO keypoint is exempt: where the oracle's top-2 gap is inside the PRN's tolerance the device may pick another
    # position, but only one whose oracle logit is within that tolerance of the oracle's maximum (a legitimate near-tie)
    flat = want["logits"].reshape(N, 56 * 36, 17)
    yx = np.rint(kp * np.array([56, 36], np.float32)).astype(np.int64)
    at_dev = np.take_along_axis(flat, (yx[..., 0] * 36 + yx[..., 1])[:, None, :], 1)[:, 0]
    short = (flat.max(1) - at_dev) / np.abs(flat).max((1, 2))[:, None]
    differ = (kp != want["keypoint_positions"]).any(-1)
    print(f"{mode}: {int(differ.sum())} of {differ.size} keypoints differ from the oracle's argmax, worst shortfall {short.max():.2e}")
    assert short.max() <= 2 * (RTOL_FP32 if mode == "fp32" else RTOL_BF16), f"shortfall {short.max():.2e}"
    tol = RTOL_FP32 if mode == "fp32" else RTOL_BF16
    np.testing.assert_allclose(ks[decided], want["keypoint_scores"][decided], rtol=50 * tol, atol=1e-7)
    return got, want


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("host", [False, True])
def test_full_path_tiny(det6, prn_weights, mode, host):
    wl = synthetic.WORKLOADS["tiny"]
    _full_case(det6, wl, synthetic.make_inputs(wl), prn_weights, mode, host)


def test_full_path_c1_fp32(det6, prn_weights):
    wl = synthetic.WORKLOADS["c1"]
    _full_case(det6, wl, synthetic.make_inputs(wl), prn_weights, "fp32", False)


def test_full_path_c2_bf16(det9, prn_weights):
    wl = synthetic.WORKLOADS["c2"]
    _full_case(det9, wl, synthetic.make_inputs(wl, batch=2), prn_weights, "bf16", True)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_full_path_with_no_person_at_all(det6, prn_weights, mode):
    """Nothing confident in any image: zero persons flow through crop / PRN / decode (every grid exits on the device-side
    count, the cooperative PRN kernel included), the seven outputs keep their contract, and the handle is fine for the
    next, normal call (counters re-armed, in-place buffers untouched)."""
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    cold = dict(inp, class_logits=np.full_like(inp["class_logits"], -9.0))
    dev = [_cuda(cold[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):                               # direct launches + capture, then graph replay
            out = det6.run_device(*dev, score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold,
                                  max_boxes=wl.max_detections, prn_mode=mode)
    side.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    assert (got["num_boxes"] == 0).all() and (got["person_offsets"] == 0).all()
    assert not got["boxes"].any() and not got["scores"].any()
    want_kh, want_seg, _, _ = oracle.heatmaps(cold["heatmap_logits"])
    assert_bit_equal(got["keypoint_heatmaps"], want_kh, "keypoint_heatmaps")
    assert_bit_equal(got["segmentation_masks"], want_seg, "segmentation_masks")
    _full_case(det6, wl, inp, prn_weights, mode, host=False)


def test_host_pipeline_pinned_inputs_three_in_flight(det6, prn_weights):
    """mpn_submit_host / mpn_wait with pinned inputs (box codes gathered in place over PCIe, never copied) and three
    calls in flight give the same bits as one call at a time with pageable inputs (staged copies)."""
    wl = synthetic.WORKLOADS["tiny"]
    sets = [synthetic.make_inputs(wl, replicate=r) for r in range(5)]
    names = ("encoded_boxes", "class_logits", "heatmap_logits")
    ref = []
    for s in sets:
        bufs = det6.run_host_async(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], prn_mode="bf16")
        det6.synchronize()
        h2d_copy, _ = det6.host_traffic()
        ref.append({k: v.numpy().copy() for k, v in bufs.items()})
    pinned = [{k: torch.from_numpy(s[k]).pin_memory() for k in names} for s in sets]
    got, pend = [], []
    for s in pinned:
        pend.append(det6.submit_host(s["encoded_boxes"], s["class_logits"], s["heatmap_logits"], prn_mode="bf16"))
        if len(pend) == 3:
            t, bufs = pend.pop(0)
            det6.wait(t)
            got.append({k: v.numpy().copy() for k, v in bufs.items()})
    for t, bufs in pend:
        det6.wait(t)
        got.append({k: v.numpy().copy() for k, v in bufs.items()})
    h2d_zero_copy, d2h = det6.host_traffic()
    assert h2d_copy - h2d_zero_copy == sets[0]["encoded_boxes"].nbytes and d2h > 0
    assert len(got) == len(ref) == 5
    for g, r in zip(got, ref):
        n = int(r["person_offsets"][-1])
        for k in r:
            rows = n if k.startswith("keypoint_s") or k.startswith("keypoint_p") else None
            assert_bit_equal(g[k][:rows], r[k][:rows], k)
    assert not all(np.array_equal(ref[0]["boxes"], r["boxes"]) for r in ref[1:])


def test_graph_replay_on_a_side_stream_equals_direct_launches(det6, prn_weights):
    """On a non-default stream mpn_run captures the kernel sequence (forked front halves included) into a CUDA graph
    and replays it; the results must be the bits of the direct launches on the default stream."""
    wl = synthetic.WORKLOADS["tiny"]
    sets = [synthetic.make_inputs(wl, replicate=r) for r in range(2)]
    dev = [{k: _cuda(s[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")} for s in sets]
    direct = []
    for d in dev:
        out = det6.run_device(d["encoded_boxes"], d["class_logits"], d["heatmap_logits"], prn_mode="bf16")
        torch.cuda.synchronize()
        direct.append({k: v.cpu().numpy().copy() for k, v in out.items()})
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(3):                        # first pass captures, later passes replay
        for d, want in zip(dev, direct):
            with torch.cuda.stream(side):
                out = det6.run_device(d["encoded_boxes"], d["class_logits"], d["heatmap_logits"], prn_mode="bf16")
            side.synchronize()
            n = int(want["person_offsets"][-1])
            for k, v in out.items():
                rows = n if k in ("keypoint_scores", "keypoint_positions") else None
                assert_bit_equal(v.cpu().numpy()[:rows], want[k][:rows], f"{k} (pass {rep})")
    last, _ = det6.launch_count()
    assert last >= 7


def test_detector_call_keeps_the_reference_output_contract(det6, prn_weights):
    """inference/detector.py:36-61 for one image: batch dimension stripped, rows filtered by score > threshold,
    num_boxes left unfiltered."""
    wl = synthetic.WORKLOADS["c1"]
    inp = synthetic.make_inputs(wl)
    from multiposenet_b200 import OUTPUT_NAMES
    want = oracle.full_path(inp["class_logits"], inp["encoded_boxes"], inp["heatmap_logits"], 512, 512, *prn_weights,
                            thr=0.3, iou_thr=0.6, max_det=128)
    n = int(want["num_boxes"][0])
    cut = float(np.median(want["scores"][0, :n]))       # a post-filter threshold that removes about half of the rows
    out = det6(inp["encoded_boxes"], inp["class_logits"], inp["heatmap_logits"], score_threshold=cut)
    assert sorted(out) == sorted(OUTPUT_NAMES)
    keep = want["scores"][0, :n] > cut
    assert 0 < keep.sum() < n
    assert int(out["num_boxes"]) == n
    assert_bit_equal(out["boxes"], want["boxes"][0, :n][keep], "boxes")
    assert_bit_equal(out["scores"], want["scores"][0, :n][keep], "scores")
    decided = want["keypoint_gap"][keep] > ARGMAX_GAP
    assert_bit_equal(out["keypoint_positions"][decided], want["keypoint_positions"][keep][decided], "keypoint_positions")
    assert out["keypoint_scores"].shape == (int(keep.sum()), 17)
    assert out["keypoint_positions"].shape == (int(keep.sum()), 17, 2)
    assert out["keypoint_heatmaps"].shape == (128, 128, 17) and out["segmentation_masks"].shape == (128, 128)
    # same call with device tensors
    out2 = det6(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), _cuda(inp["heatmap_logits"]),
                score_threshold=cut)
    for k in OUTPUT_NAMES:
        assert_bit_equal(out2[k], out[k], k)


def test_error_behaviour(det6):
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl)
    with pytest.raises(ValueError):          # inference/detector.py:45
        det6.detect(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), (200, 256))
    with pytest.raises(ValueError):          # wrong anchor count for the image size
        det6.detect(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), (384, 256))
    with pytest.raises(ValueError):          # capacity
        det6.detect(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), (256, 256), max_boxes=4096)
    with pytest.raises(ValueError):          # host tensors on the device path
        det6.run_device(inp["encoded_boxes"], inp["class_logits"], inp["heatmap_logits"])
    last, total = det6.launch_count()
    assert total > 0
