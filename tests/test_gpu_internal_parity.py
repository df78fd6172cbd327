"""GPU parity of the kernels that ACTUALLY RUN inside mpn_run, element by element, and of the whole path at BASELINE.json's
full sizes through row-sampled oracle checks.

  * the two-pass heatmap stage (logit min / max -> activation + normalisation in one pass: create_pb.py:73-76, 90-94) and the
    padded crop kernel (create_pb.py:106-109), as stages (mpn_heatmaps_normalised, mpn_crop_padded) and as the internal
    buffers a full run leaves behind (mpn_debug_fetch): bit-identical to the oracle, bf16 copy == RNE of the fp32 crop;
  * c2 (B = 8), c3 (B = 32), c4 (B = 64, per-tap crop path) through mpn_run: detection / heatmaps against the full oracle,
    crops / PRN logits / keypoints against the oracle on ~64 sampled persons (rows are independent: detector/prn.py:17-24,
    create_pb.py:106-142);
  * c5 (PRN only) at 1 000 / 10 000 / 100 000 persons, sampled rows;
  * bf16 mode: EVERY keypoint is checked, not only the clearly decided ones -- the device's position must be a legitimate
    near-tie of the oracle's logits (within the bf16 tolerance of the oracle's maximum).
"""
import numpy as np
import pytest
import torch

import oracle
from multiposenet_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL_FP32 = 1e-4      # north_star: PRN outputs within 1e-4 relative (fp32)
RTOL_BF16 = 1e-2      # north_star: 1e-2 (bf16 PRN)
RTOL_BF16_EMUL = 2e-3  # against the oracle that rounds the same operands to bf16: accumulation order + y1 rounding flips


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits(got, want, what, nan_ok=False):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = _bits(got) != _bits(want) if want.dtype == np.float32 else got != want
    if nan_ok:
        bad &= ~(np.isnan(got) & np.isnan(want))
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} differ; first at {np.argwhere(bad)[0]}: " \
                          f"{got[tuple(np.argwhere(bad)[0])]!r} vs {want[tuple(np.argwhere(bad)[0])]!r}"


def oracle_normalised(kh, mn, mx):
    """create_pb.py:93-94 over whole maps with the oracle: an identity crop at the map's own size (all lerp weights 0)."""
    B, h, w, _ = kh.shape
    ident = np.array([[0, 0, 1, 1]], np.float32)
    return np.stack([oracle.crop_and_resize(kh, ident, np.array([b], np.int32), (h, w), mn, mx)[0] for b in range(B)])


@pytest.fixture(scope="module")
def weights():
    return synthetic.make_prn_weights(bias_std=0.01)


def make_det(weights, wl, batch, **kw):
    from multiposenet_b200 import Detector, DetectorConfig
    cfg = dict(max_batch=batch, max_height=wl.height, max_width=wl.width, max_boxes=wl.max_detections,
               score_threshold=wl.score_threshold, iou_threshold=wl.iou_threshold, scale_multipliers=wl.multipliers,
               prn_mode="bf16")
    cfg.update(kw)
    return Detector(weights, DetectorConfig(**cfg))


# ------------------------------------------------------------------------------------------- exhaustive monotonicity
def test_device_sigmoid_is_monotone_over_all_floats(weights):
    """What the two-pass heatmap stage rests on: max sigmoid(l) == sigmoid(max l) bit for bit.  All 2^32 ordered float
    keys, neighbours compared; the packed-pair form of the recipe must also equal the scalar form everywhere."""
    det = make_det(weights, synthetic.WORKLOADS["tiny"], 1)
    try:
        total = 0
        for q in range(16):                                   # 16 launches of 2^28 neighbour pairs (key, key + 1)
            total += det.sigmoid_monotone_violations(q << 28, 1 << 28)
        assert total == 0
    finally:
        det.close()


# ------------------------------------------------------------------------------------------- heatmap stage, two-pass form
@pytest.mark.parametrize("key,batch", [("tiny", None), ("c1", None), ("c2", 8), ("c3", 3)])
def test_heatmaps_normalised_bit_exact(weights, key, batch):
    wl = synthetic.WORKLOADS[key]
    B = wl.batch if batch is None else batch
    hml = synthetic.make_inputs(wl, batch=B)["heatmap_logits"]
    det = make_det(weights, wl, B)
    try:
        kh, seg, mn, mx = oracle.heatmaps(hml)
        gkh, gseg, gmm, gnh = det.heatmaps_normalised(_cuda(hml))
        assert_bits(gkh.cpu().numpy(), kh, "keypoint_heatmaps")
        assert_bits(gseg.cpu().numpy(), seg, "segmentation_masks")
        gmm = gmm.cpu().numpy()
        assert_bits(gmm[..., 0], mn, "min (from the minimum logit)")
        assert_bits(gmm[..., 1], mx, "max (from the maximum logit)")
        gnh = gnh.cpu().numpy()
        assert_bits(gnh[..., :17], oracle_normalised(kh, mn, mx), "normalised map")
        assert not gnh[..., 17:].any()                        # padding is never written
        # the one-pass kernel (per-tap crop path) gives the same activations and the same min / max
        kh1, seg1, mm1 = det.heatmaps(_cuda(hml))
        assert torch.equal(kh1, gkh) and torch.equal(seg1, gseg)
        assert_bits(mm1.cpu().numpy(), gmm, "min / max: one-pass == two-pass")
    finally:
        det.close()


def test_heatmaps_normalised_edge_channels(weights):
    """Weak channel (max <= 0.2 -> zeroed), constant channel (M == m -> 0/0 = NaN, which the reference propagates:
    create_pb.py:93), a channel whose range is tiny (true division instead of the reciprocal shortcut), saturated logits,
    negative-only and positive-only channels, +-0."""
    rng = np.random.default_rng(21)
    hml = rng.normal(-2.0, 1.5, (2, 64, 64, 18)).astype(np.float32)
    hml[0, :, :, 0] = rng.normal(-5.0, 0.3, (64, 64))           # weak: sigmoid max < 0.2
    hml[0, :, :, 1] = 0.75                                       # constant
    hml[0, :, :, 2] = np.float32(-80.0) + rng.integers(0, 3, (64, 64)).astype(np.float32) * np.float32(1e-5)   # tiny activations
    hml[0, :, :, 3] = rng.choice(np.array([-100.0, -30.0, 30.0, 100.0], np.float32), (64, 64))               # saturated
    hml[1, :, :, 4] = -np.abs(hml[1, :, :, 4])
    hml[1, :, :, 5] = np.abs(hml[1, :, :, 5])
    hml[1, :, :, 6] = rng.choice(np.array([0.0, -0.0], np.float32), (64, 64))
    hml[1, :, :, 7] = np.float32(1.0) + rng.integers(0, 2, (64, 64)).astype(np.float32) * np.float32(1.2e-7)  # 1-ulp range
    det = make_det(weights, synthetic.WORKLOADS["tiny"], 2)
    try:
        kh, seg, mn, mx = oracle.heatmaps(hml)
        gkh, gseg, gmm, gnh = det.heatmaps_normalised(_cuda(hml))
        assert_bits(gkh.cpu().numpy(), kh, "keypoint_heatmaps")
        gmm = gmm.cpu().numpy()
        assert_bits(gmm[..., 0], mn, "min")
        assert_bits(gmm[..., 1], mx, "max")
        want = oracle_normalised(kh, mn, mx)
        got = gnh.cpu().numpy()[..., :17]
        assert np.isnan(want[0, :, :, 1]).all() and np.isnan(got[0, :, :, 1]).all()
        assert not got[0, :, :, 0].any()
        assert_bits(got, want, "normalised map", nan_ok=True)
    finally:
        det.close()


# ------------------------------------------------------------------------------------------- padded crop kernel
def _extra_boxes(rng, B, n):
    special = np.array([[0, 0, 1, 1], [-0.2, -0.1, 0.5, 0.6], [0.5, 0.5, 1.3, 1.2], [0.3, 0.3, 0.3, 0.3], [0.9, 0.9, 0.1, 0.1],
                        [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0]], np.float32)
    boxes = np.concatenate([special, rng.uniform(0, 1, (n, 4)).astype(np.float32)])
    return boxes, rng.integers(0, B, len(boxes)).astype(np.int32)


@pytest.mark.parametrize("key,batch", [("tiny", None), ("c1", None), ("c2", 8), ("c3", 2)])
def test_crop_padded_bit_exact(weights, key, batch):
    """crop_padded_kernel on the device's own normalised map: every crop value == oracle.crop_and_resize of the
    normalised taps (create_pb.py:93-94, 106-109), inside, on and outside the image; the bf16 copy is the RNE rounding."""
    wl = synthetic.WORKLOADS[key]
    B = wl.batch if batch is None else batch
    inp = synthetic.make_inputs(wl, batch=B)
    det = make_det(weights, wl, B)
    try:
        kh, _, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
        _, _, _, nh = det.heatmaps_normalised(_cuda(inp["heatmap_logits"]))
        rng = np.random.default_rng(5)
        gt = np.concatenate(inp["gt_boxes"]).astype(np.float32)
        gi = np.concatenate([np.full(len(g), b, np.int32) for b, g in enumerate(inp["gt_boxes"])])
        if len(gt) > 200:
            pick = rng.choice(len(gt), 200, replace=False)
            gt, gi = gt[pick], gi[pick]
        xb, xi = _extra_boxes(rng, B, 10)
        boxes, ind = np.concatenate([gt, xb]), np.concatenate([gi, xi])
        want = oracle.crop_and_resize(kh, boxes, ind, (56, 36), mn, mx)
        f, b16 = det.crop_padded(nh, _cuda(boxes), _cuda(ind))
        assert_bits(f.cpu().numpy(), want, "padded crops")
        assert torch.equal(b16, f.to(torch.bfloat16))
        assert want.any()
        # more than 640 boxes in one call: the large-call form of the kernel (8 rows per CTA, horizontally interpolated
        # source rows kept in registers across crop rows) -- tall, flat, tiny, flipped and out-of-image boxes included
        many, many_i = _extra_boxes(rng, B, 693)
        many[7:107, 2] = many[7:107, 0] + rng.uniform(0.0, 0.05, 100).astype(np.float32)      # flat: many crop rows per source row
        many[107:207, 2] = np.minimum(many[107:207, 0] + 0.9, 1.2)                            # tall: source rows skipped
        boxes2, ind2 = np.concatenate([boxes, many]), np.concatenate([ind, many_i])
        assert len(boxes2) > 640
        want2 = oracle.crop_and_resize(kh, boxes2, ind2, (56, 36), mn, mx)
        f2, b2 = det.crop_padded(nh, _cuda(boxes2), _cuda(ind2))
        assert_bits(f2.cpu().numpy(), want2, "padded crops, large call")
        assert torch.equal(b2, f2.to(torch.bfloat16))
    finally:
        det.close()


# ------------------------------------------------------------------------------------------- whole path, internal buffers
def _sample_rows(n, k, rng):
    return np.sort(rng.choice(n, min(k, n), replace=False))


def _check_keypoints_all(dev_pos, dev_scores, logits_exact, logits_emul, what, rtol):
    """EVERY keypoint: the device's position must hold an oracle logit within the bf16 tolerance of the oracle's maximum
    (a legitimate near-tie); reports how many differ from the oracle's own argmax and how far away they are."""
    n = logits_exact.shape[0]
    flat = logits_exact.reshape(n, 2016, 17)
    arg_o = flat.argmax(1)                                                           # first index, as tf.argmax
    yx = np.rint(dev_pos * np.array([56, 36], np.float32)).astype(np.int64)
    arg_d = yx[..., 0] * 36 + yx[..., 1]
    assert (arg_d >= 0).all() and (arg_d < 2016).all()
    scale = np.abs(flat).max((1, 2))[:, None]                                        # per person: what the PRN bars are relative to
    top = np.take_along_axis(flat, arg_o[:, None, :], 1)[:, 0]
    at_dev = np.take_along_axis(flat, arg_d[:, None, :], 1)[:, 0]
    short = (top - at_dev) / np.maximum(scale, 1e-30)
    differ = arg_d != arg_o
    dist = np.maximum(np.abs(arg_d // 36 - arg_o // 36), np.abs(arg_d % 36 - arg_o % 36))
    msg = (f"{what}: {int(differ.sum())} of {differ.size} keypoints differ from the fp32 oracle's argmax "
           f"(max shortfall {short.max():.2e} of the channel's logit scale, {int((differ & (dist > 1)).sum())} of them more than "
           f"one grid cell away)")
    print(msg)
    assert short.max() <= rtol, msg
    if logits_emul is not None:
        fe = logits_emul.reshape(n, 2016, 17)
        top_e = fe.max(1)
        at_e = np.take_along_axis(fe, arg_d[:, None, :], 1)[:, 0]
        short_e = (top_e - at_e) / np.maximum(np.abs(fe).max((1, 2))[:, None], 1e-30)
        assert short_e.max() <= 2 * RTOL_BF16_EMUL, f"{what}: shortfall {short_e.max():.2e} against the bf16-operand oracle"
    return int(differ.sum())


def _full_run_case(weights, key, batch, mode, sample):
    wl = synthetic.WORKLOADS[key]
    B = wl.batch if batch is None else batch
    inp = synthetic.make_inputs(wl, batch=B)
    det = make_det(weights, wl, B, prn_mode=mode, prn_modes_allocated=(mode,))
    rng = np.random.default_rng(99)
    try:
        dev_in = [_cuda(inp[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
        # ---- run 1: PRN and decode skipped -> the crops stay in the workspace
        det.debug_skip(16 | 32)
        out = det.run_device(*dev_in)
        torch.cuda.synchronize()
        det.debug_skip(0)
        got = {k: v.cpu().numpy().copy() for k, v in out.items()}
        anc = oracle.anchors(wl.height, wl.width, multipliers=wl.multipliers)
        want = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, wl.score_threshold, wl.iou_threshold,
                             wl.max_detections)
        assert_bits(got["num_boxes"], want["num_boxes"], "num_boxes")
        assert_bits(got["boxes"], want["boxes"], "boxes")
        assert_bits(got["scores"], want["scores"], "scores")
        kh, seg, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
        assert_bits(got["keypoint_heatmaps"], kh, "keypoint_heatmaps")
        assert_bits(got["segmentation_masks"], seg, "segmentation_masks")
        N = int(want["num_boxes"].sum())
        assert int(got["person_offsets"][-1]) == N
        pb = np.concatenate([want["boxes"][b, :want["num_boxes"][b]] for b in range(B)])
        pi = np.concatenate([np.full(want["num_boxes"][b], b, np.int32) for b in range(B)])
        assert_bits(det.debug_fetch("person_box").cpu().numpy().reshape(-1, 4)[:N], pb, "person list: boxes")
        assert_bits(det.debug_fetch("person_image").cpu().numpy()[:N], pi, "person list: box_ind")
        mm = det.debug_fetch("minmax").cpu().numpy().reshape(B, 17, 2)
        assert_bits(mm[..., 0], mn, "min")
        assert_bits(mm[..., 1], mx, "max")
        padded = B * (wl.height // 4) * (wl.width // 4) <= 1500 * B * wl.max_detections
        if padded:
            nh = det.debug_fetch("normalised").cpu().numpy().reshape(B, wl.height // 4, wl.width // 4, 20)
            imgs = _sample_rows(B, 4, rng)                    # whole maps of up to 4 images
            assert_bits(nh[imgs][..., :17], oracle_normalised(kh[imgs], mn[imgs], mx[imgs]), "normalised map (internal)")
        else:
            with pytest.raises(Exception):
                det.debug_fetch("normalised")                 # per-tap path: the map is never built
        rows = _sample_rows(N, sample, rng)
        crops_dev = det.debug_fetch("crops_f32").reshape(-1, 56 * 36 * 17)
        crops_want = oracle.crop_and_resize(kh, pb[rows], pi[rows], (56, 36), mn, mx).reshape(len(rows), -1)
        crops_got = crops_dev[torch.as_tensor(rows).cuda()].cpu().numpy()
        assert_bits(crops_got, crops_want, f"crops of {len(rows)} sampled persons (internal buffer)")
        if mode == "bf16":
            cb = det.debug_fetch("crops_bf16").reshape(-1, 56 * 36 * 17)[torch.as_tensor(rows).cuda()]
            assert torch.equal(cb.cpu(), torch.from_numpy(crops_want).to(torch.bfloat16)), "bf16 crops != RNE(fp32 crops)"
        # ---- run 2: the whole path
        out = det.run_device(*dev_in)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy().copy() for k, v in out.items()}
        logits_dev = det.debug_fetch("logits").reshape(-1, 56 * 36 * 17)[torch.as_tensor(rows).cuda()].cpu().numpy()
        exact = oracle.prn(crops_want, *weights, mode=0)
        tol = RTOL_FP32 if mode == "fp32" else RTOL_BF16
        assert np.abs(logits_dev - exact).max() <= tol * np.abs(exact).max(), "PRN logits vs fp64-accumulating oracle"
        emul = None
        if mode == "bf16":
            emul = oracle.prn(crops_want, *weights, mode=1)
            assert np.abs(logits_dev - emul).max() <= RTOL_BF16_EMUL * np.abs(emul).max(), "PRN logits vs bf16-operand oracle"
        kp, ks = got["keypoint_positions"][:N][rows], got["keypoint_scores"][:N][rows]
        # the decode itself is exact on the device's own logits (create_pb.py:115-142)
        s_o, pos_o, _, _ = oracle.keypoint_decode(logits_dev.reshape(-1, 56, 36, 17))
        assert_bits(kp, pos_o, "keypoint_positions == oracle decode of the device's logits")
        np.testing.assert_allclose(ks, s_o, rtol=RTOL_FP32)
        # ... and legitimate against the oracle's logits for EVERY sampled keypoint
        n_diff = _check_keypoints_all(kp, ks, exact.reshape(-1, 56, 36, 17), None if emul is None else emul.reshape(-1, 56, 36, 17),
                                      f"{key} B={B} {mode}", 2 * (RTOL_FP32 if mode == "fp32" else RTOL_BF16))
        return n_diff, len(rows) * 17
    finally:
        det.close()


@pytest.mark.parametrize("key,batch,mode,sample", [
    ("tiny", None, "bf16", 64), ("tiny", None, "fp32", 64), ("c1", None, "bf16", 64), ("c1", None, "fp32", 16),
    ("c2", 8, "bf16", 64), ("c2", 8, "fp32", 24), ("c3", 32, "bf16", 64), ("c4", 64, "bf16", 64),
    ("c3", 1, "fp32", 24), ("c3", 2, "fp32", 24)])   # fp32 mode with 2 and 3 person groups of the tensor-core PRN
def test_full_run_internal_buffers_and_sampled_rows(weights, key, batch, mode, sample):
    """BASELINE configs[0..3] at their full batch sizes (c4 on ONE GPU takes the per-tap crop path)."""
    _full_run_case(weights, key, batch, mode, sample)


# ------------------------------------------------------------------------------------------- c5: PRN-only, to 100 000 persons
@pytest.mark.parametrize("n", [1000, 10000, 100000])
def test_prn_sweep_sampled_rows(weights, n):
    """BASELINE configs[4]: 1k - 100k person crops through mpn_prn (bf16, in place, the large-batch tcgen05 kernels);
    64 sampled rows against the oracle (1e-2 vs fp64 accumulation, 2e-3 vs the bf16-operand oracle), decode of those rows."""
    from multiposenet_b200 import Detector, DetectorConfig
    det = Detector(weights, DetectorConfig(max_batch=max(1, n // 1000), max_height=128, max_width=128, max_boxes=1000,
                                           prn_mode="bf16", prn_modes_allocated=("bf16",)))
    try:
        g = torch.Generator(device="cuda").manual_seed(20240500 + n)
        D = 56 * 36 * 17
        x = torch.empty((n, D), dtype=torch.float32, device="cuda")
        step = 8192
        for lo in range(0, n, step):              # sparse U(0,1) crops (3 % density), generated on the device
            hi = min(n, lo + step)
            v = torch.rand((hi - lo, D), generator=g, device="cuda")
            m = torch.rand((hi - lo, D), generator=g, device="cuda") < 0.03
            x[lo:hi] = v * m
        rows = _sample_rows(n, 64, np.random.default_rng(n))
        idx = torch.as_tensor(rows).cuda()
        x_rows = x[idx].cpu().numpy()
        out = det.prn(x.view(n, 56, 36, 17), "bf16", inplace=True)
        torch.cuda.synchronize()
        got = out.view(n, D)[idx].cpu().numpy()
        exact = oracle.prn(x_rows, *weights, mode=0)
        emul = oracle.prn(x_rows, *weights, mode=1)
        assert np.abs(got - exact).max() <= RTOL_BF16 * np.abs(exact).max()
        assert np.abs(got - emul).max() <= RTOL_BF16_EMUL * np.abs(emul).max()
        s, pos, arg = det.keypoint_decode(out.view(n, 56, 36, 17))
        s_o, pos_o, arg_o, _ = oracle.keypoint_decode(got.reshape(-1, 56, 36, 17))
        assert_bits(pos[idx].cpu().numpy(), pos_o, "decode of the sampled rows")
        assert_bits(arg[idx].cpu().numpy(), arg_o, "argmax of the sampled rows")
        _check_keypoints_all(pos[idx].cpu().numpy(), s[idx].cpu().numpy(), exact.reshape(-1, 56, 36, 17),
                             emul.reshape(-1, 56, 36, 17), f"c5 n={n}", 2 * RTOL_BF16)
    finally:
        det.close()


# ------------------------------------------------------------------- crop_and_resize inside the single-kernel PRN (CropFuse)
@pytest.mark.parametrize("key,batch,max_boxes", [("tiny", None, None), ("c1", None, None), ("c2", 8, None), ("c2", 3, None),
                                                 ("c3", 2, 128)])
def test_fused_crop_equals_the_crop_kernel_bit_for_bit(weights, monkeypatch, key, batch, max_boxes):
    """MPN_FUSE_CROP=1: calls that fit the single-kernel PRN (capacity <= 256 persons) sample their crops inside it
    (create_pb.py:106-109 fused into the producer of fc1 and the residual of fc2, north_star) and never launch the crop
    kernel.  The crops it leaves in memory (bf16 operand, fp32 residual: skip bit 64 stops the kernel after sampling), the
    flat person list and every output of the whole call must equal the crop-kernel path bit for bit -- that path is the
    one checked against the oracle above."""
    wl = synthetic.WORKLOADS[key]
    B = wl.batch if batch is None else batch
    inp = synthetic.make_inputs(wl, batch=B)
    kw = dict(prn_mode="bf16", prn_modes_allocated=("bf16",))
    if max_boxes is not None:
        kw["max_boxes"] = max_boxes
    res = []
    for fuse in ("0", "1"):
        monkeypatch.setenv("MPN_FUSE_CROP", fuse)
        det = make_det(weights, wl, B, **kw)
        try:
            dev_in = [_cuda(inp[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
            det.debug_skip((64 | 32) if fuse == "1" else (16 | 32))
            out = det.run_device(*dev_in)
            torch.cuda.synchronize()
            det.debug_skip(0)
            N = int(out["person_offsets"].cpu().numpy()[-1])
            r = {"N": N,
                 "crops_f32": det.debug_fetch("crops_f32").reshape(-1, 56 * 36 * 17)[:N].cpu().numpy().copy(),
                 "crops_bf16": det.debug_fetch("crops_bf16").reshape(-1, 56 * 36 * 17)[:N].view(torch.int16).cpu().numpy().copy(),
                 "person_box": det.debug_fetch("person_box").cpu().numpy().reshape(-1, 4)[:N].copy(),
                 "person_image": det.debug_fetch("person_image").cpu().numpy()[:N].copy()}
            for rep in range(2):                       # second call = graph replay
                out = det.run_device(*dev_in)
            torch.cuda.synchronize()
            r["out"] = {k: v.cpu().numpy().copy() for k, v in out.items()}
            r["launches"] = det.launch_count()[0] if hasattr(det, "launch_count") else None
            res.append(r)
        finally:
            det.close()
    a, b = res
    assert a["N"] == b["N"] and a["N"] > 0
    assert B * (max_boxes or wl.max_detections) <= 256, "the case must fit the single-kernel PRN"
    N = a["N"]
    assert_bits(b["person_box"], a["person_box"], "person list: boxes")
    assert np.array_equal(b["person_image"], a["person_image"]), "person list: box_ind"
    assert_bits(b["crops_f32"], a["crops_f32"], "fp32 crops sampled inside the PRN kernel")
    assert np.array_equal(b["crops_bf16"], a["crops_bf16"]), "bf16 crops sampled inside the PRN kernel"
    for k in a["out"]:
        rows = N if k in ("keypoint_scores", "keypoint_positions") else None
        ga, gb = a["out"][k][:rows], b["out"][k][:rows]
        if ga.dtype == np.float32:
            assert_bits(gb, ga, k)
        else:
            assert np.array_equal(gb, ga), k
    if a["launches"] is not None:
        assert b["launches"] == a["launches"] - 1, (a["launches"], b["launches"])      # no crop kernel


def test_prn_red_add_variant_matches_the_fixed_order_reduce(weights, monkeypatch):
    """MPN_PRN_RED_ADD=1 (measured variant, off by default): the K splits of fc1 are added up by the L2 instead of being
    stored and reduced in a fixed order.  Same logits up to the order of 37 fp32 additions (detector/prn.py:20), call after
    call (every reader clears the accumulator), and keypoints that are legitimate maxima of the default path's logits."""
    wl = synthetic.WORKLOADS["c2"]
    inp = synthetic.make_inputs(wl, batch=4)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("MPN_PRN_RED_ADD", flag)
        det = make_det(weights, wl, 4, prn_mode="bf16", prn_modes_allocated=("bf16",))
        try:
            dev_in = [_cuda(inp[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")]
            got = []
            for rep in range(3):
                out = det.run_device(*dev_in)
                torch.cuda.synchronize()
                N = int(out["person_offsets"].cpu().numpy()[-1])
                got.append((det.debug_fetch("logits").reshape(-1, 56 * 36 * 17)[:N].cpu().numpy().copy(),
                            out["keypoint_positions"].cpu().numpy()[:N].copy()))
            outs.append(got)
        finally:
            det.close()
    ref_logits, ref_pos = outs[0][0]
    scale = np.abs(ref_logits).max()
    for rep, (lg, pos) in enumerate(outs[1]):
        assert lg.shape == ref_logits.shape
        err = np.abs(lg - ref_logits).max() / scale
        # (a different order of 37 fp32 additions flips the bf16 rounding of a few y1 values: RTOL_BF16_EMUL's effect)
        assert err < RTOL_BF16_EMUL, f"call {rep}: logits differ by {err:.2e} of the scale"
        flat = ref_logits.reshape(len(lg), 2016, 17)
        yx = np.rint(pos * np.array([56, 36], np.float32)).astype(np.int64)
        at = np.take_along_axis(flat, (yx[..., 0] * 36 + yx[..., 1])[:, None, :], 1)[:, 0]
        assert ((flat.max(1) - at) / scale).max() < 2 * RTOL_BF16_EMUL
