"""GPU parity at BASELINE.json's full sizes through size-independent properties (the oracle is too slow, or its fp64 PRN
too slow, to be run at these sizes inside a test): batch independence, permutation equivariance, NMS post-conditions,
idempotence of NMS on its own output, stage composition == whole path, determinism, host path == device path."""
import numpy as np
import pytest
import torch

import oracle
from multiposenet_b200 import synthetic

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _run(det, inp, **kw):
    out = det.run_device(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), _cuda(inp["heatmap_logits"]), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in out.items()}


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    same = (a.view(np.uint32) == b.view(np.uint32)) if a.dtype == np.float32 else (a == b)
    assert same.all(), f"{what}: {int((~same).sum())} of {same.size} differ"


def _iou_matrix(b):
    ymin = np.maximum(b[:, None, 0], b[None, :, 0]); xmin = np.maximum(b[:, None, 1], b[None, :, 1])
    ymax = np.minimum(b[:, None, 2], b[None, :, 2]); xmax = np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.clip(ymax - ymin, 0, None) * np.clip(xmax - xmin, 0, None)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / np.maximum(area[:, None] + area[None, :] - inter, 1e-30)


@pytest.fixture(scope="module")
def weights():
    return synthetic.make_prn_weights(bias_std=0.01)


@pytest.fixture(scope="module")
def det_c2(weights):
    """BASELINE configs[1]: 640x640, batch 8, 9 anchors / location, thr 0.3, IoU 0.5, 25 boxes."""
    from multiposenet_b200 import Detector, DetectorConfig
    wl = synthetic.WORKLOADS["c2"]
    d = Detector(weights, DetectorConfig(max_batch=8, max_height=640, max_width=640, max_boxes=25, score_threshold=0.3,
                                         iou_threshold=0.5, scale_multipliers=wl.multipliers, prn_mode="bf16"))
    yield d
    d.close()


def test_c2_full_size_properties(det_c2, weights):
    wl = synthetic.WORKLOADS["c2"]
    inp = synthetic.make_inputs(wl)
    full = _run(det_c2, inp)
    B = wl.batch
    N = int(full["person_offsets"][-1])
    assert np.array_equal(np.diff(full["person_offsets"]), full["num_boxes"]) and N >= 40
    # (a) determinism: a second run gives the same bits
    again = _run(det_c2, inp)
    for k in full:
        rows = N if k in ("keypoint_scores", "keypoint_positions") else None
        _eq(again[k][:rows], full[k][:rows], f"determinism {k}")
    # (b) NMS post-conditions per image: scores descending, above threshold, kept boxes pairwise IoU <= thr, zero padding
    for b in range(B):
        n = int(full["num_boxes"][b])
        sc, bx = full["scores"][b], full["boxes"][b]
        assert (np.diff(sc[:n]) <= 0).all() and (sc[:n] > wl.score_threshold).all()
        assert not sc[n:].any() and not bx[n:].any()
        assert (bx[:n] >= 0).all() and (bx[:n] <= 1).all()
        iou = _iou_matrix(bx[:n].astype(np.float64))
        np.fill_diagonal(iou, 0)
        assert iou.max() <= wl.iou_threshold + 1e-6
    # (c) batch independence: every image alone gives its slice of the batch result
    for b in (0, 3, 7):
        one = {k: inp[k][b:b + 1] for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
        r = _run(det_c2, one)
        _eq(r["boxes"][0], full["boxes"][b], "boxes of a single image")
        _eq(r["scores"][0], full["scores"][b], "scores of a single image")
        _eq(r["keypoint_heatmaps"][0], full["keypoint_heatmaps"][b], "heatmaps of a single image")
        lo, hi = full["person_offsets"][b], full["person_offsets"][b + 1]
        _eq(r["keypoint_positions"][:hi - lo], full["keypoint_positions"][lo:hi], "keypoints of a single image")
        _eq(r["keypoint_scores"][:hi - lo], full["keypoint_scores"][lo:hi], "keypoint scores of a single image")
    # (d) permutation equivariance over the batch
    perm = np.array([5, 2, 7, 0, 3, 6, 1, 4])
    pin = {k: np.ascontiguousarray(inp[k][perm]) for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
    pr = _run(det_c2, pin)
    _eq(pr["boxes"], full["boxes"][perm], "permuted boxes")
    _eq(pr["num_boxes"], full["num_boxes"][perm], "permuted num_boxes")
    for i, b in enumerate(perm):
        lo, hi = full["person_offsets"][b], full["person_offsets"][b + 1]
        plo = pr["person_offsets"][i]
        _eq(pr["keypoint_positions"][plo:plo + hi - lo], full["keypoint_positions"][lo:hi], "permuted keypoints")
    # (e) detection + heatmap stages against the oracle at full size (these oracle stages take milliseconds)
    anc = oracle.anchors(wl.height, wl.width, multipliers=wl.multipliers)
    want = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, wl.score_threshold, wl.iou_threshold, 25)
    _eq(full["boxes"], want["boxes"], "boxes vs oracle")
    _eq(full["scores"], want["scores"], "scores vs oracle")
    _eq(full["num_boxes"], want["num_boxes"], "num_boxes vs oracle")
    kh, seg, _, _ = oracle.heatmaps(inp["heatmap_logits"])
    _eq(full["keypoint_heatmaps"], kh, "keypoint_heatmaps vs oracle")
    _eq(full["segmentation_masks"], seg, "segmentation_masks vs oracle")
    # (f) keypoints: positions on the 56 x 36 grid, scores are softmax maxima
    kp, ks = full["keypoint_positions"][:N], full["keypoint_scores"][:N]
    gy, gx = kp[..., 0].astype(np.float64) * 56, kp[..., 1].astype(np.float64) * 36
    assert np.abs(gy - np.round(gy)).max() < 1e-4 and np.abs(gx - np.round(gx)).max() < 1e-4
    assert (kp >= 0).all() and (kp < 1).all() and (ks > 1 / 2016 - 1e-7).all() and (ks <= 1).all()
    # (g) the host path (three calls in flight, codes gathered from pinned memory) gives the same bits
    pinned = {k: torch.from_numpy(inp[k]).pin_memory() for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
    t, bufs = det_c2.submit_host(pinned["encoded_boxes"], pinned["class_logits"], pinned["heatmap_logits"])
    det_c2.wait(t)
    for k in full:
        rows = N if k in ("keypoint_scores", "keypoint_positions") else None
        _eq(bufs[k].numpy()[:rows], full[k][:rows], f"host path {k}")


def test_nms_is_idempotent_on_its_own_output(det_c2):
    """Feeding the kept boxes back (as anchors with zero codes is not expressible, so: re-running detection with the
    kept anchors' logits only) keeps every one of them: no kept box suppresses another."""
    wl = synthetic.WORKLOADS["c2"]
    inp = synthetic.make_inputs(wl, replicate=3)
    a = det_c2.detect(_cuda(inp["encoded_boxes"]), _cuda(inp["class_logits"]), (wl.height, wl.width))
    sel = a["sel_anchor"].cpu().numpy()
    cls2 = np.full_like(inp["class_logits"], -20.0)
    for b in range(wl.batch):
        keep = sel[b][sel[b] >= 0]
        cls2[b, keep] = inp["class_logits"][b, keep]
    b2 = det_c2.detect(_cuda(inp["encoded_boxes"]), _cuda(cls2), (wl.height, wl.width))
    for k in ("boxes", "scores", "num_boxes", "sel_anchor"):
        _eq(b2[k].cpu().numpy(), a[k].cpu().numpy(), f"idempotence {k}")


def test_crowded_scene_and_large_person_counts(weights):
    """BASELINE configs[2]-style: 100+ persons per image, 128 boxes kept, several thousand candidates per image (the
    chunked NMS path) and > 256 persons per call (the tiled tcgen05 GEMM path instead of the single-kernel PRN)."""
    from multiposenet_b200 import Detector, DetectorConfig
    wl = synthetic.WORKLOADS["c3"]
    B = 4
    inp = synthetic.make_inputs(wl, batch=B)
    det = Detector(weights, DetectorConfig(max_batch=B, max_height=640, max_width=640, max_boxes=128, score_threshold=0.3,
                                           iou_threshold=0.5, scale_multipliers=wl.multipliers, prn_mode="bf16"))
    try:
        full = _run(det, inp)
        N = int(full["person_offsets"][-1])
        assert N > 256 and full["num_boxes"].max() > 80
        anc = oracle.anchors(wl.height, wl.width, multipliers=wl.multipliers)
        want = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, 0.3, 0.5, 128)
        assert want["n_conf"].max() > 1024            # more candidates than one rank-sort / one NMS chunk
        _eq(full["boxes"], want["boxes"], "crowded boxes vs oracle")
        _eq(full["num_boxes"], want["num_boxes"], "crowded num_boxes vs oracle")
        # one image alone (<= 128 persons: single-kernel PRN) == its slice of the batch (tiled GEMM PRN) on every keypoint
        # whose decision margin is clear of bf16 accumulation-order noise
        one = {k: inp[k][1:2] for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
        r = _run(det, one)
        lo, hi = full["person_offsets"][1], full["person_offsets"][2]
        same = (r["keypoint_positions"][:hi - lo] == full["keypoint_positions"][lo:hi]).all(-1)
        assert same.mean() > 0.97
        np.testing.assert_allclose(r["keypoint_scores"][:hi - lo][same], full["keypoint_scores"][lo:hi][same], rtol=2e-2)
    finally:
        det.close()


def test_1024_image_batch_of_8(weights):
    """BASELINE configs[3] per-GPU shard at 8 GPUs: 1024x1024, 8 images."""
    from multiposenet_b200 import Detector, DetectorConfig
    wl = synthetic.WORKLOADS["c4"]
    inp = synthetic.make_inputs(wl, batch=8)
    det = Detector(weights, DetectorConfig(max_batch=8, max_height=1024, max_width=1024, max_boxes=25,
                                           score_threshold=0.3, iou_threshold=0.5, prn_mode="bf16"))
    try:
        assert det.num_anchors(1024, 1024) == 130944
        full = _run(det, inp)
        anc = oracle.anchors(1024, 1024)
        want = oracle.detect(inp["class_logits"], inp["encoded_boxes"], anc, 0.3, 0.5, 25)
        _eq(full["boxes"], want["boxes"], "1024 boxes vs oracle")
        _eq(full["scores"], want["scores"], "1024 scores vs oracle")
        kh, seg, _, _ = oracle.heatmaps(inp["heatmap_logits"])
        _eq(full["keypoint_heatmaps"], kh, "1024 heatmaps vs oracle")
        one = {k: inp[k][5:6] for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
        r = _run(det, one)
        lo, hi = full["person_offsets"][5], full["person_offsets"][6]
        _eq(r["keypoint_positions"][:hi - lo], full["keypoint_positions"][lo:hi], "1024 keypoints of a single image")
    finally:
        det.close()


def test_lanes_interleaved_calls_reproduce_the_lone_detector(det_c2, weights):
    """DetectorLanes: batches fed round-robin to 3 handles on 3 streams, many calls in flight at once (cooperative PRN
    kernels of different handles queue behind each other), every result bit-identical to the same batch through one
    handle alone."""
    from multiposenet_b200 import DetectorConfig, DetectorLanes
    wl = synthetic.WORKLOADS["c2"]
    sets = [synthetic.make_inputs(wl, replicate=r) for r in range(4)]
    alone = [_run(det_c2, s) for s in sets]
    dev = [{k: _cuda(s[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")} for s in sets]
    lanes = DetectorLanes(weights, DetectorConfig(max_batch=8, max_height=640, max_width=640, max_boxes=25,
                                                  score_threshold=0.3, iou_threshold=0.5,
                                                  scale_multipliers=wl.multipliers, prn_mode="bf16"), lanes=3)
    try:
        names = ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions", "person_offsets",
                 "keypoint_heatmaps", "segmentation_masks")
        lanes.fork()
        for rnd in range(6):                       # 24 calls queued without a host synchronisation
            kept = []
            for i, d in enumerate(dev):
                lane, out = lanes.submit(d["encoded_boxes"], d["class_logits"], d["heatmap_logits"])
                if rnd == 5:
                    with torch.cuda.stream(lanes.streams[lane]):      # snapshot before the lane reuses its buffers
                        kept.append((i, {k: out[k].clone() for k in names}))
        lanes.join()
        torch.cuda.synchronize()
        assert len(kept) == len(sets)
        for i, snap in kept:
            n = int(alone[i]["person_offsets"][-1])
            for k in names:
                a, b = snap[k].cpu().numpy(), alone[i][k]
                if k in ("keypoint_scores", "keypoint_positions"):
                    a, b = a[:n], b[:n]            # rows past the person count are stale by contract
                _eq(a, b, f"set {i}: {k}")
    finally:
        lanes.close()


def test_lanes_stress_many_queued_calls_mixed_person_counts(weights):
    """Robustness of the cooperative single-kernel PRN (hand-rolled grid barrier on a never-reset counter), programmatic
    dependent launch and graph replay under load: 4 lanes x 5 000 calls queued without a host synchronisation, the calls
    cycling through batches with 0 persons (every grid exits on the device-side count), ~120 persons (single-kernel PRN)
    and > 256 persons (large-batch kernels).  Sampled results must be the bits of a lone handle."""
    from multiposenet_b200 import Detector, DetectorConfig, DetectorLanes
    wl = synthetic.WORKLOADS["c3"]
    cfg = DetectorConfig(max_batch=3, max_height=640, max_width=640, max_boxes=128, score_threshold=0.3, iou_threshold=0.5,
                         scale_multipliers=wl.multipliers, prn_mode="bf16", prn_modes_allocated=("bf16",))
    crowd = synthetic.make_inputs(wl, batch=3)
    one = {k: np.ascontiguousarray(crowd[k][:1]) for k in ("class_logits", "encoded_boxes", "heatmap_logits")}
    none = dict(crowd, class_logits=np.full_like(crowd["class_logits"], -9.0))
    kinds = [none, one, crowd]
    names = ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions", "person_offsets")
    lone = Detector(weights, cfg)
    try:
        alone = [_run(lone, k) for k in kinds]
    finally:
        lone.close()
    counts = [int(a["person_offsets"][-1]) for a in alone]
    assert counts[0] == 0 and 0 < counts[1] <= 256 < counts[2]
    dev = [{k: _cuda(s[k]) for k in ("encoded_boxes", "class_logits", "heatmap_logits")} for s in kinds]
    n_lanes, per_lane = 4, 5000
    lanes = DetectorLanes(weights, cfg, lanes=n_lanes)
    try:
        lanes.fork()
        snaps = []
        for i in range(n_lanes * per_lane):
            kind = (i // n_lanes + i % n_lanes) % 3          # every lane sees every kind, neighbours differ
            d = dev[kind]
            lane, out = lanes.submit(d["encoded_boxes"], d["class_logits"], d["heatmap_logits"])
            if i % 1777 == 0 or i >= n_lanes * (per_lane - 1):
                with torch.cuda.stream(lanes.streams[lane]):
                    snaps.append((kind, {k: out[k].clone() for k in names}))
        lanes.join()
        torch.cuda.synchronize()
        assert len(snaps) >= 12
        for kind, snap in snaps:
            n = counts[kind]
            for k in names:
                a, b = snap[k].cpu().numpy(), alone[kind][k]
                if k in ("keypoint_scores", "keypoint_positions"):
                    a, b = a[:n], b[:n]
                _eq(a, b, f"kind {kind}: {k}")
    finally:
        lanes.close()
