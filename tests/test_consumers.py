"""the host-side consumers of a Detector result.  The detection metric is pinned by
golden vectors produced by the reference's own metrics.py (tests/golden/make_metrics_golden.py)."""
import os

import numpy as np

from multiposenet_b200 import metrics, results

HERE = os.path.dirname(os.path.abspath(__file__))


def test_evaluator_reproduces_reference_metrics_golden():
    g = np.load(os.path.join(HERE, "golden", "metrics.npz"))
    assert int(g["n"]) >= 6
    for k in range(int(g["n"])):
        ev = metrics.DetectionEvaluator()
        for i in range(int(g[f"n_images_{k}"])):
            ev.add_image(g[f"gt_{k}_{i}"], g[f"det_{k}_{i}"], g[f"score_{k}_{i}"])
        m = ev.evaluate(0.5)
        got = np.array([float(m[name]) for name in metrics.METRIC_NAMES], np.float64)
        assert np.array_equal(got, g[f"metrics_{k}"]), (k, got, g[f"metrics_{k}"])


def test_metric_known_answers():
    ev = metrics.DetectionEvaluator()
    gt = np.array([[0.1, 0.1, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9]], np.float32)
    ev.add_image(gt, gt.copy(), np.array([0.9, 0.8], np.float32))                 # perfect detector
    m = ev.evaluate()
    assert m["AP"] == 1.0 and m["total_FP"] == 0 and m["total_FN"] == 0 and abs(m["mean_iou_for_TP"] - 1.0) < 1e-6
    ev.reset()
    ev.add_image(gt, np.array([[0.1, 0.1, 0.5, 0.5], [0.1, 0.1, 0.5, 0.5]], np.float32), np.array([0.9, 0.8], np.float32))
    m = ev.evaluate()                                                            # the duplicate is a false positive
    assert m["total_FP"] == 1 and m["total_FN"] == 1 and m["AP"] == 0.5
    assert metrics.DetectionEvaluator().evaluate()["AP"] == 0.0                 # nothing at all (metrics.py:140, :199-200)
    # a detection whose best box is already claimed is a false positive even if it overlaps a second box well enough
    two = np.array([[0.0, 0.0, 0.5, 0.5], [0.0, 0.1, 0.5, 0.6]], np.float32)
    ev.reset()
    ev.add_image(two, np.array([[0.0, 0.0, 0.5, 0.5], [0.0, 0.04, 0.5, 0.54]], np.float32), np.array([0.9, 0.8], np.float32))
    m = ev.evaluate()
    assert m["total_FP"] == 1 and m["total_FN"] == 1
    iou = metrics.pairwise_iou(two, two)
    assert iou[0, 0] == 1.0 and abs(iou[0, 1] - 0.4 / 0.6) < 1e-6 and metrics.pairwise_iou(two[:0], two).shape == (0, 2)
    # num_boxes cuts the padded rows (metrics.py:45-50)
    ev.reset()
    ev.add_batch([gt], {"boxes": np.concatenate([gt, np.zeros((3, 4), np.float32)])[None],
                        "scores": np.array([[0.9, 0.8, 0, 0, 0]], np.float32), "num_boxes": np.array([2])})
    assert ev.evaluate()["total_FP"] == 0


def test_keypoints_to_image_follows_the_notebook_arithmetic():
    boxes = np.array([[0.25, 0.125, 0.75, 0.625]], np.float32)
    pos = np.zeros((1, 17, 2), np.float32)
    pos[0, 3] = (28 / 56, 9 / 36)                       # (y, x) relative to the box
    pboxes, kps = results.keypoints_to_image(boxes, pos, (640, 512))
    assert np.allclose(pboxes, [[160.0, 64.0, 480.0, 320.0]])
    assert np.allclose(kps[0, 3], [64 + 0.25 * 256, 160 + 0.5 * 320, 1.0])      # (x, y, visible)
    assert np.allclose(kps[0, 0], [64.0, 160.0, 1.0])
    segs = results.skeleton_segments(kps[0])
    assert len(segs) == len(results.EDGES) == 16
    hidden = kps[0].copy(); hidden[5, 2] = 0                                    # left shoulder not visible
    assert len(results.skeleton_segments(hidden)) == 16 - 3                     # edges (5,7), (3,5), (5,11)
    coco = results.to_coco({"boxes": boxes, "scores": np.array([0.9], np.float32), "keypoint_positions": pos,
                            "keypoint_scores": np.full((1, 17), 0.5, np.float32)}, 7, (640, 512))
    assert coco[0]["image_id"] == 7 and coco[0]["bbox"] == [64.0, 160.0, 256.0, 320.0] and len(coco[0]["keypoints"]) == 51
