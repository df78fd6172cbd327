"""Person-detector evaluation on the new decode + NMS: average precision at one IoU
threshold, the operating point that balances precision and recall, and the false-positive / false-negative totals --
the numbers the reference's training-time evaluation reports (metrics.py:63-66, computed there by metrics.py:125-254).

Array-based: an image is three arrays (ground-truth boxes [G, 4], detected boxes [K, 4], confidences [K]); evaluation is
  1. one vectorised IoU matrix [K, G] per image,
  2. per image, every detection points at its best-overlapping ground-truth box, and for every ground-truth box the most
     confident detection pointing at it with enough overlap is the true positive (`np.unique` on the sorted pointers: no
     loop over detections) -- which is what a greedy pass in descending confidence with "a box is matched once" computes,
     because matches never cross images,
  3. one global stable sort by confidence and running sums for precision / recall / AP.
Outputs are bit-identical to the reference's Python-float implementation on the goldens of tests/golden/metrics.npz
(intermediate products are formed in the boxes' own dtype and the ratio in float64, running sums are sequential).
Inputs are what `Detector` returns: boxes (ymin, xmin, ymax, xmax), scores, num_boxes.
"""
import numpy as np

METRIC_NAMES = ("AP", "precision", "recall", "mean_iou_for_TP", "best_threshold", "total_FP", "total_FN")


def pairwise_iou(dets, gts):
    """IoU matrix [K, G] (float64) of two box arrays (ymin, xmin, ymax, xmax).  Boxes that do not overlap with positive
    width AND height get exactly 0.  Widths, heights, areas and the union are formed in the input dtype, the final ratio in
    float64."""
    dets, gts = np.asarray(dets), np.asarray(gts)
    if dets.shape[0] == 0 or gts.shape[0] == 0:
        return np.zeros((dets.shape[0], gts.shape[0]), np.float64)
    d, g = dets[:, None, :], gts[None, :, :]
    w = np.minimum(d[..., 3], g[..., 3]) - np.maximum(d[..., 1], g[..., 1])
    h = np.minimum(d[..., 2], g[..., 2]) - np.maximum(d[..., 0], g[..., 0])
    overlap = (w > 0) & (h > 0)
    inter = w * h
    area_d = (dets[:, 3] - dets[:, 1]) * (dets[:, 2] - dets[:, 0])
    area_g = (gts[:, 3] - gts[:, 1]) * (gts[:, 2] - gts[:, 0])
    union = (area_d[:, None] + area_g[None, :]) - inter
    out = np.zeros(inter.shape, np.float64)
    np.divide(inter.astype(np.float64), union.astype(np.float64), out=out, where=overlap)
    return out


def match_image(dets, scores, gts, iou_threshold):
    """True-positive flags and matched IoUs of one image's detections.

    Every detection points at the ground-truth box it overlaps most (first one on ties; none if it overlaps nothing).  A
    ground-truth box is claimed by the most confident detection that points at it with IoU >= iou_threshold (earlier
    detection on equal confidence); every other detection is a false positive -- also one whose best box is already
    claimed, even if it overlaps a second box well enough.
    Returns (is_tp [K] bool, iou_of_match [K] float64)."""
    K = len(scores)
    tp = np.zeros(K, bool)
    iou = pairwise_iou(dets, gts)
    if K == 0 or iou.shape[1] == 0:
        return tp, np.zeros(K, np.float64)
    pointer = iou.argmax(axis=1)
    best = iou[np.arange(K), pointer]
    eligible = (best > 0) & (best >= iou_threshold)
    by_confidence = np.argsort(-np.asarray(scores), kind="stable")
    ranked = by_confidence[eligible[by_confidence]]              # eligible detections, most confident first
    _, first = np.unique(pointer[ranked], return_index=True)     # first claim of every ground-truth box
    tp[ranked[first]] = True
    return tp, best


def average_precision(is_tp, num_groundtruth):
    """(precision [n], recall [n], AP) along a confidence-ranked list of true-positive flags; AP is the area under the
    precision-recall staircase, sum of precision_k * (recall_k - recall_{k-1})."""
    is_tp = np.asarray(is_tp, bool)
    n = is_tp.size
    if n == 0:
        return np.zeros(0), np.zeros(0), 0.0
    hits = np.cumsum(is_tp)
    precision = hits / np.arange(1, n + 1)
    recall = hits / num_groundtruth
    steps = precision * np.diff(recall, prepend=0.0)
    return precision, recall, float(np.cumsum(steps)[-1])        # running (left-to-right) sum


class DetectionEvaluator:
    """Accumulates images, then `evaluate()`.  One instance per evaluation run; `reset()` starts over."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._gt, self._det, self._score = [], [], []

    def __len__(self):
        return len(self._gt)

    def add_image(self, gt_boxes, boxes, scores, num_boxes=None):
        """One image: ground-truth boxes [G, 4], detected boxes [K', 4] and confidences [K'] of which the first
        `num_boxes` rows are real (a Detector result is zero-padded to max_boxes)."""
        n = len(scores) if num_boxes is None else int(num_boxes)
        self._gt.append(np.asarray(gt_boxes).reshape(-1, 4))
        self._det.append(np.asarray(boxes).reshape(-1, 4)[:n])
        self._score.append(np.asarray(scores).reshape(-1)[:n])

    def add_batch(self, gt_boxes_per_image, outputs):
        """A batched Detector result (boxes [B, max_det, 4], scores [B, max_det], num_boxes [B])."""
        for b, gt in enumerate(gt_boxes_per_image):
            self.add_image(gt, outputs["boxes"][b], outputs["scores"][b], outputs["num_boxes"][b])

    def evaluate(self, iou_threshold=0.5):
        num_gt = max(sum(len(g) for g in self._gt), 1)
        flags, ious, confs = [np.zeros(0, bool)], [np.zeros(0)], [np.zeros(0, np.float32)]
        for gt, det, score in zip(self._gt, self._det, self._score):
            tp, best = match_image(det, score, gt, iou_threshold)
            flags.append(tp); ious.append(best); confs.append(score)
        flags, ious, confs = np.concatenate(flags), np.concatenate(ious), np.concatenate(confs)
        rank = np.argsort(-confs, kind="stable")                 # images in insertion order break confidence ties
        flags, ious, confs = flags[rank], ious[rank], confs[rank]
        precision, recall, ap = average_precision(flags, num_gt)
        n_tp = int(flags.sum())
        matched = ious[flags]
        mean_iou = (float(np.cumsum(matched)[-1]) if n_tp else 0.0) / max(n_tp, 1)
        if flags.size:
            # operating point: the confidence at which precision * recall * (1 - |precision - recall|) peaks
            k = int(np.argmax(precision * recall * (1.0 - np.abs(precision - recall))))
            point = (confs[k], precision[k], recall[k])
        else:
            point = (0.0, 0.0, 0.0)
        self.metrics = {"AP": ap, "precision": point[1], "recall": point[2], "best_threshold": point[0],
                        "mean_iou_for_TP": mean_iou, "total_FP": int(flags.size) - n_tp, "total_FN": num_gt - n_tp}
        return self.metrics
