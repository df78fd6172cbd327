"""Person-detector evaluation on the new decode + NMS: average precision @ IoU 0.5 "like in PASCAL VOC" -- the metric
the reference computes during training-time evaluation.

Host-side mirror of the reference's metrics.py: `Evaluator` keeps the accumulate / evaluate / reset cycle of
metrics.py:15-98 (without the tf.metrics plumbing of get_metric_ops, :27-77, which only exists to call these methods
from a TensorFlow graph), `evaluate_detector` restates metrics.py:125-185 and its helpers :188-254 operation for
operation (Python float arithmetic, the same greedy matching, the same tie behaviour of list.sort).
Inputs are what `Detector` returns: boxes (ymin, xmin, ymax, xmax), scores, num_boxes.
"""
import numpy as np

METRIC_NAMES = ("AP", "precision", "recall", "mean_iou_for_TP", "best_threshold", "total_FP", "total_FN")   # metrics.py:63-66


def _box(box, image_name=None, score=None):
    """metrics.py:101-122"""
    ymin, xmin, ymax, xmax = box
    d = {"ymin": ymin, "xmin": xmin, "ymax": ymax, "xmax": xmax}
    if score is not None and image_name is not None:
        d["image_name"] = image_name
        d["confidence"] = score
    else:
        d["is_matched"] = False
    return d


def compute_iou(box1, box2):
    """metrics.py:212-225"""
    w = min(box1["xmax"], box2["xmax"]) - max(box1["xmin"], box2["xmin"])
    if w > 0:
        h = min(box1["ymax"], box2["ymax"]) - max(box1["ymin"], box2["ymin"])
        if h > 0:
            intersection = w * h
            w1 = box1["xmax"] - box1["xmin"]
            h1 = box1["ymax"] - box1["ymin"]
            w2 = box2["xmax"] - box2["xmin"]
            h2 = box2["ymax"] - box2["ymin"]
            union = (w1 * h1 + w2 * h2) - intersection
            return float(intersection) / float(union)
    return 0.0


def match(detection, groundtruth_boxes):
    """metrics.py:228-244: index and IoU of the best ground-truth box (matched or not), (-1, 0.0) if none overlaps."""
    best_i, max_iou = -1, 0.0
    for i, box in enumerate(groundtruth_boxes):
        iou = compute_iou(detection, box)
        if iou > max_iou:
            best_i, max_iou = i, iou
    return best_i, max_iou


def compute_ap(precision, recall):
    """metrics.py:247-254: sum of precision * recall increments (recall is non-decreasing)."""
    previous, ap = 0.0, 0.0
    for p, r in zip(precision, recall):
        ap += p * (r - previous)
        previous = r
    return ap


def compute_best_threshold(precision, recall, confidences):
    """metrics.py:188-209"""
    if len(confidences) == 0:
        return 0.0, 0.0, 0.0
    precision, recall, confidences = np.array(precision), np.array(recall), np.array(confidences)
    diff = np.abs(precision - recall)
    best_i = np.argmax(precision * recall * (1.0 - diff))
    return confidences[best_i], precision[best_i], recall[best_i]


def evaluate_detector(groundtruth, detections, iou_threshold=0.5):
    """metrics.py:125-185.  groundtruth: image -> list of boxes; detections: list of boxes (sorted in place)."""
    num_groundtruth_boxes = max(sum(len(b) for b in groundtruth.values()), 1)
    detections.sort(key=lambda box: box["confidence"], reverse=True)
    num_correct, num_detections, mean_iou = 0, 0, 0.0
    precision, recall = [0.0] * len(detections), [0.0] * len(detections)
    confidences = [box["confidence"] for box in detections]
    for k, detection in enumerate(detections):
        num_detections += 1
        gt = groundtruth.get(detection["image_name"], [])
        best_i, max_iou = match(detection, gt)
        if best_i >= 0 and max_iou >= iou_threshold:
            if not gt[best_i]["is_matched"]:
                gt[best_i]["is_matched"] = True
                num_correct += 1
                mean_iou += max_iou
        precision[k] = num_correct / num_detections
        recall[k] = num_correct / num_groundtruth_boxes
    ap = compute_ap(precision, recall)
    best_threshold, best_precision, best_recall = compute_best_threshold(precision, recall, confidences)
    mean_iou /= max(num_correct, 1)
    return {"AP": ap, "precision": best_precision, "recall": best_recall, "best_threshold": best_threshold,
            "mean_iou_for_TP": mean_iou, "total_FP": num_detections - num_correct,
            "total_FN": num_groundtruth_boxes - num_correct}


class Evaluator:
    """metrics.py:15-98 without the TensorFlow op plumbing: add images, evaluate(), read .metrics."""

    def __init__(self):
        self.initialize()

    def initialize(self):
        self.detections = []
        self.groundtruth = {}
        self.unique_image_id = 0

    def add_groundtruth(self, image_name, boxes):
        for box in boxes:
            self.groundtruth.setdefault(image_name, []).append(_box(box))

    def add_detections(self, image_name, boxes, scores):
        for box, score in zip(boxes, scores):
            self.detections.append(_box(box, image_name, score))

    def add_image(self, gt_boxes, boxes, scores, num_boxes=None):
        """One image, as update_op_func does (metrics.py:39-43): boxes/scores are cut to num_boxes (:45-50)."""
        name = "{}".format(self.unique_image_id)
        self.unique_image_id += 1
        n = len(scores) if num_boxes is None else int(num_boxes)
        self.add_groundtruth(name, gt_boxes)
        self.add_detections(name, boxes[:n], scores[:n])

    def add_batch(self, gt_boxes_per_image, outputs):
        """A batched Detector result (boxes [B,max_det,4], scores [B,max_det], num_boxes [B])."""
        for b, gt in enumerate(gt_boxes_per_image):
            self.add_image(gt, outputs["boxes"][b], outputs["scores"][b], outputs["num_boxes"][b])

    def evaluate(self, iou_threshold=0.5):
        self.metrics = evaluate_detector(self.groundtruth, self.detections, iou_threshold)
        return self.metrics
