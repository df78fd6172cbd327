"""Result consumers: what the reference does with a `Detector` result on the host.

  * keypoints_to_image: box-relative keypoint_positions (y / 56, x / 36) -> image pixel (x, y, visible), the arithmetic of
    the notebook's draw_everything (inference/predict.ipynb, cell 12: boxes scaled by [height, width, height, width],
    keypoints[:, [1, 0]] * [box width, box height] + [xmin, ymin], visibility 1).
  * EDGES / skeleton_segments: the skeleton of inference/utils.py:19-26 and the "both visible" rule of draw_pose (:55-81).
  * to_coco: the same persons as COCO keypoint-detection results (the format the official evaluation scripts, which the
    reference defers to in metrics.py:6-10, consume).
No drawing: the reference's PIL code is visualisation only.
"""
import numpy as np

EDGES = [(0, 1), (0, 2), (1, 3), (2, 4), (5, 7), (7, 9), (6, 8), (8, 10), (11, 13), (13, 15), (12, 14), (14, 16),
         (3, 5), (4, 6), (5, 11), (6, 12)]                                  # inference/utils.py:19-26
KEYPOINT_NAMES = ("nose", "left eye", "right eye", "left ear", "right ear", "left shoulder", "right shoulder",
                  "left elbow", "right elbow", "left wrist", "right wrist", "left hip", "right hip", "left knee",
                  "right knee", "left ankle", "right ankle")               # inference/utils.py:5-16


def keypoints_to_image(boxes, keypoint_positions, image_hw):
    """boxes [N,4] normalised (ymin,xmin,ymax,xmax), keypoint_positions [N,17,2] box-relative (y,x) ->
    (pixel boxes [N,4] float64, keypoints [N,17,3] float32 rows (x, y, visible))."""
    height, width = image_hw
    scaler = np.array([height, width, height, width])
    pboxes = scaler * np.asarray(boxes)                                     # cell 12: boxes = scaler * outputs['boxes']
    out = np.zeros((len(pboxes), 17, 3), np.float32)
    for i, box in enumerate(pboxes):
        ymin, xmin, ymax, xmax = box
        kp = np.asarray(keypoint_positions[i])[:, [1, 0]].copy()            # (x, y)
        kp *= np.array([xmax - xmin, ymax - ymin])
        kp += np.array([xmin, ymin])
        out[i] = np.concatenate([kp, np.ones([17, 1], dtype=np.float32)], axis=1)
    return pboxes, out


def skeleton_segments(keypoints):
    """keypoints [17,3] (x, y, visible) -> list of ((x1, y1), (x2, y2)) for the EDGES whose both ends are visible."""
    segs = []
    for p, q in EDGES:
        x1, y1, v1 = keypoints[p]
        x2, y2, v2 = keypoints[q]
        if v1 > 0 and v2 > 0:
            segs.append(((float(x1), float(y1)), (float(x2), float(y2))))
    return segs


def to_coco(outputs, image_id, image_hw, category_id=1):
    """A single-image Detector result -> list of COCO keypoint results
    {image_id, category_id, bbox [x, y, w, h], score, keypoints [x1, y1, v1, ...], keypoint_scores}."""
    pboxes, kps = keypoints_to_image(outputs["boxes"], outputs["keypoint_positions"], image_hw)
    res = []
    for i, (ymin, xmin, ymax, xmax) in enumerate(pboxes):
        res.append({"image_id": image_id, "category_id": category_id,
                    "bbox": [float(xmin), float(ymin), float(xmax - xmin), float(ymax - ymin)],
                    "score": float(outputs["scores"][i]),
                    "keypoints": [float(v) for v in kps[i].reshape(-1)],
                    "keypoint_scores": [float(v) for v in outputs["keypoint_scores"][i]]})
    return res
