"""Generates tests/golden/graph_*.npz by EXECUTING THE REFERENCE'S OWN CODE for the post-backbone path under a
numpy-backed `tensorflow.compat.v1` stand-in (tests/golden/tf_numpy_shim.py).  Build container only (needs
/root/reference; TensorFlow 1.15 itself cannot be installed here):

    python tests/golden/make_graph_goldens.py

Reference code that runs (imported from its file, or exec'd by line range with the text of the range checked first):
  detector/anchor_generator.py:12-166       AnchorGenerator.__call__, tile_anchors
  detector/utils/box_utils.py:63-139        to_center_coordinates, encode, decode
  detector/box_predictor.py:53-90           reshape_and_concatenate (NCHW head outputs -> [B, A, 4] / [B, A])
  detector/retinanet.py:56-81               RetinaNet.get_predictions (sigmoid + batch_non_max_suppression)
  detector/utils/nms.py:6-61                batch_non_max_suppression (threshold, mask, decode, clip, NMS op, gather, pad)
  create_pb.py:90-94                        min-max normalisation + weak-channel mask
  create_pb.py:96-103                       flat person list (boxes, box_ind)
  create_pb.py:106-109                      crop_and_resize call (argument order / crop size)
  detector/prn.py:5-25                      prn()
  create_pb.py:115-142                      softmax over positions, argmax_2d, keypoint_scores / keypoint_positions
  inference/detector.py:49-59               Detector.__call__ post-filter

Third-party TensorFlow kernels are supplied to the stand-in from the CPU oracle (exp / sigmoid recipe, NMS op,
crop_and_resize op) or plain numpy (softmax, matmul): for those the vectors pin the reference's logic AROUND the op, not
the op's last ulp (see the stand-in's docstring).  Everything else is the reference's own float32 arithmetic.
"""
import hashlib
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_numpy_shim as shim  # noqa: E402

f32 = np.float32


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_lines(rel, first, last, must_contain):
    """Source text of lines first..last (1-based, inclusive) of a reference file, dedented; fails loudly when the
    reference no longer has the expected statements there."""
    with open(os.path.join(REF, rel)) as f:
        lines = f.read().split("\n")[first - 1:last]
    text = textwrap.dedent("\n".join(lines))
    for token in must_contain:
        assert token in text, f"{rel}:{first}-{last} no longer contains {token!r}"
    return text


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import oracle
    from multiposenet_b200 import synthetic

    tf = shim.install()

    def np_softmax(x, axis):
        # tf.nn.softmax: exp(x - max) / sum; float32, numpy pairwise summation (order differs from any other kernel:
        # scores are compared at 1e-6, the argmax does not depend on the scale)
        m = x.max(axis=axis, keepdims=True)
        e = oracle.expf((x - m).astype(f32)).reshape(x.shape)
        return (e / e.sum(axis=axis, keepdims=True, dtype=f32)).astype(f32)

    def nms_op(boxes, scores, max_out, iou_thr, score_thr):
        return oracle.nms(boxes, scores, score_thr, iou_thr, max_out)

    def crop_op(img, boxes, box_ind, crop_size):
        return oracle.crop_and_resize(img, boxes, box_ind, crop_size)

    shim.set_kernels(exp=lambda x: oracle.expf(x).reshape(x.shape), sigmoid=lambda x: oracle.sigmoidf(x).reshape(x.shape),
                     softmax=np_softmax, non_max_suppression=nms_op, crop_and_resize=crop_op)

    # ---- the reference's modules, loaded from their files (package __init__ files import the networks: bypassed)
    for pkg in ("detector", "detector.utils"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    constants = load("detector.constants", "detector/constants.py")
    box_utils = load("detector.utils.box_utils", "detector/utils/box_utils.py")
    nms = load("detector.utils.nms", "detector/utils/nms.py")
    sys.modules["detector.utils"].batch_non_max_suppression = nms.batch_non_max_suppression
    sys.modules["detector.utils"].batch_norm_relu = None
    sys.modules["detector.utils"].conv2d_same = None
    anchor_generator = load("detector.anchor_generator", "detector/anchor_generator.py")
    for name in ("detector.training_target_creation", "detector.fpn"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["detector.training_target_creation"].get_training_targets = None
    sys.modules["detector.fpn"].feature_pyramid_network = None
    box_predictor = load("detector.box_predictor", "detector/box_predictor.py")
    retinanet = load("detector.retinanet", "detector/retinanet.py")
    prn_mod = load("detector.prn", "detector/prn.py")
    assert constants.SCALE_FACTORS == [10.0, 10.0, 5.0, 5.0]

    out = {}

    # ---- 1. anchors (anchor_generator.py:12-166).  Full tensors for the small sizes, sha256 + every 61st row otherwise.
    anchor_cases = [(128, 128, "n6"), (256, 384, "n6"), (384, 256, "n9"), (512, 512, "n6"), (640, 640, "n6"), (640, 640, "n9"),
                    (1024, 1024, "n6")]
    for i, (H, W, kind) in enumerate(anchor_cases):
        mult = [1.0, 1.4142] if kind == "n6" else [float(m) for m in synthetic.MULT_9]
        gen = anchor_generator.AnchorGenerator(scale_multipliers=mult)
        a = np.asarray(gen(H, W), f32)
        assert a.dtype == f32 and a.shape[1] == 4
        out[f"anchors_hw_{i}"] = np.array([H, W], np.int32)
        out[f"anchors_mult_{i}"] = np.array(mult, np.float64)
        out[f"anchors_sha_{i}"] = np.array(sha(a))
        out[f"anchors_count_{i}"] = np.int64(a.shape[0])
        out[f"anchors_per_level_{i}"] = np.array([int(v) for v in gen.num_anchors_per_feature_map], np.int64)
        out[f"anchors_rows_{i}"] = a if a.shape[0] < 20000 else a[::61].copy()
    out["anchors_n"] = np.int64(len(anchor_cases))

    # ---- 2. box code round trip (box_utils.py:63-139): encode, decode (tf.exp / tf.log are library kernels)
    rng = np.random.Generator(np.random.PCG64(20241101))
    anc = np.asarray(anchor_generator.AnchorGenerator()(256, 384), f32)
    pick = rng.choice(anc.shape[0], 4000, replace=False)
    codes = np.clip(rng.normal(0, 1.0, (4000, 4)), -4, 4).astype(f32)
    dec = np.asarray(box_utils.decode(shim.T(codes.copy()), shim.T(anc[pick].copy())), f32)
    out["decode_codes"], out["decode_anchors"], out["decode_boxes"] = codes, anc[pick], dec
    gt = synthetic._plant_boxes(rng, 4000).astype(f32)
    out["encode_boxes"] = gt
    out["encode_codes"] = np.asarray(box_utils.encode(shim.T(gt.copy()), shim.T(anc[pick].copy())), f32)

    # ---- 3. layout contract + predictions (box_predictor.py:53-90, retinanet.py:56-81, nms.py:6-61)
    wl = synthetic.WORKLOADS["tiny"]
    inp = synthetic.make_inputs(wl, replicate=7)
    B, n_loc = wl.batch, wl.n_loc
    cls_levels, box_levels, off = [], [], 0
    for s_ in wl.strides:          # what class_net / box_net emit: NCHW, channel = k (class) / k*4 + coord (boxes)
        gh, gw = -(-wl.height // s_), -(-wl.width // s_)
        n = gh * gw * n_loc
        cls_levels.append(inp["class_logits"][:, off:off + n].reshape(B, gh, gw, n_loc).transpose(0, 3, 1, 2).copy())
        box_levels.append(inp["encoded_boxes"][:, off:off + n].reshape(B, gh, gw, n_loc * 4).transpose(0, 3, 1, 2).copy())
        off += n
    raw = box_predictor.reshape_and_concatenate([shim.T(b) for b in box_levels], [shim.T(c) for c in cls_levels], n_loc)
    assert np.array_equal(np.asarray(raw["encoded_boxes"]), inp["encoded_boxes"])        # the layout the C ABI documents
    assert np.array_equal(np.asarray(raw["class_predictions"]), inp["class_logits"])
    # the inputs are synthetic.make_inputs(WORKLOADS["tiny"], replicate=7): the tests regenerate them (and the NCHW level
    # tensors, by the rule above) and check these digests instead of carrying 0.7 MB of random numbers
    out["pred_hw"] = np.array([wl.height, wl.width], np.int32)
    out["pred_levels_sha"] = np.array([sha(c) + sha(b) for c, b in zip(cls_levels, box_levels)])
    out["pred_inputs_sha"] = np.array([sha(inp["class_logits"]), sha(inp["encoded_boxes"]), sha(inp["heatmap_logits"])])

    class FakeRetinaNet:
        pass
    net = FakeRetinaNet()
    net.raw_predictions = raw
    net.anchors = anchor_generator.AnchorGenerator()(wl.height, wl.width)
    k = 0
    for thr, iou_thr, max_det in [(0.3, 0.6, 25), (0.3, 0.5, 25), (0.05, 0.5, 10), (0.5, 0.0, 25), (0.3, 1.0, 25), (0.3, 0.6, 1)]:
        p = retinanet.RetinaNet.get_predictions(net, score_threshold=thr, iou_threshold=iou_thr, max_detections=max_det)
        out[f"pred_params_{k}"] = np.array([thr, iou_thr, max_det], np.float64)
        out[f"pred_boxes_{k}"] = np.asarray(p["boxes"], f32)
        out[f"pred_scores_{k}"] = np.asarray(p["scores"], f32)
        out[f"pred_num_boxes_{k}"] = np.asarray(p["num_boxes"], np.int32)
        k += 1
    out["pred_n"] = np.int64(k)

    # ---- 4. create_pb.py blocks, exec'd from the reference file
    norm_src = ref_lines("create_pb.py", 90, 94, ["M = tf.reduce_max(heatmaps, [1, 2], keepdims=True)", "mask = tf.to_float(M > 0.2)",
                                                  "heatmaps = (heatmaps - m)/(M - m)", "heatmaps *= mask"])
    list_src = ref_lines("create_pb.py", 96, 103, ["for i in range(batch_size):", "boxes.append(predicted_boxes[i][:n])",
                                                   "box_ind = tf.concat(box_ind, axis=0)"])
    crop_src = ref_lines("create_pb.py", 106, 109, ["crops = tf.image.crop_and_resize(", "heatmaps, boxes, box_ind,",
                                                    "crop_size=CROP_SIZE"])
    dec_src = ref_lines("create_pb.py", 115, 142, ["probabilities = tf.nn.softmax(logits, axis=1)", "def argmax_2d(x):",
                                                   "argmax_y = argmax // w", "keypoint_scores = tf.reduce_max(probabilities, axis=[1, 2])",
                                                   "keypoint_positions /= tf.to_float(scaler)"])
    crop_size_src = ref_lines("create_pb.py", 19, 19, ["CROP_SIZE = [56, 36]"])
    kh, seg, mn, mx = oracle.heatmaps(inp["heatmap_logits"])
    kh = kh.copy()
    kh[1, :, :, 4] *= f32(0.15)                  # a weak channel (max <= 0.2): masked by create_pb.py:91,94
    p0 = retinanet.RetinaNet.get_predictions(net, score_threshold=0.3, iou_threshold=0.6, max_detections=25)
    ns = {"tf": tf, "heatmaps": shim.T(kh.copy()), "batch_size": B, "predicted_boxes": p0["boxes"], "num_boxes": p0["num_boxes"]}
    exec(crop_size_src, ns)
    exec(norm_src, ns)
    out["graph_keypoint_heatmaps_sha"] = np.array(sha(kh))       # = oracle.heatmaps(inputs) with channel 4 of image 1 scaled
    out["graph_keypoint_heatmaps_image1"] = kh[1].copy()
    out["graph_normalised_sha"] = np.array(sha(np.asarray(ns["heatmaps"], f32)))
    out["graph_normalised_image1"] = np.asarray(ns["heatmaps"], f32)[1].copy()
    out["graph_min"], out["graph_max"] = np.asarray(ns["m"], f32).reshape(B, 17), np.asarray(ns["M"], f32).reshape(B, 17)
    assert not np.isnan(np.asarray(ns["heatmaps"])).any() and not out["graph_normalised_image1"][:, :, 4].any()
    exec(list_src, ns)
    out["graph_person_boxes"] = np.asarray(ns["boxes"], f32)
    out["graph_person_image"] = np.asarray(ns["box_ind"], np.int32)
    exec(crop_src, ns)
    crops = np.asarray(ns["crops"], f32)
    out["graph_crops_sha"] = np.array(sha(crops))
    out["graph_crops_rows"] = crops[:2].copy()
    N = crops.shape[0]
    assert N >= 4 and crops.shape[1:] == (56, 36, 17)

    # ---- 5. prn (detector/prn.py:5-25) with seeded weights (regenerated by the tests, not stored)
    W1, b1, W2, b2 = synthetic.make_prn_weights(bias_std=0.01)
    shim.set_kernels(fully_connected_weights={"PRN/fc1": (W1, b1), "PRN/fc2": (W2, b2)})
    logits = np.asarray(prn_mod.prn(shim.T(crops.copy()), False), f32)
    assert logits.shape == crops.shape
    out["graph_prn_logits_sha"] = np.array(sha(logits))
    out["graph_prn_logits_rows"] = logits[:2].copy()

    # ---- 6. softmax / argmax_2d / scores / positions (create_pb.py:115-142), on the PRN logits with planted ties
    lg = logits.copy()
    lg[0, :, :, 5] = f32(0.25)                                   # all-equal channel -> first index
    lg[1, 10, 3, 9] = lg[1, 30, 30, 9] = lg[1].max() + f32(1.0)  # exact tie -> first index
    lg[2, 55, 35, 16] = f32(40.0)                                # last position, last channel
    ns.update({"logits": shim.T(lg.copy()), "num_boxes": np.int32(N)})
    exec(dec_src, ns)
    out["graph_decode_logits_rows"] = lg[:3].copy()              # rows 3.. are the PRN logits unchanged
    out["graph_keypoint_scores"] = np.asarray(ns["keypoint_scores"], f32)
    out["graph_keypoint_positions"] = np.asarray(ns["keypoint_positions"], f32)
    assert out["graph_keypoint_positions"].shape == (N, 17, 2)
    assert (out["graph_keypoint_positions"][0, 5] == 0).all()
    assert np.allclose(out["graph_keypoint_positions"][1, 9], [10 / 56, 3 / 36])

    # ---- 7. Detector.__call__ post-filter (inference/detector.py:49-59): plain numpy in the reference
    filt_src = ref_lines("inference/detector.py", 49, 59, ["outputs.update({", "n = outputs['num_boxes']",
                                                          "to_keep = outputs['scores'][:n] > score_threshold",
                                                          "outputs['keypoint_scores'] = outputs['keypoint_scores'][to_keep]"])
    one = {"boxes": out["pred_boxes_0"][:1].copy(), "scores": out["pred_scores_0"][:1].copy(),
           "num_boxes": out["pred_num_boxes_0"][:1].copy(), "keypoint_heatmaps": kh[:1].copy(), "segmentation_masks": seg[:1].copy()}
    n0 = int(one["num_boxes"][0])
    one["keypoint_scores"] = out["graph_keypoint_scores"][:n0].copy()
    one["keypoint_positions"] = out["graph_keypoint_positions"][:n0].copy()
    cut = float(np.sort(one["scores"][0, :n0])[n0 // 2])           # removes about half of the rows; equal score is NOT kept
    for name, v in one.items():
        if name not in ("keypoint_heatmaps", "segmentation_masks"):      # passed through as v[0]
            out[f"filter_in_{name}"] = v
    fns = {"outputs": dict(one), "score_threshold": cut}
    exec(filt_src, fns)
    out["filter_threshold"] = np.float64(cut)
    for name, v in fns["outputs"].items():
        if name in ("keypoint_heatmaps", "segmentation_masks"):
            assert np.array_equal(np.asarray(v), one[name][0])
        else:
            out[f"filter_out_{name}"] = np.asarray(v)
    assert 0 < len(fns["outputs"]["scores"]) < n0

    path = os.path.join(HERE, "graph_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
