"""Host-side mirror of the reference's inference interface, over the C ABI.

`Detector` keeps the output contract of the reference's `inference/detector.py::Detector`
(the seven named outputs of create_pb.py:23-27, batch dimension stripped and rows filtered
by score for a single image exactly as inference/detector.py:49-59), but its inputs are
the tensors at the graph cut -- what the networks hand to the post-processing in
create_pb.py:73-81 -- instead of an image fed to a frozen TensorFlow graph:

    encoded_boxes  [B, A, 4]          retinanet.raw_predictions['encoded_boxes']      (detector/retinanet.py:47-54)
    class_logits   [B, A]             retinanet.raw_predictions['class_predictions']
    heatmap_logits [B, H/4, W/4, 18]  subnet.heatmaps                                  (detector/keypoint_subnet.py:49-58)

All arithmetic happens in libmpn_b200.so (hand-written sm_100a kernels).  PyTorch is used
only for device / pinned memory and streams.  There is no CPU fallback: constructing a
Detector without the library or without a B200 raises.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import MpnConfig, MpnInputs, MpnOutputs, MpnParams, check

OUTPUT_NAMES = ("boxes", "scores", "num_boxes", "keypoint_heatmaps", "segmentation_masks",
                "keypoint_scores", "keypoint_positions")          # create_pb.py:23-27


@dataclass
class DetectorConfig:
    """PARAMS / constants of the reference (create_pb.py:16-36, detector/retinanet.py:38-43, detector/constants.py)."""
    max_batch: int = 1
    max_height: int = 640
    max_width: int = 640
    score_threshold: float = 0.3          # create_pb.py:33
    iou_threshold: float = 0.6            # create_pb.py:34
    max_boxes: int = 25                   # create_pb.py:35
    strides: Sequence[int] = (8, 16, 32, 64, 128)
    scales: Sequence[float] = (32, 64, 128, 256, 512)
    scale_multipliers: Sequence[float] = (1.0, 1.4142)
    aspect_ratios: Sequence[float] = (1.0, 2.0, 0.5)
    scale_factors: Sequence[float] = (10.0, 10.0, 5.0, 5.0)
    crop_size: Tuple[int, int] = (56, 36)
    prn_mode: str = "fp32"                # "fp32" (1e-4 parity mode) or "bf16" (tcgen05 tensor cores, 1e-2)
    prn_modes_allocated: Sequence[str] = ("fp32", "bf16")
    device: int = 0
    extra: dict = field(default_factory=dict)

    @property
    def n_loc(self):
        return len(self.scale_multipliers) * len(self.aspect_ratios)


def _mode_id(name):
    if name in ("fp32", 0):
        return _lib.PRN_FP32
    if name in ("bf16", 1):
        return _lib.PRN_BF16
    raise ValueError(f"unknown prn_mode {name!r}")


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def post_filter(raw, score_threshold):
    """The host-side tail of the reference's Detector.__call__ (inference/detector.py:49-59) on a dict of numpy outputs
    (the seven named tensors + person_offsets): keypoint rows cut to the person count; for ONE image the batch dimension
    is stripped from everything except keypoint_scores / keypoint_positions (:49-52) and the rows with
    `score > score_threshold` are kept (:54-59, strict; `num_boxes` stays unfiltered).  For B > 1 the padded per-image
    tensors of create_pb.py:53-61 are returned with `person_offsets`."""
    raw = dict(raw)
    B = raw["num_boxes"].shape[0]
    N = int(raw["person_offsets"][-1])
    raw["keypoint_scores"] = raw["keypoint_scores"][:N]
    raw["keypoint_positions"] = raw["keypoint_positions"][:N]
    if B > 1:
        return raw
    out = {k: (v if k in ("keypoint_scores", "keypoint_positions", "person_offsets") else v[0]) for k, v in raw.items()}
    out.pop("person_offsets")
    n = int(out["num_boxes"])
    to_keep = out["scores"][:n] > score_threshold
    out["boxes"] = out["boxes"][:n][to_keep]
    out["scores"] = out["scores"][:n][to_keep]
    out["keypoint_positions"] = out["keypoint_positions"][to_keep]
    out["keypoint_scores"] = out["keypoint_scores"][to_keep]
    return out


class Detector:
    """B200 replacement of the post-network part of the reference's frozen graph."""

    def __init__(self, prn_weights=None, config: Optional[DetectorConfig] = None, device: Optional[int] = None):
        """
        Arguments:
            prn_weights: (W1 [D,1024], b1 [1024], W2 [1024,D], b2 [D]) float32 arrays, the variables
                PRN/fc1/{weights,biases}, PRN/fc2/{weights,biases} (detector/prn.py:13,20,22); may be set later.
            config: a DetectorConfig.
            device: CUDA device ordinal (overrides config.device).
        """
        self.config = config or DetectorConfig()
        if device is not None:
            self.config.device = int(device)
        cfg = self.config
        self._lib = _lib.load()
        c = MpnConfig()
        check(self._lib.mpn_default_config(C.byref(c)))
        c.device = cfg.device
        c.max_batch, c.max_height, c.max_width, c.max_detections = cfg.max_batch, cfg.max_height, cfg.max_width, cfg.max_boxes
        c.num_levels = len(cfg.strides)
        for i, (s, sc) in enumerate(zip(cfg.strides, cfg.scales)):
            c.strides[i] = int(s)
            c.scales[i] = float(sc)
        c.num_multipliers = len(cfg.scale_multipliers)
        for i, m in enumerate(cfg.scale_multipliers):
            c.multipliers[i] = float(m)
        c.num_ratios = len(cfg.aspect_ratios)
        for i, r in enumerate(cfg.aspect_ratios):
            c.ratios[i] = float(r)
        for i in range(4):
            c.scale_factors[i] = float(cfg.scale_factors[i])
        c.crop_height, c.crop_width = cfg.crop_size
        c.prn_modes = sum(1 << _mode_id(m) for m in cfg.prn_modes_allocated)
        self._handle = C.c_void_p()
        check(self._lib.mpn_create(C.byref(c), C.byref(self._handle)))
        self.device = torch.device("cuda", cfg.device)
        self.D = cfg.crop_size[0] * cfg.crop_size[1] * 17
        self._host_out = None
        self._next_slot = 0
        self._dev_out = {}
        if prn_weights is not None:
            self.set_prn_weights(*prn_weights)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.mpn_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        check(rc, self._handle)

    # ------------------------------------------------------------------ weights
    def set_prn_weights(self, W1, b1, W2, b2):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (W1, b1, W2, b2)]
        D, Hd = self.D, 1024
        if arrs[0].shape != (D, Hd) or arrs[1].shape != (Hd,) or arrs[2].shape != (Hd, D) or arrs[3].shape != (D,):
            raise ValueError("PRN weights must have shapes W1 [D,1024], b1 [1024], W2 [1024,D], b2 [D], D = %d" % D)
        self._check(self._lib.mpn_set_prn_weights(self._handle, *[_ptr(a) for a in arrs]))

    # ------------------------------------------------------------------ helpers
    def num_anchors(self, height, width):
        n = self._lib.mpn_num_anchors(self._handle, int(height), int(width))
        if n < 0:
            raise ValueError("bad image size")
        return n

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _params(self, score_threshold=None, iou_threshold=None, max_boxes=None, prn_mode=None):
        cfg = self.config
        p = MpnParams()
        p.score_threshold = cfg.score_threshold if score_threshold is None else score_threshold
        p.iou_threshold = cfg.iou_threshold if iou_threshold is None else iou_threshold
        p.max_detections = cfg.max_boxes if max_boxes is None else max_boxes
        p.prn_mode = _mode_id(cfg.prn_mode if prn_mode is None else prn_mode)
        return p

    def _inputs(self, encoded_boxes, class_logits, heatmap_logits, image_hw):
        per_level = isinstance(class_logits, (list, tuple))
        inp = MpnInputs()
        keep = []
        if per_level:
            B = int(class_logits[0].shape[0])
            n = len(class_logits)
            cls_arr = (C.c_void_p * n)(*[t.data_ptr() for t in class_logits])
            box_arr = (C.c_void_p * n)(*[t.data_ptr() for t in encoded_boxes])
            inp.level_class = C.cast(cls_arr, C.POINTER(C.c_void_p))
            inp.level_boxes = C.cast(box_arr, C.POINTER(C.c_void_p))
            keep += [cls_arr, box_arr]
        else:
            B = int(class_logits.shape[0])
            inp.class_logits = _ptr(class_logits)
            inp.encoded_boxes = _ptr(encoded_boxes)
        if image_hw is None:
            if heatmap_logits is None:
                raise ValueError("image_hw is required when heatmap_logits is not given")
            image_hw = (int(heatmap_logits.shape[1]) * 4, int(heatmap_logits.shape[2]) * 4)
        H, W = int(image_hw[0]), int(image_hw[1])
        # the reference asserts this (inference/detector.py:44-45)
        if H % 128 != 0 or W % 128 != 0:
            raise ValueError("image height and width must be divisible by 128")
        inp.batch, inp.height, inp.width = B, H, W
        inp.heatmap_logits = _ptr(heatmap_logits)
        return inp, keep, B, H, W

    def _validate(self, encoded_boxes, class_logits, heatmap_logits, B, H, W, want_cuda):
        A = self.num_anchors(H, W)
        tensors = []
        if isinstance(class_logits, (list, tuple)):
            n_loc = self.config.n_loc
            for lvl, (cl, bx) in enumerate(zip(class_logits, encoded_boxes)):
                gh, gw = -(-H // self.config.strides[lvl]), -(-W // self.config.strides[lvl])
                if tuple(cl.shape) != (B, n_loc, gh, gw) or tuple(bx.shape) != (B, 4 * n_loc, gh, gw):
                    raise ValueError(f"level {lvl}: expected NCHW [{B},{n_loc},{gh},{gw}] / [{B},{4 * n_loc},{gh},{gw}]")
                tensors += [cl, bx]
        else:
            if tuple(class_logits.shape) != (B, A) or tuple(encoded_boxes.shape) != (B, A, 4):
                raise ValueError(f"expected class_logits [{B},{A}] and encoded_boxes [{B},{A},4] for a {H}x{W} image")
            tensors += [class_logits, encoded_boxes]
        if heatmap_logits is not None:
            if tuple(heatmap_logits.shape) != (B, H // 4, W // 4, 18):
                raise ValueError(f"expected heatmap_logits [{B},{H // 4},{W // 4},18]")
            tensors.append(heatmap_logits)
        for t in tensors:
            if isinstance(t, torch.Tensor):
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise ValueError("inputs must be contiguous float32")
                if t.is_cuda != want_cuda:
                    raise ValueError("inputs must all live on the same side (all CUDA or all host)")
            else:
                if want_cuda or t.dtype != np.float32 or not t.flags["C_CONTIGUOUS"]:
                    raise ValueError("inputs must be contiguous float32 (numpy arrays are host inputs)")

    # ------------------------------------------------------------------ device-resident path
    def _device_outputs(self, B, H, W, max_det):
        key = (B, H, W, max_det)
        if key not in self._dev_out:
            dev, f32, i32 = self.device, torch.float32, torch.int32
            h4, w4, NP = H // 4, W // 4, B * max_det
            if len(self._dev_out) >= 8:         # a caller cycling through many shapes: keep the most recent few
                self._dev_out.pop(next(iter(self._dev_out)))
            self._dev_out[key] = {
                "boxes": torch.zeros((B, max_det, 4), dtype=f32, device=dev),
                "scores": torch.zeros((B, max_det), dtype=f32, device=dev),
                "num_boxes": torch.zeros((B,), dtype=i32, device=dev),
                "keypoint_heatmaps": torch.empty((B, h4, w4, 17), dtype=f32, device=dev),
                "segmentation_masks": torch.empty((B, h4, w4), dtype=f32, device=dev),
                "keypoint_scores": torch.zeros((NP, 17), dtype=f32, device=dev),
                "keypoint_positions": torch.zeros((NP, 17, 2), dtype=f32, device=dev),
                "person_offsets": torch.zeros((B + 1,), dtype=i32, device=dev),
            }
        return self._dev_out[key]

    def run_device(self, encoded_boxes, class_logits, heatmap_logits, image_hw=None, score_threshold=None,
                   iou_threshold=None, max_boxes=None, prn_mode=None, outputs=None):
        """Whole path on device tensors, asynchronous on torch's current stream.  Returns a dict of CUDA tensors
        padded to max_boxes (rows of keypoint_* beyond person_offsets[-1] are stale); buffers are reused per shape."""
        inp, keep, B, H, W = self._inputs(encoded_boxes, class_logits, heatmap_logits, image_hw)
        self._validate(encoded_boxes, class_logits, heatmap_logits, B, H, W, want_cuda=True)
        p = self._params(score_threshold, iou_threshold, max_boxes, prn_mode)
        out = outputs if outputs is not None else self._device_outputs(B, H, W, p.max_detections)
        o = MpnOutputs()
        for name in ("boxes", "scores", "num_boxes", "keypoint_heatmaps", "segmentation_masks", "keypoint_scores",
                     "keypoint_positions", "person_offsets"):
            setattr(o, name, _ptr(out.get(name)))
        self._check(self._lib.mpn_run(self._handle, C.byref(inp), C.byref(p), C.byref(o), self._stream()))
        return out

    # ------------------------------------------------------------------ host path (feed_dict / fetch semantics)
    def _host_outputs(self, B, H, W, max_det, slot):
        key = (B, H, W, max_det)
        if self._host_out is None or self._host_out[0] != key:
            self._host_out = (key, {})
        ring = self._host_out[1]
        if slot not in ring:
            h4, w4, NP = H // 4, W // 4, B * max_det
            pin = dict(pin_memory=True)
            ring[slot] = {
                "boxes": torch.zeros((B, max_det, 4), dtype=torch.float32, **pin),
                "scores": torch.zeros((B, max_det), dtype=torch.float32, **pin),
                "num_boxes": torch.zeros((B,), dtype=torch.int32, **pin),
                "keypoint_heatmaps": torch.empty((B, h4, w4, 17), dtype=torch.float32, **pin),
                "segmentation_masks": torch.empty((B, h4, w4), dtype=torch.float32, **pin),
                "keypoint_scores": torch.zeros((NP, 17), dtype=torch.float32, **pin),
                "keypoint_positions": torch.zeros((NP, 17, 2), dtype=torch.float32, **pin),
                "person_offsets": torch.zeros((B + 1,), dtype=torch.int32, **pin),
            }
        return ring[slot]

    def submit_host(self, encoded_boxes, class_logits, heatmap_logits, image_hw=None, score_threshold=None,
                    iou_threshold=None, max_boxes=None, prn_mode=None, return_heatmaps=True):
        """Enqueue copy-in, the path and copy-out (inference/detector.py:47-48) without waiting; up to
        _lib.HOST_DEPTH calls may be in flight, so the PCIe copies of neighbouring calls overlap the kernels.
        Host inputs (numpy or torch CPU, ideally pinned; they must stay untouched until wait()).  Returns
        (ticket, dict of pinned output buffers); the buffers of ticket t are re-used by ticket t + HOST_DEPTH."""
        inp, keep, B, H, W = self._inputs(encoded_boxes, class_logits, heatmap_logits, image_hw)
        self._validate(encoded_boxes, class_logits, heatmap_logits, B, H, W, want_cuda=False)
        p = self._params(score_threshold, iou_threshold, max_boxes, prn_mode)
        out = self._host_outputs(B, H, W, p.max_detections, self._next_slot)
        self._next_slot = (self._next_slot + 1) % _lib.HOST_DEPTH
        o = MpnOutputs()
        for name in ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions", "person_offsets"):
            setattr(o, name, _ptr(out[name]))
        if return_heatmaps:
            o.keypoint_heatmaps = _ptr(out["keypoint_heatmaps"])
            o.segmentation_masks = _ptr(out["segmentation_masks"])
        ticket = C.c_int64(-1)
        self._check(self._lib.mpn_submit_host(self._handle, C.byref(inp), C.byref(p), C.byref(o), C.byref(ticket)))
        return ticket.value, out

    def wait(self, ticket):
        """Block until the outputs of submit_host ticket `ticket` are in its host buffers."""
        self._check(self._lib.mpn_wait(self._handle, int(ticket)))

    def run_host_async(self, *args, **kwargs):
        """submit_host without the ticket (kept for one-call-at-a-time use): returns the pinned output buffers;
        call synchronize() before reading them."""
        return self.submit_host(*args, **kwargs)[1]

    def synchronize(self):
        self._check(self._lib.mpn_synchronize(self._handle))

    def host_traffic(self):
        """(host->device, device->host) bytes the copy engines moved for the most recent host call."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._lib.mpn_host_traffic(self._handle, C.byref(a), C.byref(b))
        return a.value, b.value

    # ------------------------------------------------------------------ the reference-facing call
    def __call__(self, encoded_boxes, class_logits, heatmap_logits, image_hw=None, score_threshold=0.05,
                 return_heatmaps=True, copy=True):
        """
        Arguments:
            encoded_boxes, class_logits, heatmap_logits: the tensors at the graph cut (see module docstring); numpy
                arrays / torch CPU tensors (copied in, like feed_dict) or torch CUDA tensors (used in place).
            image_hw: (H, W) of the network input; inferred from heatmap_logits when omitted.
            score_threshold: the post-filter of inference/detector.py:36,54-59 (NOT the graph's NMS threshold,
                which is config.score_threshold).
        Returns:
            dict with the keys of OUTPUT_NAMES (+ 'person_offsets' for B > 1), numpy arrays.  For one image the
            shapes and the filtering are those of inference/detector.py:49-59; for B > 1 they are those of
            create_pb.py:53-61 (padded per-image boxes/scores, persons of all images concatenated).
        """
        cuda_in = isinstance(heatmap_logits, torch.Tensor) and heatmap_logits.is_cuda
        if cuda_in:
            dev = self.run_device(encoded_boxes, class_logits, heatmap_logits, image_hw)
            torch.cuda.current_stream(self.device).synchronize()
            raw = {k: v.cpu().numpy() for k, v in dev.items() if return_heatmaps or k not in
                   ("keypoint_heatmaps", "segmentation_masks")}
        else:
            bufs = self.run_host_async(encoded_boxes, class_logits, heatmap_logits, image_hw,
                                       return_heatmaps=return_heatmaps)
            self.synchronize()
            raw = {k: (v.numpy().copy() if copy else v.numpy()) for k, v in bufs.items()
                   if return_heatmaps or k not in ("keypoint_heatmaps", "segmentation_masks")}
        return post_filter(raw, score_threshold)

    # ------------------------------------------------------------------ single stages (CUDA tensors)
    def anchors(self, height, width):
        """detector/anchor_generator.py:40-116 -> [A,4] CUDA tensor."""
        A = self.num_anchors(height, width)
        out = torch.empty((A, 4), dtype=torch.float32, device=self.device)
        self._check(self._lib.mpn_anchors(self._handle, int(height), int(width), _ptr(out), self._stream()))
        return out

    def detect(self, encoded_boxes, class_logits, image_hw, score_threshold=None, iou_threshold=None, max_boxes=None):
        """detector/retinanet.py:56-81 -> dict(boxes, scores, num_boxes, sel_anchor, n_candidates) of CUDA tensors."""
        inp, keep, B, H, W = self._inputs(encoded_boxes, class_logits, None, image_hw)
        self._validate(encoded_boxes, class_logits, None, B, H, W, want_cuda=True)
        p = self._params(score_threshold, iou_threshold, max_boxes)
        md, dev = p.max_detections, self.device
        out = {"boxes": torch.empty((B, md, 4), dtype=torch.float32, device=dev),
               "scores": torch.empty((B, md), dtype=torch.float32, device=dev),
               "num_boxes": torch.empty((B,), dtype=torch.int32, device=dev),
               "sel_anchor": torch.empty((B, md), dtype=torch.int32, device=dev),
               "n_candidates": torch.empty((B,), dtype=torch.int32, device=dev)}
        self._check(self._lib.mpn_detect(self._handle, C.byref(inp), C.byref(p), _ptr(out["boxes"]), _ptr(out["scores"]),
                                         _ptr(out["num_boxes"]), _ptr(out["sel_anchor"]), _ptr(out["n_candidates"]),
                                         self._stream()))
        return out

    def heatmaps(self, heatmap_logits):
        """create_pb.py:73-76,90,92 -> (keypoint_heatmaps, segmentation_masks, minmax [B,17,2])."""
        B, h, w, _ = heatmap_logits.shape
        dev = self.device
        kh = torch.empty((B, h, w, 17), dtype=torch.float32, device=dev)
        seg = torch.empty((B, h, w), dtype=torch.float32, device=dev)
        mm = torch.empty((B, 17, 2), dtype=torch.float32, device=dev)
        self._check(self._lib.mpn_heatmaps(self._handle, _ptr(heatmap_logits), B, h, w, _ptr(kh), _ptr(seg), _ptr(mm),
                                           self._stream()))
        return kh, seg, mm

    def heatmaps_normalised(self, heatmap_logits):
        """create_pb.py:73-76 + :90-94 in the two-pass form of the full path (min / max from the logits, then one pass)
        -> (keypoint_heatmaps, segmentation_masks, minmax [B,17,2], normalised [B,h,w,20]; channels 17..19 are padding)."""
        B, h, w, _ = heatmap_logits.shape
        dev = self.device
        kh = torch.empty((B, h, w, 17), dtype=torch.float32, device=dev)
        seg = torch.empty((B, h, w), dtype=torch.float32, device=dev)
        mm = torch.empty((B, 17, 2), dtype=torch.float32, device=dev)
        nh = torch.zeros((B, h, w, 20), dtype=torch.float32, device=dev)
        self._check(self._lib.mpn_heatmaps_normalised(self._handle, _ptr(heatmap_logits), B, h, w, _ptr(kh), _ptr(seg), _ptr(mm),
                                                      _ptr(nh), self._stream()))
        return kh, seg, mm, nh

    def crop_padded(self, normalised, boxes, box_ind):
        """create_pb.py:106-109 on a normalised map: [B,h,w,17] (padded here) or already padded [B,h,w,20]
        -> (crops f32 [N,56,36,17], the same crops in bfloat16), by the crop kernel the full path uses."""
        if normalised.shape[-1] == 17:
            normalised = torch.nn.functional.pad(normalised, (0, 3)).contiguous()
        B, h, w, _ = normalised.shape
        N = int(boxes.shape[0])
        ch, cw = self.config.crop_size
        f = torch.empty((N, ch, cw, 17), dtype=torch.float32, device=self.device)
        b = torch.empty((N, ch, cw, 17), dtype=torch.bfloat16, device=self.device)
        if N:
            self._check(self._lib.mpn_crop_padded(self._handle, _ptr(normalised), B, h, w, _ptr(boxes), _ptr(box_ind), N,
                                                  _ptr(f), _ptr(b), self._stream()))
        return f, b

    def heatmap_head(self, features, weight, bias, want_logits=True):
        """detector/keypoint_subnet.py:49-58 (1x1 conv 64 -> 18 + bias, NCHW -> NHWC) fused with heatmaps():
        features [B,64,h,w] CUDA f32 -> (heatmap_logits [B,h,w,18] or None, keypoint_heatmaps, segmentation_masks, minmax)."""
        B, cin, h, w = features.shape
        if cin != 64 or tuple(weight.shape) != (64, 18) or tuple(bias.shape) != (18,):
            raise ValueError("expected features [B,64,h,w], weight [64,18], bias [18]")
        dev = self.device
        lg = torch.empty((B, h, w, 18), dtype=torch.float32, device=dev) if want_logits else None
        kh = torch.empty((B, h, w, 17), dtype=torch.float32, device=dev)
        seg = torch.empty((B, h, w), dtype=torch.float32, device=dev)
        mm = torch.empty((B, 17, 2), dtype=torch.float32, device=dev)
        self._check(self._lib.mpn_heatmap_head(self._handle, _ptr(features), _ptr(weight), _ptr(bias), B, h, w, _ptr(lg),
                                               _ptr(kh), _ptr(seg), _ptr(mm), self._stream()))
        return lg, kh, seg, mm

    def crop(self, keypoint_heatmaps, boxes, box_ind, minmax=None):
        """create_pb.py:90-94,106-109 -> crops [N,56,36,17]."""
        B, h, w, _ = keypoint_heatmaps.shape
        N = int(boxes.shape[0])
        ch, cw = self.config.crop_size
        out = torch.empty((N, ch, cw, 17), dtype=torch.float32, device=self.device)
        if N:
            self._check(self._lib.mpn_crop(self._handle, _ptr(keypoint_heatmaps), _ptr(minmax), B, h, w, _ptr(boxes),
                                           _ptr(box_ind), N, _ptr(out), self._stream()))
        return out

    def prn(self, crops, prn_mode=None, inplace=False):
        """detector/prn.py:5-25 -> logits, same shape as crops.  inplace=True (bf16 mode) overwrites `crops` with the
        logits, as the full path does internally."""
        N = int(crops.shape[0])
        out = crops if inplace else torch.empty_like(crops)
        mode = _mode_id(self.config.prn_mode if prn_mode is None else prn_mode)
        self._check(self._lib.mpn_prn(self._handle, _ptr(crops), N, mode, _ptr(out), self._stream()))
        return out

    def keypoint_decode(self, logits):
        """create_pb.py:115-142 -> (keypoint_scores [N,17], keypoint_positions [N,17,2], argmax [N,17])."""
        N = int(logits.shape[0])
        dev = self.device
        s = torch.empty((N, 17), dtype=torch.float32, device=dev)
        pos = torch.empty((N, 17, 2), dtype=torch.float32, device=dev)
        arg = torch.empty((N, 17), dtype=torch.int32, device=dev)
        if N:
            self._check(self._lib.mpn_keypoint_decode(self._handle, _ptr(logits), N, _ptr(s), _ptr(pos), _ptr(arg),
                                                      self._stream()))
        return s, pos, arg

    def get_keypoints(self, heatmaps, box, threshold):
        """inference/utils.py:29-52 on the device -> numpy int32 [17,3] rows (x, y, visible)."""
        hm = torch.as_tensor(np.ascontiguousarray(heatmaps, dtype=np.float32)).to(self.device) \
            if not isinstance(heatmaps, torch.Tensor) else heatmaps
        h, w, _ = hm.shape
        out = torch.empty((17, 3), dtype=torch.int32, device=self.device)
        b = (C.c_double * 4)(*[float(v) for v in box])
        self._check(self._lib.mpn_get_keypoints(self._handle, _ptr(hm), h, w, b, float(threshold), _ptr(out),
                                                self._stream()))
        return out.cpu().numpy()

    def device_exp(self, x):
        y = torch.empty_like(x)
        self._check(self._lib.mpn_test_exp(self._handle, _ptr(x), _ptr(y), x.numel(), self._stream()))
        return y

    def device_sigmoid(self, x):
        y = torch.empty_like(x)
        self._check(self._lib.mpn_test_sigmoid(self._handle, _ptr(x), _ptr(y), x.numel(), self._stream()))
        return y

    def sigmoid_monotone_violations(self, key_begin=0, count=1 << 32):
        """Test hook (mpn_test_sigmoid_monotone): neighbouring floats x < x' with sigmoid(x) > sigmoid(x') -- or on which
        the packed-pair form of the recipe differs from the scalar one -- among `count` floats in increasing order."""
        counter = torch.zeros((1,), dtype=torch.int64, device=self.device)
        self._check(self._lib.mpn_test_sigmoid_monotone(self._handle, int(key_begin), int(count), _ptr(counter), self._stream()))
        return int(counter.item())

    def debug_fetch(self, what):
        """Test hook (mpn_debug_fetch): one internal buffer of the most recent run_device / submit_host as a CUDA tensor
        ("normalised", "crops_f32", "crops_bf16", "logits", "minmax", "person_box", "person_image")."""
        code = _lib.DEBUG_BUFFERS[what]
        n = C.c_int64(0)
        self._check(self._lib.mpn_debug_fetch(self._handle, code, None, 0, C.byref(n)))
        dtype = {"crops_bf16": torch.bfloat16, "person_image": torch.int32}.get(what, torch.float32)
        out = torch.empty((n.value // torch.empty((), dtype=dtype).element_size(),), dtype=dtype, device=self.device)
        self._check(self._lib.mpn_debug_fetch(self._handle, code, _ptr(out), n.value, C.byref(n)))
        return out

    def set_profiling(self, enable=True):
        """Per-kernel CUDA-event timing of run_device / run_host_async (bench.py's roofline leg)."""
        self._check(self._lib.mpn_set_profiling(self._handle, 1 if enable else 0))

    def profile(self):
        """[(kernel name, milliseconds)] of the most recent run (waits for it to finish)."""
        cap = 32
        names, ms, n = (C.c_char_p * cap)(), (C.c_float * cap)(), C.c_int32(0)
        self._check(self._lib.mpn_get_profile(self._handle, cap, names, ms, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def fused_trace(self, enable=True):
        """Development aid (mpn_debug_fused_trace): numpy uint64 [grid, 16] of per-CTA phase timestamps in ns."""
        cap = 1024 * 16
        buf = (C.c_uint64 * cap)()
        g = C.c_int32(0)
        self._check(self._lib.mpn_debug_fused_trace(self._handle, 1 if enable else 0, buf, cap, C.byref(g)))
        return np.frombuffer(buf, dtype=np.uint64, count=g.value * 16).reshape(g.value, 16).copy()

    def nms_trace(self, enable=True):
        """Development aid (mpn_debug_nms_trace): numpy uint64 [max_batch, 16] of per-image phase timestamps in ns."""
        n = self.config.max_batch * 16
        buf = (C.c_uint64 * n)()
        self._check(self._lib.mpn_debug_nms_trace(self._handle, 1 if enable else 0, buf, n))
        return np.frombuffer(buf, dtype=np.uint64, count=n).reshape(-1, 16).copy()

    def debug_skip(self, mask):
        """Development aid (mpn_debug_skip): do not launch the stages whose bit is set (1 detect, 2 heatmap stage, 8 crop,
        16 PRN, 32 keypoint decode); their outputs keep the previous call's values."""
        self._check(self._lib.mpn_debug_skip(self._handle, int(mask)))

    def launch_count(self):
        last, total = C.c_int64(0), C.c_int64(0)
        self._lib.mpn_launch_count(self._handle, C.byref(last), C.byref(total))
        return last.value, total.value


class DetectorLanes:
    """Throughput mode for a stream of device-resident batches: `lanes` Detector handles on one GPU, each with its own
    CUDA stream, fed round-robin.  Calls of neighbouring batches overlap: the latency-bound front half of one call
    (candidate scan, sort / NMS, heatmap pass) and its keypoint decode run beside another call's kernels instead of
    leaving the device to one short grid at a time.  The PRN kernel owns every SM while it runs, so the gain is bounded by
    the time of the other six kernels (c2, per batch, round 2: 79.9 us alone, 73.5 with 2 lanes, 79.1 with 3, 75.5 with 4;
    tools/two_streams.py).  Every lane is an
    independent handle (own workspace and weight copy, 0.3 GB each), so results are those of a lone Detector, bit for
    bit; a handle is not re-entrant (include/mpn_b200.h), the lanes are what makes concurrent calls legal."""

    def __init__(self, prn_weights, config: Optional[DetectorConfig] = None, lanes: int = 2, device: Optional[int] = None):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.detectors = [Detector(prn_weights, config, device) for _ in range(lanes)]
        self.device = self.detectors[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self._next = 0

    def __len__(self):
        return len(self.detectors)

    def fork(self, event=None):
        """Every lane waits for `event` (default: everything submitted so far to the caller's current stream)."""
        if event is None:
            event = torch.cuda.current_stream(self.device).record_event()
        for st in self.streams:
            st.wait_event(event)

    def submit(self, encoded_boxes, class_logits, heatmap_logits, image_hw=None, **kwargs):
        """run_device on the next lane's stream; returns (lane index, that lane's output dict).  The dict is reused by the
        lane's next call: consume it (on that lane's stream, or after join()) before `lanes` more submits."""
        lane = self._next
        self._next = (lane + 1) % len(self.detectors)
        with torch.cuda.stream(self.streams[lane]):
            out = self.detectors[lane].run_device(encoded_boxes, class_logits, heatmap_logits, image_hw, **kwargs)
        return lane, out

    def join(self):
        """The caller's current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def launch_count(self):
        return sum(d.launch_count()[1] for d in self.detectors)

    def close(self):
        for d in self.detectors:
            d.close()
