"""ctypes binding of libmpn_b200.so (include/mpn_b200.h).

The library is loaded from the package directory.  If it is missing or cannot be
loaded this module raises: there is no CPU fallback and none is attempted.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MPN_LIB: a variant build of the same ABI (multiposenet_b200/build.py: VARIANTS) for measurements; default: the product library
LIB_PATH = os.environ.get("MPN_LIB") or os.path.join(HERE, "libmpn_b200.so")

MPN_MAX_LEVELS = 8
MPN_MAX_ANCHOR_SHAPES = 16

MPN_OK = 0
MPN_ERR_INVALID_ARGUMENT = -1
MPN_ERR_CUDA = -2
MPN_ERR_UNSUPPORTED = -3
MPN_ERR_NO_WEIGHTS = -4
MPN_ERR_CAPACITY = -5

PRN_FP32 = 0
PRN_BF16 = 1
HOST_DEPTH = 3          # MPN_HOST_DEPTH
DEBUG_BUFFERS = {"normalised": 0, "crops_f32": 1, "crops_bf16": 2, "logits": 3, "minmax": 4, "person_box": 5,
                 "person_image": 6}          # MPN_DEBUG_*

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class MpnConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32),
        ("max_height", C.c_int32), ("max_width", C.c_int32), ("max_detections", C.c_int32),
        ("num_levels", C.c_int32), ("strides", C.c_int32 * MPN_MAX_LEVELS), ("scales", C.c_double * MPN_MAX_LEVELS),
        ("num_multipliers", C.c_int32), ("multipliers", C.c_double * MPN_MAX_ANCHOR_SHAPES),
        ("num_ratios", C.c_int32), ("ratios", C.c_double * MPN_MAX_ANCHOR_SHAPES),
        ("scale_factors", C.c_float * 4), ("crop_height", C.c_int32), ("crop_width", C.c_int32),
        ("num_keypoints", C.c_int32), ("downsample", C.c_int32), ("prn_hidden", C.c_int32), ("prn_modes", C.c_int32),
    ]


class MpnParams(C.Structure):
    _fields_ = [("score_threshold", C.c_float), ("iou_threshold", C.c_float), ("max_detections", C.c_int32),
                ("prn_mode", C.c_int32)]


class MpnInputs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("class_logits", C.c_void_p), ("encoded_boxes", C.c_void_p), ("heatmap_logits", C.c_void_p),
                ("level_class", C.POINTER(C.c_void_p)), ("level_boxes", C.POINTER(C.c_void_p))]


class MpnOutputs(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("scores", C.c_void_p), ("num_boxes", C.c_void_p),
                ("keypoint_heatmaps", C.c_void_p), ("segmentation_masks", C.c_void_p),
                ("keypoint_scores", C.c_void_p), ("keypoint_positions", C.c_void_p), ("person_offsets", C.c_void_p)]


# every symbol include/mpn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mpn_version": (C.c_int, []),
    "mpn_last_error": (C.c_char_p, [C.c_void_p]),
    "mpn_default_config": (C.c_int, [C.POINTER(MpnConfig)]),
    "mpn_create": (C.c_int, [C.POINTER(MpnConfig), C.POINTER(C.c_void_p)]),
    "mpn_destroy": (None, [C.c_void_p]),
    "mpn_num_anchors": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "mpn_set_prn_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_run": (C.c_int, [C.c_void_p, C.POINTER(MpnInputs), C.POINTER(MpnParams), C.POINTER(MpnOutputs), C.c_void_p]),
    "mpn_run_host": (C.c_int, [C.c_void_p, C.POINTER(MpnInputs), C.POINTER(MpnParams), C.POINTER(MpnOutputs)]),
    "mpn_synchronize": (C.c_int, [C.c_void_p]),
    "mpn_submit_host": (C.c_int, [C.c_void_p, C.POINTER(MpnInputs), C.POINTER(MpnParams), C.POINTER(MpnOutputs),
                                  C.POINTER(C.c_int64)]),
    "mpn_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "mpn_host_traffic": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mpn_anchors": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mpn_detect": (C.c_int, [C.c_void_p, C.POINTER(MpnInputs), C.POINTER(MpnParams), C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_heatmaps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "mpn_heatmaps_normalised": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_crop_padded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_heatmap_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_crop": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                           C.c_int32, C.c_void_p, C.c_void_p]),
    "mpn_prn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mpn_keypoint_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpn_get_keypoints": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_double,
                                    C.c_void_p, C.c_void_p]),
    "mpn_test_exp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mpn_test_sigmoid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mpn_test_sigmoid_monotone": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mpn_debug_fetch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "mpn_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "mpn_get_profile": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "mpn_debug_fused_trace": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int32)]),
    "mpn_debug_nms_trace": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), C.c_int32]),
    "mpn_debug_skip": (C.c_int, [C.c_void_p, C.c_uint32]),
    "mpn_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load libmpn_b200.so and set the prototypes.  Raises if the library is absent: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m multiposenet_b200.build` "
            "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MpnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"mpn status {code}: {message}")
        self.code = code


def check(rc, handle=None):
    """Map a status to the reference's error behaviour: bad shapes -> ValueError / AssertionError-like,
    everything else -> RuntimeError."""
    if rc == MPN_OK:
        return
    msg = load().mpn_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc in (MPN_ERR_INVALID_ARGUMENT, MPN_ERR_CAPACITY):
        raise ValueError(f"mpn status {rc}: {msg}")
    raise MpnError(rc, msg)
