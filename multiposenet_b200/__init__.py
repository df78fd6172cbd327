"""multiposenet_b200 -- the post-backbone inference path of TropComplique/MultiPoseNet
(anchors -> decode -> NMS -> heatmap normalise -> crop_and_resize -> PRN -> keypoint
decode) as hand-written sm_100a CUDA kernels behind a C ABI (include/mpn_b200.h).

    from multiposenet_b200 import Detector, DetectorConfig

`Detector` mirrors the reference's inference/detector.py::Detector output contract.
Importing the package does not load the CUDA library; constructing a Detector does,
and raises if the library or a B200 is missing (there is no CPU fallback).
"""
__all__ = ["Detector", "DetectorConfig", "DetectorLanes", "OUTPUT_NAMES"]


def __getattr__(name):
    if name in __all__:
        from . import detector
        return getattr(detector, name)
    raise AttributeError(name)
