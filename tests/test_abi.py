"""The C-ABI library loads and exports every symbol include/mpn_b200.h declares; host-only entry points work and
the product path fails loudly without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from multiposenet_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from multiposenet_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in mpn_b200.h but not exported"


def test_default_config_and_struct_layout(lib):
    from multiposenet_b200._lib import MpnConfig
    c = MpnConfig()
    assert lib.mpn_default_config(C.byref(c)) == 0
    assert c.struct_size == C.sizeof(MpnConfig)
    assert list(c.strides)[:5] == [8, 16, 32, 64, 128] and c.num_levels == 5
    assert list(c.scales)[:5] == [32, 64, 128, 256, 512]
    assert (c.num_multipliers, c.num_ratios) == (2, 3) and abs(c.multipliers[1] - 1.4142) < 1e-12
    assert list(c.scale_factors) == [10.0, 10.0, 5.0, 5.0]
    assert (c.crop_height, c.crop_width, c.num_keypoints, c.prn_hidden, c.max_detections) == (56, 36, 17, 1024, 25)
    assert lib.mpn_version() >= 100


def test_create_rejects_bad_config_before_touching_the_gpu(lib):
    from multiposenet_b200._lib import MpnConfig, MPN_ERR_INVALID_ARGUMENT, MPN_ERR_UNSUPPORTED
    c = MpnConfig()
    lib.mpn_default_config(C.byref(c))
    h = C.c_void_p()
    c.struct_size = 12
    assert lib.mpn_create(C.byref(c), C.byref(h)) == MPN_ERR_INVALID_ARGUMENT
    assert b"size mismatch" in lib.mpn_last_error(None)
    lib.mpn_default_config(C.byref(c))
    c.max_height = 600                      # not a multiple of 128 (inference/detector.py:45)
    assert lib.mpn_create(C.byref(c), C.byref(h)) == MPN_ERR_INVALID_ARGUMENT
    lib.mpn_default_config(C.byref(c))
    c.num_keypoints = 12
    assert lib.mpn_create(C.byref(c), C.byref(h)) == MPN_ERR_UNSUPPORTED
    assert not h.value


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; the loud-failure path is for GPU-less hosts")
    from multiposenet_b200 import Detector
    from multiposenet_b200._lib import MpnError
    with pytest.raises(MpnError) as e:
        Detector()
    assert "no CUDA device" in str(e.value) and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multiposenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "exact_math.h" not in src, f
