"""A numpy-backed stand-in for `tensorflow.compat.v1` (+ `tensorflow.contrib.slim`), just large enough to IMPORT AND
EXECUTE the reference's own graph-construction code for the post-backbone path eagerly, in IEEE float32:

    detector/anchor_generator.py, detector/utils/box_utils.py, detector/utils/nms.py, detector/box_predictor.py
    (reshape_and_concatenate), detector/retinanet.py (get_predictions), detector/prn.py, and the blocks of create_pb.py /
    inference/detector.py that tests/golden/make_graph_goldens.py executes by line range.

TEST INFRASTRUCTURE ONLY (golden-vector generation in the build container; TensorFlow 1.15 itself cannot be installed).

What is and is not pinned by vectors made with it:
  * everything the reference spells with element-wise / shape ops (to_float, ceil, range, meshgrid, tile, stack, concat,
    sqrt, +, -, *, /, reduce_max / reduce_min, comparisons, boolean_mask, gather, pad, argmax, //, %) is evaluated by
    numpy in float32, one IEEE rounding per operation, exactly what TensorFlow's CPU kernels do for these ops -> the
    vectors pin the oracle and the device BIT FOR BIT;
  * the library kernels `tf.exp`, `tf.sigmoid`, `tf.nn.softmax`, `tf.image.non_max_suppression`,
    `tf.image.crop_and_resize`, `slim.fully_connected` are third-party arithmetic that is not in /root/reference; the shim
    routes them to callables the generator supplies (`set_kernels`), so for those the vectors pin the reference's own
    LOGIC AROUND the op (argument order, thresholds, masks, padding, layouts), not the op's last ulp.
"""
import contextlib
import sys
import types

import numpy as np


class Dim(int):
    """A static dimension: behaves like an int and has `.value` (TF1 Dimension)."""

    @property
    def value(self):
        return int(self)


class Shape(tuple):
    def as_list(self):
        return [int(d) for d in self]


class Tensor(np.ndarray):
    """An eagerly evaluated tensor: a numpy array with the few TF1 Tensor methods the reference calls."""

    def __array_finalize__(self, obj):
        pass

    @property
    def shape(self):
        return Shape(Dim(d) for d in np.ndarray.shape.__get__(self))

    def set_shape(self, shape):
        got = tuple(int(d) for d in np.ndarray.shape.__get__(self))
        want = tuple(shape)
        assert len(got) == len(want) and all(w is None or int(w) == g for g, w in zip(got, want)), (got, want)

    def get_shape(self):
        return self.shape


def T(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(Tensor)


_kernels = {}


def set_kernels(**fns):
    """exp, sigmoid, softmax, non_max_suppression, crop_and_resize, fully_connected_weights (dict scope -> (W, b))."""
    _kernels.update(fns)


@contextlib.contextmanager
def _scope(*a, **k):
    yield


_scope_stack = []


@contextlib.contextmanager
def _variable_scope(name, *a, **k):
    _scope_stack.append(name)
    try:
        yield
    finally:
        _scope_stack.pop()


def _split(value, num_or_size_splits, axis=0):
    return [T(p) for p in np.split(np.asarray(value), num_or_size_splits, axis=axis)]


def _unstack(value, axis=0):
    v = np.asarray(value)
    return [T(np.take(v, i, axis=axis).copy()) for i in range(v.shape[axis])]      # copies: the reference does `ty /= ...`


def _map_fn(fn, elems, dtype=None, **kwargs):
    n = int(np.asarray(elems[0]).shape[0])
    outs = [fn([T(np.asarray(e)[i]) for e in elems]) for i in range(n)]
    return tuple(T(np.stack([np.asarray(o[j]) for o in outs], 0)) for j in range(len(outs[0])))


def _pad(t, paddings):
    return T(np.pad(np.asarray(t), [(int(a), int(b)) for a, b in paddings]))


def _fully_connected(x, num_outputs, activation_fn=None, scope=None, **kwargs):
    W, b = _kernels["fully_connected_weights"]["/".join(_scope_stack + [scope])]
    assert W.shape[1] == num_outputs
    y = np.asarray(x, np.float32) @ W + b
    return T(activation_fn(y) if activation_fn is not None else y)


def install():
    """Put the stand-in modules into sys.modules; returns the `tensorflow.compat.v1` module object."""
    tf = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    v1 = types.ModuleType("tensorflow.compat.v1")
    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")
    tf.compat, compat.v1, tf.contrib, contrib.slim = compat, v1, contrib, slim

    f32 = np.float32
    v1.float32, v1.int32, v1.int64, v1.uint8 = np.float32, np.int32, np.int64, np.uint8
    v1.AUTO_REUSE = object()
    v1.name_scope = _scope
    v1.variable_scope = _variable_scope
    v1.to_float = lambda x: T(np.asarray(x).astype(f32))
    v1.to_int32 = lambda x: T(np.asarray(x).astype(np.int32))
    v1.ceil = lambda x: T(np.ceil(np.asarray(x)))
    v1.sqrt = lambda x: T(np.sqrt(np.asarray(x, f32)))
    v1.log = lambda x: T(np.log(np.asarray(x, f32)))
    v1.exp = lambda x: T(_kernels["exp"](np.asarray(x, f32)))
    v1.sigmoid = lambda x: T(_kernels["sigmoid"](np.asarray(x, f32)))
    v1.constant = lambda v, dtype=None, **k: T(np.asarray(v, dtype=dtype))
    v1.size = lambda x: np.int32(np.asarray(x).size)
    v1.shape = lambda x: T(np.asarray(np.asarray(x).shape, np.int32))
    v1.range = lambda *a: T(np.arange(*[int(v) for v in a], dtype=np.int32))
    v1.ones = lambda shape, dtype=f32: T(np.ones([int(s) for s in shape], dtype))
    v1.meshgrid = lambda *a, **k: [T(m) for m in np.meshgrid(*[np.asarray(x) for x in a], **k)]
    v1.stack = lambda vals, axis=0: T(np.stack([np.asarray(v) for v in vals], axis=axis))
    v1.unstack = _unstack
    v1.split = _split
    v1.concat = lambda vals, axis=0: T(np.concatenate([np.asarray(v) for v in vals], axis=axis))
    v1.expand_dims = lambda x, axis: T(np.expand_dims(np.asarray(x), axis))
    v1.tile = lambda x, m: T(np.tile(np.asarray(x), [int(v) for v in m]))
    v1.reshape = lambda x, s: T(np.reshape(np.asarray(x), [int(v) for v in s]))
    v1.transpose = lambda x, perm=None: T(np.transpose(np.asarray(x), perm))
    v1.minimum = lambda a, b: T(np.minimum(a, b))
    v1.maximum = lambda a, b: T(np.maximum(a, b))
    v1.divide = lambda a, b: T(np.asarray(a) / np.asarray(b))
    v1.clip_by_value = lambda x, lo, hi: T(np.clip(np.asarray(x), f32(lo), f32(hi)))
    v1.boolean_mask = lambda t, m: T(np.asarray(t)[np.asarray(m, bool)])
    v1.gather = lambda t, idx: T(np.asarray(t)[np.asarray(idx, np.int64)])
    v1.pad = _pad
    v1.map_fn = _map_fn
    v1.identity = lambda x, name=None: x
    v1.reduce_max = lambda x, axis=None, keepdims=False: T(np.max(np.asarray(x), axis=tuple(axis) if axis is not None else None,
                                                                  keepdims=keepdims))
    v1.reduce_min = lambda x, axis=None, keepdims=False: T(np.min(np.asarray(x), axis=tuple(axis) if axis is not None else None,
                                                                  keepdims=keepdims))
    v1.argmax = lambda x, axis=None, output_type=np.int64: T(np.argmax(np.asarray(x), axis=axis).astype(output_type))
    v1.greater_equal = lambda a, b: T(np.asarray(a) >= np.asarray(b))
    v1.variance_scaling_initializer = lambda *a, **k: None

    nn = types.ModuleType("tensorflow.compat.v1.nn")
    nn.relu = lambda x: T(np.maximum(np.asarray(x, f32), f32(0)))
    nn.softmax = lambda x, axis=-1: T(_kernels["softmax"](np.asarray(x, f32), axis))
    v1.nn = nn

    image = types.ModuleType("tensorflow.compat.v1.image")
    image.non_max_suppression = lambda boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float("-inf"): T(
        _kernels["non_max_suppression"](np.asarray(boxes, f32), np.asarray(scores, f32), int(max_output_size),
                                        float(iou_threshold), float(score_threshold)))
    image.crop_and_resize = lambda img, boxes, box_ind, crop_size, **k: T(
        _kernels["crop_and_resize"](np.asarray(img, f32), np.asarray(boxes, f32), np.asarray(box_ind, np.int32),
                                    tuple(int(c) for c in crop_size)))

    class ResizeMethod:
        BILINEAR = 0
    image.ResizeMethod = ResizeMethod
    v1.image = image

    slim.fully_connected = _fully_connected

    @contextlib.contextmanager
    def arg_scope(*a, **k):
        yield
    slim.arg_scope = arg_scope
    slim.dropout = lambda x, **k: x

    for name, mod in (("tensorflow", tf), ("tensorflow.compat", compat), ("tensorflow.compat.v1", v1),
                      ("tensorflow.contrib", contrib), ("tensorflow.contrib.slim", slim)):
        sys.modules[name] = mod
    return v1
