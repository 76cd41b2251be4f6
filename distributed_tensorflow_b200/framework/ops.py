"""Op library: symbolic builders + their evaluation kernels.

Every builder creates one graph node; every node type has a kernel registered
with :func:`register_kernel`.  Kernels receive already-evaluated input values
(``torch.Tensor``) and run on the node's resolved torch device.  On CUDA
devices the GEMM / softmax-cross-entropy / optimizer kernels route to the
hand-written sm_100a kernels in ``ops/native.py`` (tcgen05 GEMM, fused xent,
fused apply); on CPU they use plain PyTorch so the whole API is testable
without a GPU.

Op coverage follows SURVEY §2.4 (K1-K14): ``xw_plus_b``, ``relu``, ``softmax``,
``clip_by_value``, ``log``, ``reduce_sum`` (reference ``distributed_mnist.py:109-113``),
``multiply``/``square``/``reduce_mean`` (``example_between_graph.py:57-59``),
``split``/``concat``/``matmul`` (``example_in_graph.py:38,56,58``),
``expand_dims`` (``standalone.py:83-87``), ``arg_max``
(``distributed_mnist_predict.py:33``), plus conv/BN/pool for the ResNet-18
configuration named in BASELINE.json.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import shapes as _sh
from .graph import Tensor, convert_to_tensor, get_default_graph

KERNELS: Dict[str, Callable] = {}
# op types whose kernels mutate task-local resources (run under no_grad, never cached across runs)
STATEFUL_OPS = set()

float32, float64, float16, bfloat16 = torch.float32, torch.float64, torch.float16, torch.bfloat16
int32, int64, uint8, int8, bool_ = torch.int32, torch.int64, torch.uint8, torch.int8, torch.bool


def register_kernel(op_type: str, stateful: bool = False):
    def deco(fn):
        KERNELS[op_type] = fn
        if stateful:
            STATEFUL_OPS.add(op_type)
        return fn
    return deco


def as_dtype(dt) -> Optional[torch.dtype]:
    if dt is None or isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str):
        return getattr(torch, dt.replace("torch.", ""))
    if dt in (float,):
        return torch.float32
    if dt in (int,):
        return torch.int64
    if dt in (bool,):
        return torch.bool
    return torch.from_numpy(np.zeros((), dtype=dt)).dtype


def _node(op_type, inputs=(), attrs=None, name=None, dtype=None, shape=None, **kw) -> Tensor:
    ins = [convert_to_tensor(i) for i in inputs]
    return get_default_graph().create_node(op_type, ins, attrs or {}, name or op_type, dtype, shape, **kw)


def _to_torch(value, dtype=None) -> torch.Tensor:
    if isinstance(value, torch.Tensor):
        t = value
    else:
        arr = value if isinstance(value, np.ndarray) else np.asarray(value)
        if arr.dtype == np.float64 and dtype is None and not isinstance(value, np.ndarray):
            arr = arr.astype(np.float32)       # python floats: TF default float is float32
        if arr.ndim == 0:
            t = torch.tensor(arr.item(), dtype=torch.from_numpy(np.zeros(1, arr.dtype)).dtype)
        else:
            t = torch.from_numpy(arr if arr.flags["C_CONTIGUOUS"] and arr.flags["WRITEABLE"] else arr.copy())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


# ---------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name="Const") -> Tensor:
    t = _to_torch(value, as_dtype(dtype))
    if shape is not None:
        shape = tuple(shape)
        t = t.expand(shape).contiguous() if t.numel() == 1 else t.reshape(shape)
    return _node("Const", (), {"value": t}, name, t.dtype, tuple(t.shape))


@register_kernel("Const")
def _k_const(ctx, node):
    return node.attrs["value"].to(ctx.torch_device(node))


def placeholder(dtype=float32, shape=None, name="Placeholder") -> Tensor:
    return _node("Placeholder", (), {}, name, as_dtype(dtype), None if shape is None else tuple(shape))


@register_kernel("Placeholder")
def _k_placeholder(ctx, node):
    raise ValueError("You must feed a value for placeholder tensor %r" % node.name)


def placeholder_with_default(input, shape=None, name="PlaceholderWithDefault") -> Tensor:
    x = convert_to_tensor(input)
    return _node("PlaceholderWithDefault", (x,), {}, name, x.dtype, shape if shape is not None else x.shape)


@register_kernel("PlaceholderWithDefault")
def _k_pwd(ctx, node, x):
    return x


def _fill(op, shape, dtype, name, **attrs):
    shape = tuple(int(s) for s in (shape if isinstance(shape, (list, tuple)) else [shape]))
    dt = as_dtype(dtype) or float32
    return _node(op, (), dict(attrs, shape=shape, dtype=dt), name, dt, shape)


def zeros(shape, dtype=float32, name="zeros") -> Tensor:
    return _fill("Zeros", shape, dtype, name)


def ones(shape, dtype=float32, name="ones") -> Tensor:
    return _fill("Ones", shape, dtype, name)


def fill(dims, value, name="Fill") -> Tensor:
    return _fill("Fill", dims, float32 if isinstance(value, float) else None, name, value=value)


@register_kernel("Zeros")
def _k_zeros(ctx, node):
    return torch.zeros(node.attrs["shape"], dtype=node.attrs["dtype"], device=ctx.torch_device(node))


@register_kernel("Ones")
def _k_ones(ctx, node):
    return torch.ones(node.attrs["shape"], dtype=node.attrs["dtype"], device=ctx.torch_device(node))


@register_kernel("Fill")
def _k_fill(ctx, node):
    return torch.full(node.attrs["shape"], node.attrs["value"], dtype=node.attrs["dtype"],
                      device=ctx.torch_device(node))


def zeros_like(x, name="zeros_like") -> Tensor:
    x = convert_to_tensor(x)
    return _node("ZerosLike", (x,), {}, name, x.dtype, x.shape)


@register_kernel("ZerosLike")
def _k_zl(ctx, node, x):
    return torch.zeros_like(x)


def ones_like(x, name="ones_like") -> Tensor:
    x = convert_to_tensor(x)
    return _node("OnesLike", (x,), {}, name, x.dtype, x.shape)


@register_kernel("OnesLike")
def _k_ol(ctx, node, x):
    return torch.ones_like(x)


# -- random sources (K13) -----------------------------------------------------
def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name="truncated_normal") -> Tensor:
    return _fill("TruncatedNormal", shape, dtype, name, mean=float(mean), stddev=float(stddev), seed=seed)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name="random_normal") -> Tensor:
    return _fill("RandomNormal", shape, dtype, name, mean=float(mean), stddev=float(stddev), seed=seed)


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None, name="random_uniform") -> Tensor:
    return _fill("RandomUniform", shape, dtype, name, minval=float(minval), maxval=float(maxval), seed=seed)


class _OpStream:
    """The random stream of ONE op: a Philox key + the next unused block (``csrc/philox.h``)."""
    __slots__ = ("key", "offset", "lock")

    def __init__(self, key: int):
        import threading
        self.key, self.offset, self.lock = key & (2 ** 64 - 1), 0, threading.Lock()

    def take(self, blocks: int) -> int:
        with self.lock:
            at = self.offset
            self.offset += int(blocks)
            return at


def _op_stream(ctx, node) -> _OpStream:
    """TF-style per-op stream: keyed by (graph-level seed, op-level seed, the op's identity) and STATEFUL -- it lives in the
    executing task's resource store, so every execution of the op continues it (a seeded op does not return the same tensor
    every step), two same-shaped initialisers draw different values even when they sit on different tasks, and a fresh
    store (a new local Session, a restarted ps) replays the same sequence.  With neither seed the key comes from entropy.
    The stream is counter-based (Philox4x32-10): the same (key, offset) yields the same values on a CPU task and on a GPU
    task (``ops/random_ops.py``)."""
    import os as _os
    import zlib
    op_seed = node.attrs.get("seed")
    graph_seed = ctx._seed if getattr(ctx, "_seed", None) is not None else getattr(node.graph, "seed", None)
    store = ctx.store
    with store._lock:
        streams = store.__dict__.setdefault("_rng_streams", {})
        key = node.name if node.name else "node%d" % node.id
        st = streams.get(key)
        if st is None:
            if op_seed is None and graph_seed is None:
                k = int.from_bytes(_os.urandom(8), "little")
            else:
                k = (int(graph_seed or 0) * 1000003) ^ (int(op_seed or 0) * 7919 + (1 if op_seed is not None else 0)) \
                    ^ (zlib.crc32(key.encode()) << 17)
                k = (k * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & (2 ** 64 - 1)      # spread small seeds over the key bits
            st = streams[key] = _OpStream(k)
    return st


def _random_fill(ctx, node, kind: int, p0: float, p1: float, shape=None, device=None) -> torch.Tensor:
    from ..ops import random_ops
    shape = tuple(node.attrs["shape"]) if shape is None else tuple(shape)
    st = _op_stream(ctx, node)
    at = st.take(random_ops.blocks_used(shape))
    return random_ops.philox_fill(shape, kind, p0, p1, st.key, at, device if device is not None else ctx.torch_device(node))


@register_kernel("TruncatedNormal")
def _k_tn(ctx, node):
    a = node.attrs
    return _random_fill(ctx, node, 2, a["mean"], a["stddev"]).to(dtype=a["dtype"])


@register_kernel("RandomNormal")
def _k_rn(ctx, node):
    a = node.attrs
    return _random_fill(ctx, node, 1, a["mean"], a["stddev"]).to(dtype=a["dtype"])


@register_kernel("RandomUniform")
def _k_ru(ctx, node):
    a = node.attrs
    return _random_fill(ctx, node, 0, a["minval"], a["maxval"]).to(dtype=a["dtype"])


# ---------------------------------------------------------------------------
# element-wise / math
# ---------------------------------------------------------------------------
def _binary(op, a, b, name):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node(op, (a, b), {}, name, a.dtype or b.dtype, _sh.broadcast_shape(a.shape, b.shape))


def _unary(op, x, name, **attrs):
    x = convert_to_tensor(x)
    return _node(op, (x,), attrs, name, x.dtype, x.shape)


def add(a, b, name="Add"): return _binary("Add", a, b, name)
def subtract(a, b, name="Sub"): return _binary("Sub", a, b, name)
def multiply(a, b, name="Mul"): return _binary("Mul", a, b, name)
def divide(a, b, name="RealDiv"): return _binary("RealDiv", a, b, name)
def maximum(a, b, name="Maximum"): return _binary("Maximum", a, b, name)
def minimum(a, b, name="Minimum"): return _binary("Minimum", a, b, name)
def pow(a, b, name="Pow"): return _binary("Pow", a, b, name)
def squared_difference(a, b, name="SquaredDifference"): return _binary("SquaredDifference", a, b, name)
def negative(x, name="Neg"): return _unary("Neg", x, name)
def square(x, name="Square"): return _unary("Square", x, name)
def sqrt(x, name="Sqrt"): return _unary("Sqrt", x, name)
def rsqrt(x, name="Rsqrt"): return _unary("Rsqrt", x, name)
def exp(x, name="Exp"): return _unary("Exp", x, name)
def log(x, name="Log"): return _unary("Log", x, name)
def abs(x, name="Abs"): return _unary("Abs", x, name)
def sigmoid(x, name="Sigmoid"): return _unary("Sigmoid", x, name)
def tanh(x, name="Tanh"): return _unary("Tanh", x, name)
def relu(x, name="Relu"): return _unary("Relu", x, name)
def identity(x, name="Identity"): return _unary("Identity", x, name)
def stop_gradient(x, name="StopGradient"): return _unary("StopGradient", x, name)


sub, mul, div, neg = subtract, multiply, divide, negative


def _native_ew():
    """K9: fp32 tensors on /gpu run the ew_* kernels of csrc/elementwise.cu (forward and backward), see ops/native.py."""
    from ..ops import native
    return native

register_kernel("Add")(lambda ctx, n, a, b: _native_ew().binary("add", a, b))
register_kernel("Sub")(lambda ctx, n, a, b: _native_ew().binary("sub", a, b))
register_kernel("Mul")(lambda ctx, n, a, b: _native_ew().binary("mul", a, b))
register_kernel("RealDiv")(lambda ctx, n, a, b: _native_ew().binary("div", a, b))
register_kernel("Maximum")(lambda ctx, n, a, b: torch.maximum(a, b))
register_kernel("Minimum")(lambda ctx, n, a, b: torch.minimum(a, b))
register_kernel("Pow")(lambda ctx, n, a, b: torch.pow(a, b))
register_kernel("SquaredDifference")(lambda ctx, n, a, b: _native_ew().binary("sqdiff", a, b))
register_kernel("Neg")(lambda ctx, n, x: _native_ew().unary("neg", x))
register_kernel("Square")(lambda ctx, n, x: _native_ew().unary("square", x))
register_kernel("Sqrt")(lambda ctx, n, x: _native_ew().unary("sqrt", x))
register_kernel("Rsqrt")(lambda ctx, n, x: torch.rsqrt(x))
register_kernel("Exp")(lambda ctx, n, x: _native_ew().unary("exp", x))
register_kernel("Log")(lambda ctx, n, x: _native_ew().unary("log", x))
register_kernel("Abs")(lambda ctx, n, x: torch.abs(x))
register_kernel("Sigmoid")(lambda ctx, n, x: _native_ew().unary("sigmoid", x))
register_kernel("Tanh")(lambda ctx, n, x: _native_ew().unary("tanh", x))
register_kernel("Relu")(lambda ctx, n, x: _native_ew().unary("relu", x))
register_kernel("Identity")(lambda ctx, n, x: x)
register_kernel("StopGradient")(lambda ctx, n, x: x.detach())


def add_n(inputs, name="AddN") -> Tensor:
    ins = [convert_to_tensor(i) for i in inputs]
    return _node("AddN", ins, {}, name, ins[0].dtype, ins[0].shape)


@register_kernel("AddN")
def _k_addn(ctx, node, *xs):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def clip_by_value(t, clip_value_min, clip_value_max, name="clip_by_value") -> Tensor:
    x = convert_to_tensor(t)
    return _node("ClipByValue", (x,), {"lo": float(clip_value_min), "hi": float(clip_value_max)}, name, x.dtype, x.shape)


@register_kernel("ClipByValue")
def _k_clip(ctx, node, x):
    # gradient is gated outside [lo, hi], as in TF (clip at 1e-10 gates the gradient)
    return torch.clamp(x, node.attrs["lo"], node.attrs["hi"])


def cast(x, dtype, name="Cast") -> Tensor:
    x = convert_to_tensor(x)
    dt = as_dtype(dtype)
    return _node("Cast", (x,), {"dtype": dt}, name, dt, x.shape)


to_float = lambda x, name="ToFloat": cast(x, float32, name)

register_kernel("Cast")(lambda ctx, n, x: x.to(n.attrs["dtype"]))


def equal(a, b, name="Equal"):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node("Equal", (a, b), {}, name, bool_, _sh.broadcast_shape(a.shape, b.shape))


def greater(a, b, name="Greater"):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node("Greater", (a, b), {}, name, bool_, _sh.broadcast_shape(a.shape, b.shape))


def less(a, b, name="Less"):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node("Less", (a, b), {}, name, bool_, _sh.broadcast_shape(a.shape, b.shape))


register_kernel("Equal")(lambda ctx, n, a, b: a == b)
register_kernel("Greater")(lambda ctx, n, a, b: a > b)
register_kernel("Less")(lambda ctx, n, a, b: a < b)


# ---------------------------------------------------------------------------
# reductions / indexing
# ---------------------------------------------------------------------------
def _axis_attr(axis, reduction_indices=None):
    if axis is None:
        axis = reduction_indices
    if axis is None:
        return None
    return int(axis) if isinstance(axis, (int, np.integer)) else tuple(int(a) for a in axis)


def _reduce(op, x, axis, keepdims, name, reduction_indices=None, keep_dims=None):
    x = convert_to_tensor(x)
    if keep_dims is not None:           # TF-1.x spelling used at standalone.py:87
        keepdims = keep_dims
    ax = _axis_attr(axis, reduction_indices)
    return _node(op, (x,), {"axis": ax, "keepdims": bool(keepdims)}, name, x.dtype,
                 _sh.reduce_shape(x.shape, ax, bool(keepdims)))


def reduce_sum(x, axis=None, keepdims=False, name="Sum", reduction_indices=None, keep_dims=None):
    return _reduce("Sum", x, axis, keepdims, name, reduction_indices, keep_dims)


def reduce_mean(x, axis=None, keepdims=False, name="Mean", reduction_indices=None, keep_dims=None):
    return _reduce("Mean", x, axis, keepdims, name, reduction_indices, keep_dims)


def reduce_max(x, axis=None, keepdims=False, name="Max", reduction_indices=None, keep_dims=None):
    return _reduce("Max", x, axis, keepdims, name, reduction_indices, keep_dims)


def reduce_min(x, axis=None, keepdims=False, name="Min", reduction_indices=None, keep_dims=None):
    return _reduce("Min", x, axis, keepdims, name, reduction_indices, keep_dims)


def _red(fn_all, fn_dim):
    def k(ctx, node, x):
        ax, kd = node.attrs["axis"], node.attrs["keepdims"]
        if ax is None:
            out = fn_all(x)
            return out.reshape([1] * x.dim()) if kd else out
        return fn_dim(x, ax, kd)
    return k


register_kernel("Sum")(_red(lambda x: _native_ew().reduce_all(x, False), lambda x, a, k: x.sum(dim=a, keepdim=k)))
def _mean_over(x, a, k):
    # NHWC global average pooling (mean over H, W of a 4-D tensor, ResNet's classifier input): our pooling kernel on /gpu
    # when the fused NN kernels are enabled; plain mean otherwise
    if x.dim() == 4 and not k and sorted(int(i) % 4 for i in (a if isinstance(a, (list, tuple)) else [a])) == [1, 2]:
        from ..ops import native
        return native.global_avg_pool(x).to(x.dtype)
    return x.mean(dim=a, keepdim=k)


register_kernel("Mean")(_red(lambda x: _native_ew().reduce_all(x, True), _mean_over))
register_kernel("Max")(_red(lambda x: x.max(), lambda x, a, k: x.amax(dim=a, keepdim=k)))
register_kernel("Min")(_red(lambda x: x.min(), lambda x, a, k: x.amin(dim=a, keepdim=k)))


def argmax(x, axis=None, name="ArgMax", dimension=None, output_type=int64):
    x = convert_to_tensor(x)
    ax = dimension if axis is None else axis
    ax = 0 if ax is None else int(ax)
    return _node("ArgMax", (x,), {"axis": ax}, name, int64, _sh.reduce_shape(x.shape, ax, False))


arg_max = argmax     # distributed_mnist_predict.py:33 spelling

register_kernel("ArgMax")(lambda ctx, n, x: torch.argmax(x, dim=n.attrs["axis"]))


def reshape(x, shape, name="Reshape"):
    x = convert_to_tensor(x)
    shape = tuple(int(s) for s in shape)
    st = tuple(None if s == -1 else s for s in shape)
    return _node("Reshape", (x,), {"shape": shape}, name, x.dtype, st)


register_kernel("Reshape")(lambda ctx, n, x: x.reshape(n.attrs["shape"]))


def transpose(x, perm=None, name="Transpose"):
    x = convert_to_tensor(x)
    shp = None
    if x.shape is not None:
        p = perm if perm is not None else list(reversed(range(len(x.shape))))
        shp = tuple(x.shape[i] for i in p)
    return _node("Transpose", (x,), {"perm": None if perm is None else tuple(perm)}, name, x.dtype, shp)


@register_kernel("Transpose")
def _k_transpose(ctx, node, x):
    perm = node.attrs["perm"]
    return x.permute(*(perm if perm is not None else reversed(range(x.dim()))))


def expand_dims(x, axis=None, name="ExpandDims", dim=None):
    x = convert_to_tensor(x)
    ax = int(dim if axis is None else axis)
    shp = None
    if x.shape is not None:
        l = list(x.shape)
        l.insert(ax if ax >= 0 else len(l) + ax + 1, 1)
        shp = tuple(l)
    return _node("ExpandDims", (x,), {"axis": ax}, name, x.dtype, shp)


register_kernel("ExpandDims")(lambda ctx, n, x: x.unsqueeze(n.attrs["axis"]))


def squeeze(x, axis=None, name="Squeeze"):
    x = convert_to_tensor(x)
    return _node("Squeeze", (x,), {"axis": axis}, name, x.dtype, None)


@register_kernel("Squeeze")
def _k_squeeze(ctx, node, x):
    ax = node.attrs["axis"]
    return x.squeeze() if ax is None else x.squeeze(ax)


def concat(values, axis=0, name="concat"):
    ins = [convert_to_tensor(v) for v in values]
    shp = None
    if all(i.shape is not None for i in ins):
        shp = list(ins[0].shape)
        tot = 0
        for i in ins:
            d = i.shape[axis]
            tot = None if (tot is None or d is None) else tot + d
        shp[axis] = tot
        shp = tuple(shp)
    return _node("ConcatV2", ins, {"axis": int(axis)}, name, ins[0].dtype, shp)


register_kernel("ConcatV2")(lambda ctx, n, *xs: _native_ew().concat(xs, n.attrs["axis"]))


def stack(values, axis=0, name="stack"):
    ins = [convert_to_tensor(v) for v in values]
    return _node("Pack", ins, {"axis": int(axis)}, name, ins[0].dtype, None)


register_kernel("Pack")(lambda ctx, n, *xs: torch.stack(xs, dim=n.attrs["axis"]))


def split(value, num_or_size_splits, axis=0, name="split") -> List[Tensor]:
    """K10: returns a python list of slice nodes (one per part)."""
    x = convert_to_tensor(value)
    outs = []
    if isinstance(num_or_size_splits, int):
        n = num_or_size_splits
        for i in range(n):
            shp = None
            if x.shape is not None:
                l = list(x.shape)
                l[axis] = None if l[axis] is None else l[axis] // n
                shp = tuple(l)
            outs.append(_node("SplitPart", (x,), {"axis": int(axis), "num": n, "index": i}, "%s_%d" % (name, i),
                              x.dtype, shp))
    else:
        sizes = [int(s) for s in num_or_size_splits]
        off = 0
        for i, s in enumerate(sizes):
            shp = None
            if x.shape is not None:
                l = list(x.shape)
                l[axis] = s
                shp = tuple(l)
            outs.append(_node("SplitPart", (x,), {"axis": int(axis), "offset": off, "size": s},
                              "%s_%d" % (name, i), x.dtype, shp))
            off += s
    return outs


@register_kernel("SplitPart")
def _k_split(ctx, node, x):
    a = node.attrs
    ax = a["axis"]
    if "num" in a:
        if x.shape[ax] % a["num"]:
            raise ValueError("split: dimension %d (size %d) not divisible by %d" % (ax, x.shape[ax], a["num"]))
        size = x.shape[ax] // a["num"]
        off = a["index"] * size
    else:
        off, size = a["offset"], a["size"]
    return x.narrow(ax, off, size)


def slice_rows(x, begin, size, name="Slice"):
    x = convert_to_tensor(x)
    return _node("SplitPart", (x,), {"axis": 0, "offset": int(begin), "size": int(size)}, name, x.dtype, None)


def one_hot(indices, depth, dtype=float32, name="one_hot"):
    x = convert_to_tensor(indices)
    return _node("OneHot", (x,), {"depth": int(depth), "dtype": as_dtype(dtype)}, name, as_dtype(dtype),
                 None if x.shape is None else tuple(x.shape) + (int(depth),))


register_kernel("OneHot")(lambda ctx, n, x: F.one_hot(x.long(), n.attrs["depth"]).to(n.attrs["dtype"]))


# ---------------------------------------------------------------------------
# linear algebra / nn (K1, K2, K3, K11, K14)
# ---------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False, name="MatMul"):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node("MatMul", (a, b), {"ta": bool(transpose_a), "tb": bool(transpose_b)}, name, a.dtype,
                 _sh.matmul_shape(a.shape, b.shape, transpose_a, transpose_b))


@register_kernel("MatMul")
def _k_matmul(ctx, node, a, b):
    from ..ops import native
    return native.matmul(a, b, node.attrs["ta"], node.attrs["tb"])


def bias_add(value, bias, name="BiasAdd"):
    return _binary("Add", value, bias, name)


def xw_plus_b(x, weights, biases, name="xw_plus_b"):
    """K1/K2: ``x @ W + b`` as ONE node so the GPU path runs a single GEMM with bias epilogue."""
    x, w, b = convert_to_tensor(x), convert_to_tensor(weights), convert_to_tensor(biases)
    return _node("XwPlusB", (x, w, b), {"relu": False}, name, x.dtype, _sh.matmul_shape(x.shape, w.shape))


@register_kernel("XwPlusB")
def _k_xwb(ctx, node, x, w, b):
    from ..ops import native
    return native.linear(x, w, b, relu=node.attrs.get("relu", False))


def softmax(logits, axis=-1, name="Softmax"):
    x = convert_to_tensor(logits)
    return _node("Softmax", (x,), {"axis": int(axis)}, name, x.dtype, x.shape)


def log_softmax(logits, axis=-1, name="LogSoftmax"):
    x = convert_to_tensor(logits)
    return _node("LogSoftmax", (x,), {"axis": int(axis)}, name, x.dtype, x.shape)


register_kernel("Softmax")(lambda ctx, n, x: torch.softmax(x, dim=n.attrs["axis"]))
register_kernel("LogSoftmax")(lambda ctx, n, x: torch.log_softmax(x, dim=n.attrs["axis"]))


def softmax_cross_entropy_with_logits(labels=None, logits=None, name="SoftmaxCrossEntropyWithLogits"):
    l, y = convert_to_tensor(logits), convert_to_tensor(labels)
    return _node("SoftmaxXent", (l, y), {}, name, l.dtype, None if l.shape is None else l.shape[:-1])


@register_kernel("SoftmaxXent")
def _k_sxent(ctx, node, logits, labels):
    from ..ops import native
    return native.softmax_xent(logits, labels)


def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None, name="SparseSoftmaxXent"):
    l, y = convert_to_tensor(logits), convert_to_tensor(labels)
    return _node("SparseSoftmaxXent", (l, y), {}, name, l.dtype, None if l.shape is None else l.shape[:-1])


register_kernel("SparseSoftmaxXent")(
    lambda ctx, n, logits, labels: F.cross_entropy(logits.float(), labels.long(), reduction="none"))


def clipped_softmax_xent_sum(logits, labels, clip_min=1e-10, name="ClippedSoftmaxXentSum"):
    """The reference loss as one fused node: ``-sum(y_ * log(clip(softmax(logits), 1e-10, 1)))``
    (``distributed_mnist.py:112-113``).  Produced by the remapper or called directly."""
    l, y = convert_to_tensor(logits), convert_to_tensor(labels)
    return _node("ClippedSoftmaxXentSum", (l, y), {"clip_min": float(clip_min)}, name, l.dtype, ())


@register_kernel("ClippedSoftmaxXentSum")
def _k_cxent(ctx, node, logits, labels):
    from ..ops import native
    return native.clipped_softmax_xent_sum(logits, labels, node.attrs["clip_min"])


def l2_loss(t, name="L2Loss"):
    x = convert_to_tensor(t)
    return _node("L2Loss", (x,), {}, name, x.dtype, ())


register_kernel("L2Loss")(lambda ctx, n, x: 0.5 * (x * x).sum())


def conv2d(input, filter, strides=(1, 1, 1, 1), padding="SAME", data_format="NHWC", name="Conv2D"):
    """K14. ``filter`` is HWIO like TF; data is NHWC (channels-last is also the tensor-core layout)."""
    x, w = convert_to_tensor(input), convert_to_tensor(filter)
    st = tuple(int(s) for s in strides)
    shp = None
    if x.shape is not None and w.shape is not None and len(x.shape) == 4 and len(w.shape) == 4 and data_format == "NHWC":
        shp = (x.shape[0], _window_out(x.shape[1], w.shape[0], st[1], padding), _window_out(x.shape[2], w.shape[1], st[2], padding),
               w.shape[3])
    return _node("Conv2D", (x, w), {"strides": st, "padding": padding, "data_format": data_format}, name, x.dtype, shp)


def _window_out(size, k, s, padding):
    """Static output extent of a sliding window (TF's SAME / VALID arithmetic); ``None`` stays unknown."""
    if size is None or k is None:
        return None
    size, k, s = int(size), int(k), int(s)
    return -(-size // s) if str(padding).upper() == "SAME" else max(0, (size - k) // s + 1)


def _pool_shape(x, ksize, strides, padding):
    if x.shape is None or len(x.shape) != 4:
        return None
    return (x.shape[0], _window_out(x.shape[1], ksize[1], strides[1], padding), _window_out(x.shape[2], ksize[2], strides[2], padding),
            x.shape[3])


@register_kernel("Conv2D")
def _k_conv(ctx, node, x, w):
    from ..ops import native
    a = node.attrs
    return native.conv2d_nhwc(x, w, a["strides"], a["padding"], a["data_format"])


def _pool_attrs(ksize, strides, padding):
    return {"ksize": tuple(int(k) for k in ksize), "strides": tuple(int(s) for s in strides), "padding": padding}


def max_pool(value, ksize, strides, padding="SAME", name="MaxPool"):
    x = convert_to_tensor(value)
    return _node("MaxPool", (x,), _pool_attrs(ksize, strides, padding), name, x.dtype, _pool_shape(x, ksize, strides, padding))


def avg_pool(value, ksize, strides, padding="SAME", name="AvgPool"):
    x = convert_to_tensor(value)
    return _node("AvgPool", (x,), _pool_attrs(ksize, strides, padding), name, x.dtype, _pool_shape(x, ksize, strides, padding))


def _same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _pool(x, attrs, mode):
    kh, kw = attrs["ksize"][1:3]
    sh, sw = attrs["strides"][1:3]
    xn = x.permute(0, 3, 1, 2)
    if attrs["padding"].upper() == "SAME":
        pt, pb = _same_pad(xn.shape[2], kh, sh)
        pl, pr = _same_pad(xn.shape[3], kw, sw)
        if pt or pb or pl or pr:
            xn = F.pad(xn, (pl, pr, pt, pb), value=float("-inf") if mode == "max" else 0.0)
    out = F.max_pool2d(xn, (kh, kw), (sh, sw)) if mode == "max" else F.avg_pool2d(xn, (kh, kw), (sh, sw))
    return out.permute(0, 2, 3, 1)


def _k_maxpool(ctx, n, x):
    from ..ops import native                       # our NHWC kernels on /gpu when enabled, F.max_pool2d otherwise
    return native.max_pool_nhwc(x, n.attrs["ksize"], n.attrs["strides"], n.attrs["padding"])


register_kernel("MaxPool")(_k_maxpool)
register_kernel("AvgPool")(lambda ctx, n, x: _pool(x, n.attrs, "avg"))


def batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name="batchnorm"):
    ins = [convert_to_tensor(v) for v in (x, mean, variance, offset, scale)]
    return _node("BatchNorm", ins, {"eps": float(variance_epsilon)}, name, ins[0].dtype, ins[0].shape)


register_kernel("BatchNorm")(
    lambda ctx, n, x, m, v, o, s: (x - m) * torch.rsqrt(v + n.attrs["eps"]) * s + o)


def moments(x, axes, keepdims=False, name="moments"):
    x = convert_to_tensor(x)
    mean = reduce_mean(x, axis=list(axes), keepdims=True, name=name + "/mean")
    var = reduce_mean(squared_difference(x, stop_gradient(mean)), axis=list(axes), keepdims=keepdims,
                      name=name + "/variance")
    if not keepdims:
        mean = reduce_mean(x, axis=list(axes), keepdims=False, name=name + "/mean_squeezed")
    return mean, var


def fused_batch_norm_train(x, scale, offset, eps=1e-5, name="FusedBatchNorm"):
    """Batch-statistics normalisation over N,H,W of an NHWC tensor (training mode)."""
    ins = [convert_to_tensor(v) for v in (x, scale, offset)]
    return _node("FusedBatchNormTrain", ins, {"eps": float(eps)}, name, ins[0].dtype, ins[0].shape)


@register_kernel("FusedBatchNormTrain")
def _k_fbn(ctx, node, x, scale, offset):
    # ops/native.py: the fused statistics + apply kernels on /gpu when enabled (DTF_FUSED_NN=1), the plain formulation otherwise
    from ..ops import native
    return native.batch_norm_train(x, scale.float(), offset.float(), eps=node.attrs["eps"]).to(x.dtype)


def dropout(x, keep_prob=None, rate=None, name="dropout", seed=None, noise_shape=None):
    x = convert_to_tensor(x)
    r = float(rate) if rate is not None else 1.0 - float(keep_prob)
    return _node("Dropout", (x,), {"rate": r, "seed": seed}, name, x.dtype, x.shape)


@register_kernel("Dropout")
def _k_dropout(ctx, n, x):
    r = n.attrs["rate"]
    seeded = n.attrs.get("seed") is not None or getattr(ctx, "_seed", None) is not None or getattr(n.graph, "seed", None) is not None
    if not seeded or r <= 0.0:
        return F.dropout(x, r, training=True)
    # reproducible mask from the op's own counter-based stream (CPU and GPU tasks draw the same mask), inverted-dropout scaling
    keep = (_random_fill(ctx, n, 0, 0.0, 1.0, shape=x.shape, device=x.device) >= r).to(dtype=x.dtype)
    return x * keep / (1.0 - r)


# ---------------------------------------------------------------------------
# control / grouping
# ---------------------------------------------------------------------------
def no_op(name="NoOp"):
    return _node("NoOp", (), {}, name)


register_kernel("NoOp")(lambda ctx, n: None)


def group(*inputs, name="group_deps"):
    flat = []
    for i in inputs:
        if isinstance(i, (list, tuple)):
            flat.extend(i)
        elif i is not None:
            flat.append(i)
    g = get_default_graph()
    deps = [getattr(i, "_node", i) for i in flat]
    node = g.create_node("NoOp", [], {}, name)
    node.control_inputs = list(node.control_inputs) + deps
    return node


def tuple_(tensors, control_inputs=None, name="tuple"):
    outs = []
    deps = list(control_inputs or [])
    for t in tensors:
        n = identity(t, name=name)
        n.control_inputs = list(n.control_inputs) + [getattr(d, "_node", d) for d in deps]
        outs.append(n)
    return outs


def with_dependencies(dependencies, output_tensor, name="with_deps"):
    n = identity(output_tensor, name=name)
    n.control_inputs = list(n.control_inputs) + [getattr(d, "_node", d) for d in dependencies]
    return n


def shape(x, name="Shape"):
    x = convert_to_tensor(x)
    return _node("Shape", (x,), {}, name, int64, None)


register_kernel("Shape")(lambda ctx, n, x: torch.tensor(list(x.shape), dtype=torch.int64))


# ---------------------------------------------------------------------------
# gradients
# ---------------------------------------------------------------------------
def gradients(ys, xs, grad_ys=None, name="gradients", colocate_with=None) -> List[Optional[Tensor]]:
    """Symbolic ``d(sum ys)/d xs``.  Evaluated with reverse-mode autodiff over the
    forward values of the run (one backward pass shared by all returned nodes)."""
    single = not isinstance(xs, (list, tuple))
    xs_l = [xs] if single else list(xs)
    ys_l = list(ys) if isinstance(ys, (list, tuple)) else [ys]
    y_nodes = [convert_to_tensor(y) for y in ys_l]
    x_nodes = [convert_to_tensor(x) for x in xs_l]
    g = get_default_graph()
    dev = y_nodes[0].device
    group_node = g.create_node("Gradients", y_nodes + x_nodes, {"num_ys": len(y_nodes)}, name, device=dev)
    outs = []
    for i, x in enumerate(x_nodes):
        outs.append(g.create_node("GradientPart", [group_node], {"index": i}, "%s/%s_grad" % (name, x.name.split("/")[-1]),
                                  x.dtype, x.shape, device=dev))
    return outs


@register_kernel("Gradients")
def _k_gradients(ctx, node, *vals):
    ny = node.attrs["num_ys"]
    ys, xs = vals[:ny], vals[ny:]
    y = ys[0] if ny == 1 else sum(v.sum() for v in ys)
    if y.dim() > 0:
        y = y.sum()
    if not y.requires_grad:
        return [None] * len(xs)
    idx = [i for i, x in enumerate(xs) if isinstance(x, torch.Tensor) and x.requires_grad]
    got = torch.autograd.grad(y, [xs[i] for i in idx], retain_graph=True, allow_unused=True) if idx else ()
    out: List[Optional[torch.Tensor]] = [None] * len(xs)
    for i, gval in zip(idx, got):
        out[i] = gval
    return out


@register_kernel("GradientPart")
def _k_gradpart(ctx, node, grads):
    g = grads[node.attrs["index"]]
    return None if g is None else g.detach()


# ---------------------------------------------------------------------------
# python operators on symbolic tensors
# ---------------------------------------------------------------------------
def _install_operators():
    T = Tensor
    T.__add__ = lambda a, b: add(a, b)
    T.__radd__ = lambda a, b: add(b, a)
    T.__sub__ = lambda a, b: subtract(a, b)
    T.__rsub__ = lambda a, b: subtract(b, a)
    T.__mul__ = lambda a, b: multiply(a, b)
    T.__rmul__ = lambda a, b: multiply(b, a)
    T.__truediv__ = lambda a, b: divide(a, b)
    T.__rtruediv__ = lambda a, b: divide(b, a)
    T.__neg__ = lambda a: negative(a)
    T.__pow__ = lambda a, b: pow(a, b)
    T.__matmul__ = lambda a, b: matmul(a, b)
    T.__gt__ = lambda a, b: greater(a, b)
    T.__lt__ = lambda a, b: less(a, b)

    def _getitem(a, key):
        return _node("StridedSlice", (a,), {"key": key}, "strided_slice", a.dtype, None)
    T.__getitem__ = _getitem


register_kernel("StridedSlice")(lambda ctx, n, x: x[n.attrs["key"]])
_install_operators()
