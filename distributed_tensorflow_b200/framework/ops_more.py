"""Second batch of TF-1.x ops next to the reference's own surface (SURVEY section 2.2 lists that one): unstack / slice / argmin,
``einsum`` / ``tensordot`` / ``accumulate_n``, ``random_shuffle``, the numeric guards (``is_nan`` / ``is_finite`` /
``check_numerics`` / ``verify_tensor_all_finite`` / ``Assert`` / ``assert_equal``), ``sparse_to_dense`` and
``softmax_cross_entropy_with_logits_v2``.  Same construction as ``framework/ops_extra.py``: one node per builder, one kernel per
node type over torch tensors (differentiable wherever torch's op is); nothing here is on the benchmarked path."""
from __future__ import annotations

from typing import Sequence

import torch

from . import errors
from . import shapes as _sh
from .graph import convert_to_tensor
from .ops import _node, add_n, bool_, cast, int32, int64, register_kernel

__all__ = ["unstack", "slice", "argmin", "arg_min", "einsum", "tensordot", "accumulate_n", "random_shuffle", "is_nan", "is_finite", "is_inf",
           "check_numerics", "verify_tensor_all_finite", "Assert", "assert_equal", "assert_less", "assert_greater", "sparse_to_dense",
           "softmax_cross_entropy_with_logits_v2", "to_int32", "to_int64", "to_double"]


def to_int32(x, name="ToInt32"): return cast(x, int32, name=name)
def to_int64(x, name="ToInt64"): return cast(x, int64, name=name)


def to_double(x, name="ToDouble"):
    from .ops import float64
    return cast(x, float64, name=name)


def unstack(value, num=None, axis=0, name="unstack"):
    """``tf.unstack``: the ``num`` slices of ``value`` along ``axis`` with that axis removed (``num`` from the static shape)."""
    x = convert_to_tensor(value)
    if num is None:
        if x.shape is None or x.shape[axis] is None:
            raise ValueError("unstack(): the size of axis %d is not known statically; pass num=" % axis)
        num = x.shape[axis]
    shp = None
    if x.shape is not None:
        shp = tuple(d for i, d in enumerate(x.shape) if i != axis % len(x.shape))
    return [_node("UnstackPart", (x,), {"axis": int(axis), "index": i}, "%s_%d" % (name, i), x.dtype, shp) for i in range(int(num))]


register_kernel("UnstackPart")(lambda ctx, n, x: x.select(n.attrs["axis"], n.attrs["index"]))


def slice(input_, begin, size, name="Slice"):          # noqa: A001 - TF's name
    """``tf.slice(x, begin, size)`` with python-list ``begin`` / ``size`` (``-1`` = to the end of that dimension)."""
    x = convert_to_tensor(input_)
    b, s = [int(v) for v in begin], [int(v) for v in size]
    shp = None
    if x.shape is not None:
        shp = tuple((None if d is None else d - bb) if ss < 0 else ss for d, bb, ss in zip(x.shape, b, s))
    return _node("Slice", (x,), {"begin": b, "size": s}, name, x.dtype, shp)


@register_kernel("Slice")
def _k_slice(ctx, n, x):
    for ax, (b, s) in enumerate(zip(n.attrs["begin"], n.attrs["size"])):
        x = x.narrow(ax, b, (x.shape[ax] - b) if s < 0 else s)
    return x


def argmin(x, axis=None, name="ArgMin", dimension=None, output_type=int64):
    x = convert_to_tensor(x)
    ax = dimension if axis is None else axis
    ax = 0 if ax is None else int(ax)
    return _node("ArgMin", (x,), {"axis": ax}, name, int64, _sh.reduce_shape(x.shape, ax, False))


arg_min = argmin
register_kernel("ArgMin")(lambda ctx, n, x: torch.argmin(x, dim=n.attrs["axis"]))


def einsum(equation, *inputs, name="einsum"):
    xs = [convert_to_tensor(v) for v in inputs]
    return _node("Einsum", xs, {"equation": str(equation)}, name, xs[0].dtype, None)


register_kernel("Einsum")(lambda ctx, n, *xs: torch.einsum(n.attrs["equation"], *xs))


def tensordot(a, b, axes, name="Tensordot"):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    if isinstance(axes, int):
        ax = int(axes)
    else:
        ax = [[int(v) for v in (p if isinstance(p, (list, tuple)) else [p])] for p in axes]
    return _node("Tensordot", (a, b), {"axes": ax}, name, a.dtype, None)


register_kernel("Tensordot")(lambda ctx, n, a, b: torch.tensordot(a, b, dims=n.attrs["axes"]))


def accumulate_n(inputs: Sequence, shape=None, tensor_dtype=None, name="AccumulateNV2"):
    return add_n(list(inputs), name=name)


def random_shuffle(value, seed=None, name="RandomShuffle"):
    """Rows of ``value`` in a random order, a new one every run (an op-level ``seed`` or the graph seed makes the sequence of
    orders reproducible, like the other random ops)."""
    x = convert_to_tensor(value)
    return _node("RandomShuffle", (x,), {"seed": seed}, name, x.dtype, x.shape)


@register_kernel("RandomShuffle", stateful=True)
def _k_random_shuffle(ctx, n, x):
    from .ops import _random_fill
    keys = _random_fill(ctx, n, 0, 0.0, 1.0, shape=(int(x.shape[0]),), device=torch.device("cpu"))      # the op's own Philox stream
    return x.index_select(0, torch.argsort(keys).to(x.device))


def is_nan(x, name="IsNan"):
    x = convert_to_tensor(x)
    return _node("IsNan", (x,), {}, name, bool_, x.shape)


def is_inf(x, name="IsInf"):
    x = convert_to_tensor(x)
    return _node("IsInf", (x,), {}, name, bool_, x.shape)


def is_finite(x, name="IsFinite"):
    x = convert_to_tensor(x)
    return _node("IsFinite", (x,), {}, name, bool_, x.shape)


register_kernel("IsNan")(lambda ctx, n, x: torch.isnan(x))
register_kernel("IsInf")(lambda ctx, n, x: torch.isinf(x))
register_kernel("IsFinite")(lambda ctx, n, x: torch.isfinite(x))


def check_numerics(tensor, message, name="CheckNumerics"):
    """Identity that fails the run with ``InvalidArgumentError`` when the tensor holds a NaN or an infinity."""
    x = convert_to_tensor(tensor)
    return _node("CheckNumerics", (x,), {"message": str(message)}, name, x.dtype, x.shape)


def verify_tensor_all_finite(t, msg, name="VerifyFinite"):
    return check_numerics(t, msg, name=name)


@register_kernel("CheckNumerics", stateful=True)
def _k_check_numerics(ctx, n, x):
    if x.is_floating_point() and not bool(torch.isfinite(x).all()):
        kind = "NaN" if bool(torch.isnan(x).any()) else "Inf"
        raise errors.InvalidArgumentError("%s : Tensor had %s values (%s)" % (n.attrs["message"], kind, n.name))
    return x


def Assert(condition, data, summarize=None, name="Assert"):      # noqa: N802 - TF's name
    """``tf.Assert``: an op (fetch it, or put it under ``control_dependencies``) that fails the run with
    ``InvalidArgumentError`` -- printing ``data`` -- when ``condition`` does not hold."""
    c = convert_to_tensor(condition)
    data = list(data) if isinstance(data, (list, tuple)) else [data]
    ds = [convert_to_tensor(d) for d in data if not isinstance(d, str)]            # strings are message text
    return _node("Assert", [c] + ds, {"summarize": summarize, "message": " ".join(d for d in data if isinstance(d, str))}, name, None, None)


@register_kernel("Assert", stateful=True)
def _k_assert(ctx, n, c, *data):
    if not bool(torch.as_tensor(c).all()):
        k = n.attrs.get("summarize") or 3
        shown = ["%s" % (d.reshape(-1)[:k].tolist() if isinstance(d, torch.Tensor) else d,) for d in data]
        msg = n.attrs.get("message") or ""
        raise errors.InvalidArgumentError("assertion failed: %s%s (%s)" % (msg + " " if msg else "", " ".join(shown), n.name))
    return None


def _assert_cmp(op_name, x, y, data, summarize, message, name):
    x, y = convert_to_tensor(x), convert_to_tensor(y)
    cond = _node(op_name, (x, y), {}, name + "/cond", bool_, ())
    shown = [convert_to_tensor(d) for d in data if not isinstance(d, str)] if data is not None else [x, y]
    texts = [d for d in (data or []) if isinstance(d, str)] + ([str(message)] if message else [])
    return _node("Assert", [cond] + shown, {"summarize": summarize, "message": " ".join(texts)}, name, None, None)


for _op, _fn in (("AllEqual", torch.eq), ("AllLess", torch.lt), ("AllGreater", torch.gt)):
    register_kernel(_op)(lambda ctx, n, a, b, _f=_fn: _f(a, b).all())


def assert_equal(x, y, data=None, summarize=None, message=None, name="assert_equal"):
    return _assert_cmp("AllEqual", x, y, data, summarize, message, name)


def assert_less(x, y, data=None, summarize=None, message=None, name="assert_less"):
    return _assert_cmp("AllLess", x, y, data, summarize, message, name)


def assert_greater(x, y, data=None, summarize=None, message=None, name="assert_greater"):
    return _assert_cmp("AllGreater", x, y, data, summarize, message, name)


def sparse_to_dense(sparse_indices, output_shape, sparse_values, default_value=0, validate_indices=True, name="SparseToDense"):
    """Dense tensor of ``output_shape`` holding ``sparse_values`` at ``sparse_indices`` (``[n]`` for a vector, ``[n, rank]``
    otherwise) and ``default_value`` elsewhere -- the classic TF-1.x way of building one-hot labels."""
    idx, vals = convert_to_tensor(sparse_indices), convert_to_tensor(sparse_values)
    shape = [int(v) for v in output_shape]
    return _node("SparseToDense", (idx, vals), {"shape": shape, "default": default_value}, name, vals.dtype, tuple(shape))


@register_kernel("SparseToDense")
def _k_sparse_to_dense(ctx, n, idx, vals):
    shape = n.attrs["shape"]
    vals = torch.as_tensor(vals)
    out = torch.full(shape, n.attrs["default"], dtype=vals.dtype, device=vals.device)
    idx = idx.long()
    if idx.dim() == 0:
        idx = idx.reshape(1, 1)
    elif idx.dim() == 1:
        idx = idx.reshape(-1, 1) if len(shape) == 1 else idx.reshape(1, -1)
    v = vals.expand(idx.shape[0]) if vals.dim() == 0 else vals
    out[tuple(idx[:, d] for d in range(idx.shape[1]))] = v
    return out


def softmax_cross_entropy_with_logits_v2(labels=None, logits=None, axis=-1, name="softmax_cross_entropy_with_logits_v2", dim=None):
    """Per-row ``-sum(labels * log_softmax(logits))``; unlike the v1 op, gradients also flow into ``labels``."""
    from .ops import log_softmax, multiply, negative, reduce_sum
    ax = axis if dim is None else dim
    return negative(reduce_sum(multiply(convert_to_tensor(labels), log_softmax(logits, axis=ax)), axis=ax), name=name)
