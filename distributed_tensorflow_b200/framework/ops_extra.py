"""Ops a TF-1.x ps/worker program commonly uses NEXT to the ones the reference scripts call (SURVEY section 2.2 lists those):
shape manipulation, comparisons / selection, rounding, norms and gradient clipping, a few activations and losses, ``tf.Print`` /
``tf.py_func``.  Same construction as ``framework/ops.py`` (one node per builder, one kernel per node type over torch tensors, so
``tf.gradients`` differentiates through them); nothing here is on the benchmarked path.  ``tf.cond`` builds both branches and
executes the chosen one; graph loops (``tf.while_loop``) are not provided: programs of the reference's kind loop in Python around
``Session.run``."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F

from . import shapes as _sh
from .graph import Tensor, convert_to_tensor
from .ops import (_node, as_dtype, bool_, cast, constant, float32, int32, int64, multiply, reduce_sum, register_kernel, sqrt,
                  square)

__all__ = ["tile", "gather", "where", "floor", "ceil", "round", "sign", "reduce_prod", "reduce_all", "reduce_any", "logical_and",
           "logical_or", "logical_not", "greater_equal", "less_equal", "not_equal", "size", "rank", "range", "linspace", "norm",
           "global_norm", "clip_by_norm", "clip_by_global_norm", "Print", "py_func", "relu6", "elu", "leaky_relu", "softplus",
           "sigmoid_cross_entropy_with_logits", "l2_normalize", "embedding_lookup", "in_top_k", "top_k", "cumsum", "reverse",
           "pad", "matrix_transpose", "diag_part", "trace", "erf", "log1p", "expm1", "floordiv", "mod", "eye"]


def _un(op, x, name, dtype=None, shape="same", **attrs):
    x = convert_to_tensor(x)
    return _node(op, (x,), attrs, name, dtype or x.dtype, x.shape if shape == "same" else shape)


def _bin(op, a, b, name, dtype=None):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    return _node(op, (a, b), {}, name, dtype or a.dtype or b.dtype, _sh.broadcast_shape(a.shape, b.shape))


# -- element-wise -------------------------------------------------------------------------------------------------------------
def floor(x, name="Floor"): return _un("Floor", x, name)
def ceil(x, name="Ceil"): return _un("Ceil", x, name)
def round(x, name="Round"): return _un("Round", x, name)          # noqa: A001 - TF's name (banker's rounding, like torch.round)
def sign(x, name="Sign"): return _un("Sign", x, name)
def erf(x, name="Erf"): return _un("Erf", x, name)
def log1p(x, name="Log1p"): return _un("Log1p", x, name)
def expm1(x, name="Expm1"): return _un("Expm1", x, name)
def softplus(x, name="Softplus"): return _un("Softplus", x, name)
def relu6(x, name="Relu6"): return _un("Relu6", x, name)
def elu(x, name="Elu"): return _un("Elu", x, name)
def leaky_relu(x, alpha=0.2, name="LeakyRelu"): return _un("LeakyRelu", x, name, alpha=float(alpha))
def logical_not(x, name="LogicalNot"): return _un("LogicalNot", x, name, dtype=bool_)
def logical_and(a, b, name="LogicalAnd"): return _bin("LogicalAnd", a, b, name, bool_)
def logical_or(a, b, name="LogicalOr"): return _bin("LogicalOr", a, b, name, bool_)
def greater_equal(a, b, name="GreaterEqual"): return _bin("GreaterEqual", a, b, name, bool_)
def less_equal(a, b, name="LessEqual"): return _bin("LessEqual", a, b, name, bool_)
def not_equal(a, b, name="NotEqual"): return _bin("NotEqual", a, b, name, bool_)
def floordiv(a, b, name="FloorDiv"): return _bin("FloorDiv", a, b, name)
def mod(a, b, name="FloorMod"): return _bin("FloorMod", a, b, name)


for _op, _fn in (("Floor", torch.floor), ("Ceil", torch.ceil), ("Round", torch.round), ("Sign", torch.sign), ("Erf", torch.erf),
                 ("Log1p", torch.log1p), ("Expm1", torch.expm1), ("Softplus", F.softplus), ("Relu6", F.relu6), ("Elu", F.elu),
                 ("LogicalNot", torch.logical_not)):
    register_kernel(_op)(lambda ctx, n, x, _f=_fn: _f(x))
register_kernel("LeakyRelu")(lambda ctx, n, x: F.leaky_relu(x, n.attrs["alpha"]))
for _op, _fn in (("LogicalAnd", torch.logical_and), ("LogicalOr", torch.logical_or), ("GreaterEqual", torch.ge), ("LessEqual", torch.le),
                 ("NotEqual", torch.ne), ("FloorDiv", lambda a, b: torch.div(a, b, rounding_mode="floor")), ("FloorMod", torch.remainder)):
    register_kernel(_op)(lambda ctx, n, a, b, _f=_fn: _f(a, b))


def where(condition, x=None, y=None, name="Where"):
    """``tf.where(c, x, y)``: element-wise select; ``tf.where(c)``: the ``[n, rank]`` int64 coordinates of the true entries."""
    c = convert_to_tensor(condition)
    if x is None and y is None:
        return _node("WhereIndices", (c,), {}, name, int64, None)
    x, y = convert_to_tensor(x), convert_to_tensor(y)
    return _node("Select", (c, x, y), {}, name, x.dtype or y.dtype, _sh.broadcast_shape(x.shape, y.shape))


register_kernel("WhereIndices")(lambda ctx, n, c: torch.nonzero(c))


@register_kernel("Select")
def _k_select(ctx, n, c, x, y):
    if c.dim() == 1 and x.dim() > 1 and c.shape[0] == x.shape[0]:      # TF: a vector condition selects whole rows
        c = c.reshape((-1,) + (1,) * (x.dim() - 1))
    return torch.where(c.bool(), x, y)


# -- shapes, indexing ------------------------------------------------------------------------------------------------------------
def tile(x, multiples, name="Tile"):
    x = convert_to_tensor(x)
    m = [int(v) for v in multiples]
    shp = None if x.shape is None else tuple(None if d is None else d * k for d, k in zip(x.shape, m))
    return _node("Tile", (x,), {"multiples": m}, name, x.dtype, shp)


register_kernel("Tile")(lambda ctx, n, x: x.repeat(*n.attrs["multiples"]))


def gather(params, indices, axis=0, name="GatherV2"):
    p, i = convert_to_tensor(params), convert_to_tensor(indices)
    return _node("GatherV2", (p, i), {"axis": int(axis)}, name, p.dtype, None)


@register_kernel("GatherV2")
def _k_gather(ctx, n, p, i):
    ax = n.attrs["axis"] % p.dim()
    out = torch.index_select(p, ax, i.reshape(-1).long())
    return out.reshape(tuple(p.shape[:ax]) + tuple(i.shape) + tuple(p.shape[ax + 1:]))


def embedding_lookup(params, ids, name="embedding_lookup"):
    return gather(params, ids, axis=0, name=name)


def size(x, name="Size", out_type=int32): return _un("Size", x, name, dtype=as_dtype(out_type), shape=())
def rank(x, name="Rank"): return _un("Rank", x, name, dtype=int32, shape=())


register_kernel("Size")(lambda ctx, n, x: torch.tensor(x.numel(), dtype=n.dtype or torch.int32))
register_kernel("Rank")(lambda ctx, n, x: torch.tensor(x.dim(), dtype=torch.int32))


def range(start, limit=None, delta=1, dtype=None, name="Range"):      # noqa: A001
    if limit is None:
        start, limit = 0, start
    is_float = any(isinstance(v, float) for v in (start, limit, delta))
    dt = as_dtype(dtype) or (float32 if is_float else int32)
    n = max(0, int(-(-(limit - start) // delta))) if not is_float else None
    return _node("Range", (), {"start": start, "limit": limit, "delta": delta}, name, dt, None if n is None else (n,))


register_kernel("Range")(lambda ctx, n: torch.arange(n.attrs["start"], n.attrs["limit"], n.attrs["delta"], dtype=n.dtype).to(ctx.torch_device(n)))


def linspace(start, stop, num, name="LinSpace"):
    return _node("LinSpace", (), {"start": float(start), "stop": float(stop), "num": int(num)}, name, float32, (int(num),))


register_kernel("LinSpace")(lambda ctx, n: torch.linspace(n.attrs["start"], n.attrs["stop"], n.attrs["num"]).to(ctx.torch_device(n)))


def eye(num_rows, num_columns=None, dtype=float32, name="eye"):
    m = int(num_columns if num_columns is not None else num_rows)
    return constant(torch.eye(int(num_rows), m, dtype=as_dtype(dtype)), name=name)


def cumsum(x, axis=0, exclusive=False, reverse=False, name="Cumsum"):
    return _un("Cumsum", x, name, axis=int(axis), exclusive=bool(exclusive), reverse=bool(reverse))


@register_kernel("Cumsum")
def _k_cumsum(ctx, n, x):
    ax = n.attrs["axis"]
    if n.attrs["reverse"]:
        x = torch.flip(x, [ax])
    out = torch.cumsum(x, ax)
    if n.attrs["exclusive"]:
        out = out - x
    return torch.flip(out, [ax]) if n.attrs["reverse"] else out


def reverse(x, axis, name="ReverseV2"):
    ax = [int(a) for a in (axis if isinstance(axis, (list, tuple)) else [axis])]
    return _un("ReverseV2", x, name, axis=ax)


register_kernel("ReverseV2")(lambda ctx, n, x: torch.flip(x, n.attrs["axis"]))


def pad(x, paddings, mode="CONSTANT", constant_values=0, name="Pad"):
    x = convert_to_tensor(x)
    p = [[int(a), int(b)] for a, b in paddings]
    shp = None if x.shape is None else tuple(None if d is None else d + a + b for d, (a, b) in zip(x.shape, p))
    return _node("Pad", (x,), {"paddings": p, "mode": mode.upper(), "value": constant_values}, name, x.dtype, shp)


@register_kernel("Pad")
def _k_pad(ctx, n, x):
    flat = []
    for a, b in reversed(n.attrs["paddings"]):                          # torch pads from the last dimension backwards
        flat += [a, b]
    mode = {"CONSTANT": "constant", "REFLECT": "reflect", "SYMMETRIC": "replicate"}[n.attrs["mode"]]
    return F.pad(x, flat, mode=mode, value=n.attrs["value"]) if mode == "constant" else F.pad(x, flat, mode=mode)


def matrix_transpose(x, name="matrix_transpose"): return _un("MatrixTranspose", x, name, shape=None)
def diag_part(x, name="DiagPart"): return _un("DiagPart", x, name, shape=None)
def trace(x, name="Trace"): return _un("Trace", x, name, shape=None)


register_kernel("MatrixTranspose")(lambda ctx, n, x: x.transpose(-1, -2))
register_kernel("DiagPart")(lambda ctx, n, x: torch.diagonal(x, dim1=-2, dim2=-1))
register_kernel("Trace")(lambda ctx, n, x: torch.diagonal(x, dim1=-2, dim2=-1).sum(-1))


# -- reductions, norms, clipping ---------------------------------------------------------------------------------------------------
def _reduce(op, x, axis, keepdims, name):
    x = convert_to_tensor(x)
    ax = None if axis is None else ([int(a) for a in axis] if isinstance(axis, (list, tuple)) else [int(axis)])
    return _node(op, (x,), {"axis": ax, "keepdims": bool(keepdims)}, name, x.dtype if op == "Prod" else bool_,
                 _sh.reduce_shape(x.shape, ax if ax is None or len(ax) > 1 else ax[0], bool(keepdims)))


def reduce_prod(x, axis=None, keepdims=False, name="Prod"): return _reduce("Prod", x, axis, keepdims, name)
def reduce_all(x, axis=None, keepdims=False, name="All"): return _reduce("All", x, axis, keepdims, name)
def reduce_any(x, axis=None, keepdims=False, name="Any"): return _reduce("Any", x, axis, keepdims, name)


def _red_kernel(fn_all, fn_dim):
    def k(ctx, n, x):
        ax, kd = n.attrs["axis"], n.attrs["keepdims"]
        if ax is None:
            out = fn_all(x)
            return out.reshape([1] * x.dim()) if kd else out
        for a in sorted((a % x.dim() for a in ax), reverse=True):
            x = fn_dim(x, a, kd)
        return x
    return k


register_kernel("Prod")(_red_kernel(lambda x: x.prod(), lambda x, a, k: x.prod(dim=a, keepdim=k)))
register_kernel("All")(_red_kernel(lambda x: x.bool().all(), lambda x, a, k: x.bool().all(dim=a, keepdim=k)))
register_kernel("Any")(_red_kernel(lambda x: x.bool().any(), lambda x, a, k: x.bool().any(dim=a, keepdim=k)))


def norm(x, ord="euclidean", axis=None, keepdims=False, name="norm"):        # noqa: A002
    x = convert_to_tensor(x)
    if ord in ("euclidean", 2, 2.0):
        return sqrt(reduce_sum(square(x), axis=axis, keepdims=keepdims), name=name)
    if ord in (1, 1.0):
        from .ops import abs as _abs
        return reduce_sum(_abs(x), axis=axis, keepdims=keepdims, name=name)
    raise ValueError("norm: ord must be 1 or 2 ('euclidean'), got %r" % (ord,))


def global_norm(t_list: Sequence[Tensor], name="global_norm"):
    """``sqrt(sum_i ||t_i||^2)`` over the non-None tensors (TF's ``tf.global_norm``)."""
    ts = [convert_to_tensor(t) for t in t_list if t is not None]
    return _node("GlobalNorm", ts, {}, name, float32, ())


register_kernel("GlobalNorm")(lambda ctx, n, *ts: torch.sqrt(sum((t.float() * t.float()).sum() for t in ts)))


def clip_by_norm(t, clip_norm, axes=None, name="clip_by_norm"):
    t = convert_to_tensor(t)
    ax = None if axes is None else [int(a) for a in (axes if isinstance(axes, (list, tuple)) else [axes])]
    return _node("ClipByNorm", (t,), {"clip_norm": float(clip_norm), "axes": ax}, name, t.dtype, t.shape)


@register_kernel("ClipByNorm")
def _k_clip_by_norm(ctx, n, t):
    ax = n.attrs["axes"]
    l2 = torch.sqrt((t * t).sum() if ax is None else (t * t).sum(dim=ax, keepdim=True))
    c = n.attrs["clip_norm"]
    return t * c / torch.maximum(l2, torch.as_tensor(c, dtype=t.dtype, device=t.device))


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name="clip_by_global_norm"):
    """Returns ``(clipped list, global norm)``: every tensor scaled by ``clip_norm / max(global_norm, clip_norm)`` -- the usual
    guard in front of ``apply_gradients``."""
    gn = use_norm if use_norm is not None else global_norm(t_list)
    out = []
    for i, t in enumerate(t_list):
        if t is None:
            out.append(None)
            continue
        t = convert_to_tensor(t)
        out.append(_node("ScaleByGlobalNorm", (t, gn), {"clip_norm": float(clip_norm)}, "%s_%d" % (name, i), t.dtype, t.shape))
    return out, gn


@register_kernel("ScaleByGlobalNorm")
def _k_scale_gn(ctx, n, t, gn):
    c = n.attrs["clip_norm"]
    gn = gn.to(t.device)
    return t * (c / torch.maximum(gn, torch.as_tensor(c, dtype=gn.dtype, device=gn.device))).to(t.dtype)


# -- nn ---------------------------------------------------------------------------------------------------------------------------------
def sigmoid_cross_entropy_with_logits(labels=None, logits=None, name="logistic_loss"):
    z, y = convert_to_tensor(logits), convert_to_tensor(labels)
    return _node("SigmoidXent", (z, y), {}, name, z.dtype, z.shape)


register_kernel("SigmoidXent")(lambda ctx, n, z, y: F.binary_cross_entropy_with_logits(z, y.to(z.dtype), reduction="none"))


def l2_normalize(x, axis=None, epsilon=1e-12, name="l2_normalize", dim=None):
    x = convert_to_tensor(x)
    ax = dim if axis is None else axis
    ax = None if ax is None else [int(a) for a in (ax if isinstance(ax, (list, tuple)) else [ax])]
    return _node("L2Normalize", (x,), {"axis": ax, "eps": float(epsilon)}, name, x.dtype, x.shape)


@register_kernel("L2Normalize")
def _k_l2n(ctx, n, x):
    ax = n.attrs["axis"]
    ss = (x * x).sum() if ax is None else (x * x).sum(dim=ax, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=n.attrs["eps"]))


def in_top_k(predictions, targets, k, name="InTopK"):
    p, t = convert_to_tensor(predictions), convert_to_tensor(targets)
    return _node("InTopK", (p, t), {"k": int(k)}, name, bool_, None if p.shape is None else (p.shape[0],))


@register_kernel("InTopK")
def _k_in_top_k(ctx, n, p, t):
    tgt = p.gather(1, t.long().reshape(-1, 1))
    return (p > tgt).sum(dim=1) < n.attrs["k"]                      # fewer than k classes score strictly higher


def top_k(x, k=1, sorted=True, name="TopKV2"):                       # noqa: A002
    x = convert_to_tensor(x)
    both = _node("TopKV2", (x,), {"k": int(k)}, name, x.dtype, None)
    return (_node("TupleItem", (both,), {"index": 0}, name + "_values", x.dtype, None),
            _node("TupleItem", (both,), {"index": 1}, name + "_indices", int32, None))


register_kernel("TopKV2")(lambda ctx, n, x: tuple(torch.topk(x, n.attrs["k"], dim=-1)))
register_kernel("TupleItem")(lambda ctx, n, t: t[n.attrs["index"]] if n.attrs["index"] == 0 else t[1].to(torch.int32))


# -- debugging / host callbacks ------------------------------------------------------------------------------------------------------------
def Print(input_, data, message=None, first_n=None, summarize=None, name="Print"):      # noqa: N802 - TF's spelling
    """Identity on ``input_`` that prints ``message`` + the values of ``data`` to stderr when it runs (at most ``first_n`` times)."""
    x = convert_to_tensor(input_)
    ds = [convert_to_tensor(d) for d in data]
    return _node("Print", [x] + ds, {"message": message or "", "first_n": first_n, "summarize": summarize or 3, "count": [0]}, name,
                 x.dtype, x.shape)


@register_kernel("Print", stateful=True)
def _k_print(ctx, n, x, *data):
    import sys
    a = n.attrs
    a["count"][0] += 1
    if a["first_n"] is None or a["first_n"] < 0 or a["count"][0] <= a["first_n"]:
        parts = ["[%s%s]" % (" ".join("%g" % v for v in d.detach().reshape(-1)[:a["summarize"]].tolist()),
                            "..." if d.numel() > a["summarize"] else "") for d in data]
        print(a["message"] + "".join(parts), file=sys.stderr, flush=True)
    return x


def py_func(func, inp, Tout, stateful=True, name="PyFunc"):           # noqa: N803 - TF's argument name
    """Run ``func(*numpy arrays)`` on the host inside the graph; ``Tout``: a dtype or a list of dtypes.  Not differentiable."""
    ins = [convert_to_tensor(i) for i in inp]
    many = isinstance(Tout, (list, tuple))
    outs = [as_dtype(t) for t in (Tout if many else [Tout])]
    node = _node("PyFunc", ins, {"func": func, "dtypes": outs}, name, outs[0] if not many else None, None)
    if not many:
        return _node("TupleItemRaw", (node,), {"index": 0}, name + "_0", outs[0], None)
    return [_node("TupleItemRaw", (node,), {"index": i}, "%s_%d" % (name, i), dt, None) for i, dt in enumerate(outs)]


@register_kernel("PyFunc", stateful=True)
def _k_py_func(ctx, n, *xs):
    import numpy as np
    res = n.attrs["func"](*[x.detach().cpu().numpy() for x in xs])
    res = res if isinstance(res, (list, tuple)) else [res]
    dev = ctx.torch_device(n)
    return tuple(torch.as_tensor(np.asarray(r)).to(dt).to(dev) for r, dt in zip(res, n.attrs["dtypes"]))


register_kernel("TupleItemRaw")(lambda ctx, n, t: t[n.attrs["index"]])


# -- tf.cond --------------------------------------------------------------------------------------------------------------------------
def cond(pred, true_fn=None, false_fn=None, name="cond", fn1=None, fn2=None):
    """``tf.cond``: both branches are BUILT (their nodes are created, like TF), only the chosen one is EXECUTED: the branch
    sub-graphs hang off the ``Cond`` node and run inside its kernel over the values they capture from outside.  Differentiable
    (the reverse pass follows the branch that ran).  The node must run on the task that built the graph (the usual placement of
    worker-side compute); graph loops (``tf.while_loop``) are not provided."""
    from .graph import get_default_graph
    true_fn, false_fn = true_fn or fn1, false_fn or fn2
    if true_fn is None or false_fn is None:
        raise TypeError("cond(): true_fn and false_fn are required")
    g = get_default_graph()
    p = convert_to_tensor(pred)
    start = len(g.nodes)
    t_res = true_fn()
    mid = len(g.nodes)
    f_res = false_fn()
    end = len(g.nodes)

    def flat(r):
        return [convert_to_tensor(v) for v in (r if isinstance(r, (list, tuple)) else [r])]
    single = not isinstance(t_res, (list, tuple))
    t_out, f_out = flat(t_res), flat(f_res)
    end = len(g.nodes)                                     # convert_to_tensor may have added constants: they belong to the false side
    if len(t_out) != len(f_out):
        raise ValueError("cond(): true_fn and false_fn must return the same number of tensors (%d vs %d)" % (len(t_out), len(f_out)))
    t_nodes, f_nodes = list(g.nodes[start:mid]), list(g.nodes[mid:end])
    captured, seen = [], set()
    for n in t_nodes + f_nodes:
        for d in list(n.inputs) + list(n.control_inputs):
            if d.id < start and d.id not in seen:
                seen.add(d.id)
                captured.append(d)
    for o in t_out + f_out:                                   # a branch that returns an outer tensor as it is
        if o.id < start and o.id not in seen:
            seen.add(o.id)
            captured.append(o)
    node = _node("Cond", [p] + captured, {"t_nodes": t_nodes, "f_nodes": f_nodes, "t_out": [o.id for o in t_out],
                                          "f_out": [o.id for o in f_out], "captured": [c.id for c in captured]}, name, None, None)
    outs = [_node("TupleItemRaw", (node,), {"index": i}, "%s_%d" % (name, i), to.dtype or fo.dtype,
                  to.shape if to.shape == fo.shape else None) for i, (to, fo) in enumerate(zip(t_out, f_out))]
    return outs[0] if single else outs


@register_kernel("Cond")
def _k_cond(ctx, node, pred, *captured):
    from .executor import ExecContext, execute
    a = node.attrs
    take_true = bool(pred.reshape(-1)[0]) if isinstance(pred, torch.Tensor) else bool(pred)
    nodes, out_ids = (a["t_nodes"], a["t_out"]) if take_true else (a["f_nodes"], a["f_out"])
    sub = ExecContext(ctx.store, ctx.task, ctx.gpu_index, ctx.tracer, getattr(ctx, "_seed", None), ctx.allow_soft_placement)
    sub.force_device, sub.leaves = ctx.force_device, ctx.leaves
    for attr in ("cancel_event", "server"):
        if hasattr(ctx, attr):
            setattr(sub, attr, getattr(ctx, attr))
    for cid, v in zip(a["captured"], captured):
        sub.values[cid] = v
    execute(nodes, sub, torch.is_grad_enabled())
    return tuple(sub.values[i] for i in out_ids)


__all__.append("cond")


# -- tf.while_loop ----------------------------------------------------------------------------------------------------------------------
def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10, back_prop=True, swap_memory=False, name="while",
               maximum_iterations=None):
    """``tf.while_loop``: ``cond`` and ``body`` are built ONCE over stand-in nodes for the loop variables; the ``While`` node's kernel
    evaluates the condition sub-graph and, while it holds, the body sub-graph, feeding each iteration's results back in.  The
    iterations are recorded on the autograd tape as they run, so ``tf.gradients`` differentiates the unrolled loop.  Same
    placement rule as :func:`cond` (the node runs on the task that built the graph)."""
    from .graph import get_default_graph
    g = get_default_graph()
    single = not isinstance(loop_vars, (list, tuple))
    init = [convert_to_tensor(v) for v in ([loop_vars] if single else loop_vars)]
    start = len(g.nodes)
    stand_ins = [_node("LoopVar", (), {}, "%s/var_%d" % (name, i), v.dtype, v.shape) for i, v in enumerate(init)]
    c0 = len(g.nodes)
    c_out = convert_to_tensor(cond(*stand_ins))
    c1 = len(g.nodes)
    b_res = body(*stand_ins)
    b_out = [convert_to_tensor(v) for v in (b_res if isinstance(b_res, (list, tuple)) else [b_res])]
    end = len(g.nodes)
    if len(b_out) != len(init):
        raise ValueError("while_loop(): body must return as many tensors as there are loop variables (%d vs %d)" % (len(b_out), len(init)))
    c_nodes, b_nodes = list(g.nodes[c0:c1]), list(g.nodes[c1:end])
    captured, seen = [], set()
    for n in c_nodes + b_nodes:
        for d in list(n.inputs) + list(n.control_inputs):
            if d.id < start and d.id not in seen:
                seen.add(d.id)
                captured.append(d)
    for o in [c_out] + b_out:
        if o.id < start and o.id not in seen:
            seen.add(o.id)
            captured.append(o)
    node = _node("While", init + captured, {"n": len(init), "vars": [s.id for s in stand_ins], "c_nodes": c_nodes, "b_nodes": b_nodes,
                                           "c_out": c_out.id, "b_out": [o.id for o in b_out], "captured": [c.id for c in captured],
                                           "max_iter": maximum_iterations}, name, None, None)
    outs = [_node("TupleItemRaw", (node,), {"index": i}, "%s_%d" % (name, i), v.dtype, v.shape if b.shape == v.shape else None)
            for i, (v, b) in enumerate(zip(init, b_out))]
    return outs[0] if single else outs


@register_kernel("LoopVar")
def _k_loop_var(ctx, node):
    raise RuntimeError("a while_loop variable (%s) was fetched outside its loop" % node.name)


@register_kernel("While")
def _k_while(ctx, node, *args):
    from .executor import ExecContext, execute
    a = node.attrs
    vals, captured = list(args[:a["n"]]), args[a["n"]:]

    def run(nodes, out_ids):
        sub = ExecContext(ctx.store, ctx.task, ctx.gpu_index, ctx.tracer, getattr(ctx, "_seed", None), ctx.allow_soft_placement)
        sub.force_device, sub.leaves = ctx.force_device, ctx.leaves
        for attr in ("cancel_event", "server"):
            if hasattr(ctx, attr):
                setattr(sub, attr, getattr(ctx, attr))
        for cid, v in zip(a["captured"], captured):
            sub.values[cid] = v
        for vid, v in zip(a["vars"], vals):
            sub.values[vid] = v
        execute(nodes, sub, torch.is_grad_enabled())
        return [sub.values[i] for i in out_ids]
    it = 0
    while a["max_iter"] is None or it < a["max_iter"]:
        c = run(a["c_nodes"], [a["c_out"]])[0]
        if not bool(c.reshape(-1)[0] if isinstance(c, torch.Tensor) else c):
            break
        vals = run(a["b_nodes"], a["b_out"])
        it += 1
        cancel = getattr(ctx, "cancel_event", None)
        if cancel is not None and cancel.is_set():
            from . import errors
            raise errors.CancelledError("while_loop cancelled")
    return tuple(vals)


__all__.append("while_loop")
