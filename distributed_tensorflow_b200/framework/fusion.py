"""Plan-time operator fusion for the graph tier (the role of TF's grappler remapper, here two rewrites that put element-wise
work into the producing kernel -- the B200 rule "fuse elementwise/activation work into the producer"):

* ``Relu(XwPlusB(x, W, b))`` (``/root/reference/distributed_mnist.py:109-110``) -> ONE GEMM with the bias + ReLU epilogue
  (``ops/native.linear(relu=True)``: out of TMEM on ``/gpu``), the ReLU mask applied in the backward GEMMs' producer;
* ``Relu(MatMul(x, W) + b)`` / ``MatMul(x, W) + b`` with a vector or scalar ``b`` (``/root/reference/standalone.py:52-53,60``: the
  tower spelling) -> the same GEMM epilogue;
* ``reduce_mean(square(a - b))`` (``example_between_graph.py:59``, ``standalone.py:61``) -> one difference kernel + a
  sum-of-squares reduction (``ops/native.mse``), three kernels instead of seven per training step;
* ``-reduce_sum(y_ * log(clip_by_value(softmax(logits), eps, 1)))`` (``distributed_mnist.py:112-113``) -> ONE fused
  softmax + clipped cross-entropy forward/backward kernel (``ops/native.clipped_softmax_xent_sum``) instead of six node kernels
  forward and six backward.

A rewrite is planned once per ``Session`` plan (fetches x feeds) and only when it cannot be observed: no interior value is
fetched, fed, differentiated against or consumed by another node of the plan, and every node of the pattern runs on the same
task (a task keeps one execution context per run, across its segments).  The executor applies it while walking the segment (``try_execute``); whether the fused kernel
accepts the operands (2-D fp32, matching shapes) is decided when the first node of the pattern is reached -- otherwise the nodes
run one by one as before.  ``DTF_GRAPH_FUSION=0`` switches the pass off."""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Set

import torch

ENABLED = os.environ.get("DTF_GRAPH_FUSION", "1") == "1"


def _through_identity(t):
    while t is not None and t.op_type == "Identity" and t.inputs:
        t = t.inputs[0]
    return t


def plan_fusions(order: Sequence[Any], fetch_ids: Set[int], leaves: Set[int], task_of: Dict[int, Any], fed: Set[int]) -> Optional[Dict[str, list]]:
    """Wire form ``{"relu": [[xwb_id, relu_id], ...], "affine": [[matmul, add, relu or -1, x, w, b], ...],
    "mse": [[mean, square, sub, a, b], ...], "xent": [[neg, softmax, logits, labels, clip_min, [interior ids]], ...]}``
    (plain ints / floats: it travels to remote tasks inside the run options), or ``None``."""
    if not ENABLED:
        return None
    pos = {n.id: i for i, n in enumerate(order)}
    users: Dict[int, List[Any]] = {}
    for n in order:
        if n.id in fed:
            continue
        for d in list(n.inputs) + list(n.control_inputs):
            users.setdefault(d.id, []).append(n)
    protected = set(fetch_ids) | set(leaves) | set(fed)

    def private(node, consumer) -> bool:
        """``node`` is produced in this plan, seen by ``consumer`` only, and nobody outside can observe it."""
        return (node.id in pos and node.id not in protected and [u.id for u in users.get(node.id, ())] == [consumer.id]
                and node.id in task_of and consumer.id in task_of and task_of[node.id] == task_of[consumer.id])
    relu_pairs, affine, xent = [], [], []
    taken: Set[int] = set()
    for n in order:                      # MatMul + Add (+ Relu): planned first, a Relu on top joins below
        if n.id in fed or n.id not in task_of or n.op_type != "Add" or n.id in leaves or n.id in taken:
            continue
        a, b = n.inputs
        mm, bias = (a, b) if a.op_type == "MatMul" else ((b, a) if b.op_type == "MatMul" else (None, None))
        if mm is None or mm.attrs.get("ta") or mm.attrs.get("tb") or not private(mm, n) or mm.id in taken:
            continue
        if bias.id in pos and bias.id not in fed and pos[bias.id] > pos[mm.id]:
            continue                     # the bias must exist when the MatMul node (the decision point) is reached
        relu = -1
        us = users.get(n.id, ())
        if len(us) == 1 and us[0].op_type == "Relu" and private(n, us[0]) and us[0].id not in leaves and us[0].id not in taken:
            relu = us[0].id
        affine.append([mm.id, n.id, relu, mm.inputs[0].id, mm.inputs[1].id, bias.id])
        taken.update((mm.id, n.id) + ((relu,) if relu >= 0 else ()))
    mse = []
    for n in order:                      # Mean(Square(Sub(a, b))) over every element
        if n.id in fed or n.id not in task_of or n.op_type != "Mean" or n.id in leaves or n.id in taken:
            continue
        if n.attrs.get("axis") is not None or n.attrs.get("keepdims"):
            continue
        sq = n.inputs[0]
        if sq.op_type != "Square" or not private(sq, n) or sq.id in taken:
            continue
        sub = sq.inputs[0]
        if sub.op_type != "Sub" or not private(sub, sq) or sub.id in taken:
            continue
        mse.append([n.id, sq.id, sub.id, sub.inputs[0].id, sub.inputs[1].id])
        taken.update((n.id, sq.id, sub.id))
    for n in order:
        if n.id in fed or n.id not in task_of or n.id in taken:
            continue
        if n.op_type == "Relu" and n.id not in leaves:
            src = n.inputs[0]
            if src.op_type == "XwPlusB" and not src.attrs.get("relu") and private(src, n) and src.id not in taken:
                relu_pairs.append([src.id, n.id])
                taken.update((src.id, n.id))
        elif n.op_type == "Neg" and n.id not in leaves:
            red = n.inputs[0]
            if red.op_type != "Sum" or red.attrs.get("axis") is not None or red.attrs.get("keepdims") or not private(red, n):
                continue
            mul = red.inputs[0]
            if mul.op_type != "Mul" or not private(mul, red):
                continue
            a, b = mul.inputs
            log, labels = (a, b) if a.op_type == "Log" else ((b, a) if b.op_type == "Log" else (None, None))
            if log is None or not private(log, mul):
                continue
            clip = log.inputs[0]
            if clip.op_type != "ClipByValue" or not private(clip, log):
                continue
            lo, hi = clip.attrs.get("lo"), clip.attrs.get("hi")
            if not isinstance(lo, (int, float)) or not isinstance(hi, (int, float)) or float(hi) != 1.0 or not (0.0 <= float(lo) < 1e-3):
                continue
            sm = clip.inputs[0]
            if sm.op_type != "Softmax" or sm.attrs.get("axis", -1) not in (-1, 1) or not private(sm, clip):
                continue
            logits = sm.inputs[0]
            # both operands must exist when the softmax node is reached (the decision point)
            if labels.id in pos and labels.id not in fed and pos[labels.id] > pos[sm.id]:
                continue
            interior = [sm.id, clip.id, log.id, mul.id, red.id]
            if taken & set(interior + [n.id]):
                continue
            xent.append([n.id, sm.id, logits.id, labels.id, float(lo), interior])
            taken.update(interior + [n.id])
    if not relu_pairs and not xent and not affine and not mse:
        return None
    return {"relu": relu_pairs, "affine": affine, "mse": mse, "xent": xent}


class FusionState:
    """Runtime form of the planned rewrites for one run on one task."""
    __slots__ = ("xwb", "relu", "softmax", "interior", "neg", "active", "aff_first", "aff_interior", "aff_last", "mse_first", "mse_interior",
                 "mse_last", "ids")

    def __init__(self, wire: Optional[Dict[str, list]]):
        self.xwb: Dict[int, int] = {}
        self.relu: Dict[int, int] = {}
        self.softmax: Dict[int, tuple] = {}
        self.interior: Dict[int, int] = {}
        self.neg: Dict[int, tuple] = {}
        self.active: Dict[int, bool] = {}
        self.aff_first: Dict[int, tuple] = {}        # MatMul id -> spec (the decision point)
        self.aff_interior: Dict[int, int] = {}       # Add id (when a Relu closes the pattern) -> MatMul id
        self.aff_last: Dict[int, tuple] = {}         # id of the node that receives the fused value -> spec
        for mm, add, relu, x, w, b in (wire or {}).get("affine", ()):
            spec = (int(mm), int(add), int(relu), int(x), int(w), int(b))
            self.aff_first[int(mm)] = spec
            if relu >= 0:
                self.aff_interior[int(add)] = int(mm)
                self.aff_last[int(relu)] = spec
            else:
                self.aff_last[int(add)] = spec
        self.mse_first: Dict[int, tuple] = {}        # Sub id -> spec (the decision point)
        self.mse_interior: Dict[int, int] = {}       # Square id -> Sub id
        self.mse_last: Dict[int, tuple] = {}         # Mean id -> spec
        for mean, sq, sub, a, b in (wire or {}).get("mse", ()):
            spec = (int(mean), int(sq), int(sub), int(a), int(b))
            self.mse_first[int(sub)] = spec
            self.mse_interior[int(sq)] = int(sub)
            self.mse_last[int(mean)] = spec
        for x, r in (wire or {}).get("relu", ()):
            self.xwb[int(x)] = int(r)
            self.relu[int(r)] = int(x)
        for neg, sm, logits, labels, lo, interior in (wire or {}).get("xent", ()):
            spec = (int(neg), int(sm), int(logits), int(labels), float(lo))
            self.softmax[int(sm)] = spec
            self.neg[int(neg)] = spec
            for i in interior:
                self.interior[int(i)] = int(sm)

        # every node id a rewrite touches: the executor asks try_execute() only for these
        self.ids = frozenset(list(self.xwb) + list(self.relu) + list(self.softmax) + list(self.interior) + list(self.neg)
                             + list(self.aff_first) + list(self.aff_interior) + list(self.aff_last)
                             + list(self.mse_first) + list(self.mse_interior) + list(self.mse_last))

    def __bool__(self) -> bool:
        return bool(self.xwb or self.softmax or self.aff_first or self.mse_first)


def try_execute(node, ctx, values: Dict[int, Any], st: FusionState, dev, want_grad: bool):
    """Returns ``(handled, out)``: ``handled`` False = run the node's own kernel.  ``out`` None with ``handled`` True = the node is
    interior to an active fusion and has no value of its own."""
    nid = node.id
    from ..ops import native
    if nid in st.xwb:
        x, w, b = (values[i.id] for i in node.inputs)
        ok = (isinstance(x, torch.Tensor) and isinstance(w, torch.Tensor) and isinstance(b, torch.Tensor) and x.dim() == 2 and w.dim() == 2
              and x.is_floating_point())
        st.active[nid] = ok
        if not ok:
            return False, None
        x, w, b = (v.to(dev, non_blocking=True) if (dev is not None and v.device != dev) else v for v in (x, w, b))
        with (torch.enable_grad() if want_grad else torch.no_grad()):
            return True, native.linear(x, w, b, relu=True)
    if nid in st.relu:
        src = st.relu[nid]
        if st.active.get(src):
            return True, values[src]                       # the producer's epilogue already applied the activation
        return False, None
    if nid in st.aff_first:
        mm, add, relu, xid, wid, bid = st.aff_first[nid]
        x, w, b = values.get(xid), values.get(wid), values.get(bid)
        ok = (isinstance(x, torch.Tensor) and isinstance(w, torch.Tensor) and isinstance(b, torch.Tensor) and x.dim() == 2 and w.dim() == 2
              and x.dtype == torch.float32 and w.dtype == torch.float32 and b.dtype == torch.float32
              and (b.numel() == 1 or tuple(b.shape) == (w.shape[1],)))
        st.active[nid] = ok
        return (True, None) if ok else (False, None)
    if nid in st.aff_interior:
        return (True, None) if st.active.get(st.aff_interior[nid]) else (False, None)
    if nid in st.aff_last:
        mm, add, relu, xid, wid, bid = st.aff_last[nid]
        if not st.active.get(mm):
            return False, None
        x, w, b = values[xid], values[wid], values[bid]
        if dev is not None:
            x, w, b = (v.to(dev, non_blocking=True) if v.device != dev else v for v in (x, w, b))
        with (torch.enable_grad() if want_grad else torch.no_grad()):
            if b.numel() == 1 and tuple(b.shape) != (w.shape[1],):
                b = b.reshape(1).expand(w.shape[1])              # a scalar bias: one value per output column (its gradient sums back)
            return True, native.linear(x, w, b.contiguous(), relu=relu >= 0)
    if nid in st.mse_first:
        _, _, _, aid, bid = st.mse_first[nid]
        a, b = values.get(aid), values.get(bid)
        ok = (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dtype == torch.float32 and b.dtype == torch.float32
              and a.dim() > 0 and b.dim() > 0 and a.numel() > 0 and b.numel() > 0)
        st.active[nid] = ok
        return (True, None) if ok else (False, None)
    if nid in st.mse_interior:
        return (True, None) if st.active.get(st.mse_interior[nid]) else (False, None)
    if nid in st.mse_last:
        _, _, sub, aid, bid = st.mse_last[nid]
        if not st.active.get(sub):
            return False, None
        a, b = values[aid], values[bid]
        if dev is not None:
            a, b = (v.to(dev, non_blocking=True) if v.device != dev else v for v in (a, b))
        with (torch.enable_grad() if want_grad else torch.no_grad()):
            return True, native.mse(a, b)
    if nid in st.softmax:
        _, _, lid, yid, _ = st.softmax[nid]
        logits, labels = values.get(lid), values.get(yid)
        ok = (isinstance(logits, torch.Tensor) and isinstance(labels, torch.Tensor) and logits.dim() == 2 and logits.dtype == torch.float32
              and labels.dtype == torch.float32 and tuple(labels.shape) == tuple(logits.shape))
        st.active[nid] = ok
        return (True, None) if ok else (False, None)
    if nid in st.interior:
        return (True, None) if st.active.get(st.interior[nid]) else (False, None)
    if nid in st.neg:
        _, sm, lid, yid, lo = st.neg[nid]
        if not st.active.get(sm):
            return False, None
        logits, labels = values[lid], values[yid]
        if dev is not None:
            logits = logits.to(dev, non_blocking=True) if logits.device != dev else logits
            labels = labels.to(dev, non_blocking=True) if labels.device != dev else labels
        with (torch.enable_grad() if want_grad else torch.no_grad()):
            return True, native.clipped_softmax_xent_sum(logits, labels, lo)
    return False, None
