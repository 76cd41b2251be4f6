"""The reference's MNIST MLP (``distributed_mnist.py:96-113``) as reusable builders.

``build_mnist_mlp`` creates the graph-API model (variables ``hid_w, hid_b, sm_w, sm_b`` in the reference's
creation order so ``replica_device_setter`` places them identically); ``fused=True`` emits the fused
``clipped_softmax_xent_sum`` node (one kernel on GPU) instead of the softmax/clip/log/mul/sum chain.
The fabric fast path for this model is :class:`parallel.ps_engine.PSTrainEngine`.
"""
from __future__ import annotations

import math
from typing import List, Tuple

__all__ = ["build_mnist_mlp", "mnist_mlp_param_shapes"]


def mnist_mlp_param_shapes(hidden: int = 100, in_dim: int = 784, classes: int = 10) -> List[Tuple[str, Tuple[int, ...]]]:
    return [("hid_w", (in_dim, hidden)), ("hid_b", (hidden,)), ("sm_w", (hidden, classes)), ("sm_b", (classes,))]


def build_mnist_mlp(hidden: int = 100, in_dim: int = 784, classes: int = 10, fused: bool = False, seed=None):
    import distributed_tensorflow_b200 as dtf
    global_step = dtf.train.get_or_create_global_step()
    hid_w = dtf.Variable(dtf.truncated_normal([in_dim, hidden], stddev=1.0 / math.sqrt(in_dim), seed=seed), name="hid_w")
    hid_b = dtf.Variable(dtf.zeros([hidden]), name="hid_b")
    sm_w = dtf.Variable(dtf.truncated_normal([hidden, classes], stddev=1.0 / math.sqrt(hidden),
                                             seed=None if seed is None else seed + 1), name="sm_w")
    sm_b = dtf.Variable(dtf.zeros([classes]), name="sm_b")
    x = dtf.placeholder(dtf.float32, [None, in_dim], name="x")
    y_ = dtf.placeholder(dtf.float32, [None, classes], name="y_")
    hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    logits = dtf.nn.xw_plus_b(hid, sm_w, sm_b)
    y = dtf.nn.softmax(logits)
    if fused:
        loss = dtf.nn.clipped_softmax_xent_sum(logits, y_)
    else:
        loss = -dtf.reduce_sum(y_ * dtf.log(dtf.clip_by_value(y, 1e-10, 1.0)))
    return {"global_step": global_step, "x": x, "y_": y_, "y": y, "logits": logits, "loss": loss,
            "vars": (hid_w, hid_b, sm_w, sm_b)}
