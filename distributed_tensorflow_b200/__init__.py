"""distributed_tensorflow_b200 -- a B200-native parameter-server training framework.

Same capabilities and ``tf.train``-style surface as the gctian/distributed-tensorflow
examples (ClusterSpec ps/worker roles, between-graph and in-graph replication,
``replica_device_setter``, async and ``SyncReplicasOptimizer`` training,
``MonitoredTrainingSession`` + hooks, name-keyed checkpoints, chrome-trace
timelines), rebuilt for Blackwell: PyTorch tensors, hand-written sm_100a
kernels (tcgen05/TMEM/TMA GEMMs, fused softmax-xent, fused reduce+apply), and
NVLink-5 peer memory for the ps<->worker data plane.

Typical use mirrors the reference scripts::

    import distributed_tensorflow_b200 as dtf
    cluster = dtf.train.ClusterSpec({"ps": [...], "worker": [...]})
    server = dtf.train.Server(cluster, job_name=..., task_index=...)
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device=...)):
        ...
    with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=..., hooks=[...]) as sess:
        while not sess.should_stop():
            sess.run(train_op, feed_dict={...})
"""
from __future__ import annotations

import builtins as _builtins
import types as _types

import torch as _torch

from .framework import errors
from .framework.device import DeviceSpec, device
from .framework.graph import (Graph, GraphKeys, Tensor, add_to_collection, control_dependencies, convert_to_tensor,
                              get_collection, get_default_graph, name_scope, reset_default_graph)
from .framework.ops import *  # noqa: F401,F403  (op builders: matmul, relu, placeholder, ...)
from .framework.ops import (abs, arg_max, argmax, cast, float16, float32, float64, bfloat16, int32, int64, uint8,
                            int8, pow, tuple_ as tuple, gradients, bool_ as bool)  # noqa: F401,A004
from .framework.variables import (AUTO_REUSE, Variable, all_variables, assign, assign_add, assign_sub,
                                  constant_initializer, get_variable, get_variable_scope, global_variables,
                                  global_variables_initializer, glorot_uniform_initializer,
                                  initialize_all_variables, is_variable_initialized, local_variables,
                                  local_variables_initializer, ones_initializer, random_normal_initializer,
                                  random_uniform_initializer, report_uninitialized_variables,
                                  trainable_variables, truncated_normal_initializer, variable_scope,
                                  variables_initializer, variance_scaling_initializer, zeros_initializer)
from .client.session import (ConfigProto, GPUOptions, InteractiveSession, RunMetadata, RunOptions, Session,
                             get_default_session)
from .utils import flags as _flags_mod
from .utils import summary
from .utils.timeline import Timeline

__version__ = "0.1.0"


def set_random_seed(seed: int) -> None:
    """Graph-level seed: makes initialisers reproducible across tasks (same seed => same draws)."""
    get_default_graph().seed = int(seed)


# -- tf.app ------------------------------------------------------------------------------------------
def _app_run(main=None, argv=None):
    import sys
    _flags_mod.FLAGS(sys.argv if argv is None else argv)
    main = main or sys.modules["__main__"].main
    sys.exit(main(sys.argv[:1] + _flags_mod.FLAGS.unparsed))


app = _types.SimpleNamespace(flags=_flags_mod.flags, run=_app_run)
flags = _flags_mod.flags

# -- tf.nn --------------------------------------------------------------------------------------------
from .framework import ops as _ops  # noqa: E402

nn = _types.SimpleNamespace(
    relu=_ops.relu, softmax=_ops.softmax, log_softmax=_ops.log_softmax, xw_plus_b=_ops.xw_plus_b,
    bias_add=_ops.bias_add, sigmoid=_ops.sigmoid, tanh=_ops.tanh, conv2d=_ops.conv2d, max_pool=_ops.max_pool,
    avg_pool=_ops.avg_pool, dropout=_ops.dropout, l2_loss=_ops.l2_loss, moments=_ops.moments,
    batch_normalization=_ops.batch_normalization, fused_batch_norm_train=_ops.fused_batch_norm_train,
    softmax_cross_entropy_with_logits=_ops.softmax_cross_entropy_with_logits,
    sparse_softmax_cross_entropy_with_logits=_ops.sparse_softmax_cross_entropy_with_logits,
    clipped_softmax_xent_sum=_ops.clipped_softmax_xent_sum,
)

# -- dtf.fabric: NVLink parameter-server engines + the graph-API strategy -------------------------------
def _fabric_ns():
    from .parallel.fabric import Fabric
    from .parallel.generic_engine import GenericPSEngine
    from .parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from .parallel.strategy import FabricPSStrategy
    return _types.SimpleNamespace(Fabric=Fabric, GenericPSEngine=GenericPSEngine, EngineConfig=EngineConfig,
                                  MLPSpec=MLPSpec, PSTrainEngine=PSTrainEngine, FabricPSStrategy=FabricPSStrategy)


class _LazyFabric:
    def __getattr__(self, name):
        return getattr(_fabric_ns(), name)


fabric = _LazyFabric()

# -- tf.logging ----------------------------------------------------------------------------------------
def _make_logging():
    import logging as _pylog
    log = _pylog.getLogger("dtf")
    if not log.handlers and not _pylog.getLogger().handlers:
        _pylog.basicConfig(format="%(levelname)s:%(name)s:%(message)s")
    ns = _types.SimpleNamespace(DEBUG=_pylog.DEBUG, INFO=_pylog.INFO, WARN=_pylog.WARNING, ERROR=_pylog.ERROR,
                                FATAL=_pylog.CRITICAL, debug=log.debug, info=log.info, warn=log.warning,
                                warning=log.warning, error=log.error, fatal=log.critical,
                                set_verbosity=log.setLevel, get_verbosity=log.getEffectiveLevel)
    return ns


logging = _make_logging()


# -- tf.layers (the one everybody uses) -----------------------------------------------------------------
def _dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None,
           reuse=None, trainable=True):
    """``tf.layers.dense``: ``activation(inputs . kernel + bias)`` with variables ``<name>/kernel`` and ``<name>/bias``
    (glorot-uniform / zeros by default) created through ``get_variable``, so scopes, reuse and
    ``replica_device_setter`` placement behave as for hand-made variables."""
    x = convert_to_tensor(inputs)
    in_dim = int(x.get_shape()[-1])
    with variable_scope(name or "dense", reuse=reuse):
        kernel = get_variable("kernel", [in_dim, int(units)], initializer=kernel_initializer or glorot_uniform_initializer(),
                              trainable=trainable)
        if use_bias:
            bias = get_variable("bias", [int(units)], initializer=bias_initializer or zeros_initializer(), trainable=trainable)
            y = _ops.xw_plus_b(x, kernel, bias)
        else:
            y = _ops.matmul(x, kernel)
    return activation(y) if activation is not None else y


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _conv2d_layer(inputs, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True,
                  kernel_initializer=None, bias_initializer=None, name=None, reuse=None, trainable=True):
    """``tf.layers.conv2d`` (NHWC): variables ``<name>/kernel`` [kh, kw, in, filters] (glorot uniform) and ``<name>/bias``."""
    x = convert_to_tensor(inputs)
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(strides)
    cin = int(x.get_shape()[-1])
    with variable_scope(name or "conv2d", reuse=reuse):
        kernel = get_variable("kernel", [kh, kw, cin, int(filters)], initializer=kernel_initializer or glorot_uniform_initializer(),
                              trainable=trainable)
        y = _ops.conv2d(x, kernel, strides=(1, sh, sw, 1), padding=padding.upper())
        if use_bias:
            y = _ops.bias_add(y, get_variable("bias", [int(filters)], initializer=bias_initializer or zeros_initializer(), trainable=trainable))
    return activation(y) if activation is not None else y


def _pool_layer(fn):
    def layer(inputs, pool_size, strides, padding="valid", name=None):
        ph, pw = _pair(pool_size)
        sh, sw = _pair(strides)
        return fn(inputs, (1, ph, pw, 1), (1, sh, sw, 1), padding=padding.upper(), name=name or fn.__name__)
    return layer


def _flatten_layer(inputs, name=None):
    x = convert_to_tensor(inputs)
    dims = x.get_shape()[1:]
    n = 1
    for d in dims:
        n *= int(d)
    return _ops.reshape(x, [-1, n], name=name or "flatten")


def _dropout_layer(inputs, rate=0.5, noise_shape=None, seed=None, training=False, name=None):
    """``tf.layers.dropout``: active only when ``training`` (a python bool) is true."""
    return _ops.dropout(inputs, rate=rate, seed=seed, name=name or "dropout") if training else _ops.identity(inputs, name=name or "dropout")


def _batch_norm_layer(inputs, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False, name=None, reuse=None,
                      trainable=True):
    """``tf.layers.batch_normalization`` over the last axis: ``gamma`` / ``beta`` / ``moving_mean`` / ``moving_variance``;
    ``training=True`` normalises with the batch statistics and registers the moving-average updates in
    ``GraphKeys.UPDATE_OPS`` (run them with the train op, as in TF); otherwise the moving statistics are used."""
    x = convert_to_tensor(inputs)
    c = int(x.get_shape()[-1])
    red = list(_builtins.range(len(x.get_shape()) - 1))          # (the module-level name ``range`` is tf.range)
    with variable_scope(name or "batch_normalization", reuse=reuse):
        gamma = get_variable("gamma", [c], initializer=ones_initializer(), trainable=trainable and scale)
        beta = get_variable("beta", [c], initializer=zeros_initializer(), trainable=trainable and center)
        mmean = get_variable("moving_mean", [c], initializer=zeros_initializer(), trainable=False)
        mvar = get_variable("moving_variance", [c], initializer=ones_initializer(), trainable=False)
        if training:
            mean, var = _ops.moments(x, red)
            keep = float(momentum)
            for mov, cur in ((mmean, mean), (mvar, var)):
                upd = assign(mov, _ops.add(_ops.multiply(mov._node, keep), _ops.multiply(_ops.stop_gradient(cur), 1.0 - keep)))
                add_to_collection(GraphKeys.UPDATE_OPS, upd)
            return _ops.batch_normalization(x, mean, var, beta, gamma, epsilon)
        return _ops.batch_normalization(x, mmean, mvar, beta, gamma, epsilon)


layers = _types.SimpleNamespace(dense=_dense, conv2d=_conv2d_layer, max_pooling2d=_pool_layer(_ops.max_pool),
                                average_pooling2d=_pool_layer(_ops.avg_pool), flatten=_flatten_layer, dropout=_dropout_layer,
                                batch_normalization=_batch_norm_layer)

# -- tf.train -----------------------------------------------------------------------------------------
from . import train  # noqa: E402
from .python_compat import input_data, timeline  # noqa: E402,F401

# -- ops next to the reference's own (framework/ops_extra.py); bound last: some names shadow builtins (range, round) ---------------
from .framework import ops_extra as _extra  # noqa: E402

for _n in _extra.__all__:
    globals()[_n] = getattr(_extra, _n)
from .framework import ops_more as _more  # noqa: E402

for _n in _more.__all__:
    globals()[_n] = getattr(_more, _n)
setattr(nn, "softmax_cross_entropy_with_logits_v2", _more.softmax_cross_entropy_with_logits_v2)
for _n in ("relu6", "elu", "leaky_relu", "softplus", "sigmoid_cross_entropy_with_logits", "l2_normalize", "embedding_lookup", "in_top_k",
           "top_k"):
    setattr(nn, _n, getattr(_extra, _n))


def _mean_squared_error(labels, predictions, scope=None):
    return _ops.reduce_mean(_ops.square(_ops.subtract(convert_to_tensor(predictions), convert_to_tensor(labels))), name=scope or "mean_squared_error")


def _softmax_cross_entropy(onehot_labels, logits, scope=None):
    return _ops.reduce_mean(_ops.softmax_cross_entropy_with_logits(labels=onehot_labels, logits=logits), name=scope or "softmax_cross_entropy_loss")


def _sparse_softmax_cross_entropy(labels, logits, scope=None):
    return _ops.reduce_mean(_ops.sparse_softmax_cross_entropy_with_logits(labels=labels, logits=logits),
                            name=scope or "sparse_softmax_cross_entropy_loss")


def _metric_mean(values, weights=None, name="mean"):
    """``tf.metrics.mean``: ``(value, update_op)`` over two LOCAL variables (``total``, ``count``): run ``update_op`` per batch, read
    ``value`` at the end; ``tf.local_variables_initializer()`` resets them."""
    v = cast(convert_to_tensor(values), float32)
    g = get_default_graph()
    with g.name_scope(name):
        total = Variable(lambda: _ops.constant(0.0, dtype=float32), trainable=False, collections=[GraphKeys.LOCAL_VARIABLES], name="total")
        count = Variable(lambda: _ops.constant(0.0, dtype=float32), trainable=False, collections=[GraphKeys.LOCAL_VARIABLES], name="count")
        if weights is not None:
            w = cast(convert_to_tensor(weights), float32)
            num, den = _ops.reduce_sum(_ops.multiply(v, w)), _ops.reduce_sum(_ops.multiply(_ops.ones_like(v), w))
        else:
            num, den = _ops.reduce_sum(v), cast(_extra.size(v), float32)
        new_total, new_count = assign_add(total, num), assign_add(count, den)

        def safe_div(a, b, nm):
            return _extra.where(_ops.greater(b, 0.0), _ops.divide(a, _ops.maximum(b, 1e-12)), _ops.zeros_like(a), name=nm)
        return safe_div(total._node, count._node, "value"), safe_div(new_total, new_count, "update_op")


def _metric_accuracy(labels, predictions, weights=None, name="accuracy"):
    """``tf.metrics.accuracy``: running fraction of ``predictions == labels``."""
    l, p = convert_to_tensor(labels), convert_to_tensor(predictions)
    return _metric_mean(cast(_ops.equal(cast(p, int64), cast(l, int64)), float32), weights, name)


metrics = _types.SimpleNamespace(mean=_metric_mean, accuracy=_metric_accuracy)

from .utils import dataset as _dataset  # noqa: E402

from .utils import tfrecord as _tfrecord  # noqa: E402

data = _types.SimpleNamespace(Dataset=_dataset.Dataset, Iterator=_dataset.Iterator_, TFRecordDataset=_tfrecord.TFRecordDataset)
python_io = _types.SimpleNamespace(TFRecordWriter=_tfrecord.TFRecordWriter, tf_record_iterator=_tfrecord.tf_record_iterator)
io = _types.SimpleNamespace(TFRecordWriter=_tfrecord.TFRecordWriter, tf_record_iterator=_tfrecord.tf_record_iterator,
                            FixedLenFeature=_tfrecord.FixedLenFeature, VarLenFeature=_tfrecord.VarLenFeature,
                            parse_single_example=_tfrecord.parse_single_example, parse_example=_tfrecord.parse_example,
                            decode_raw=_tfrecord.decode_raw)
FixedLenFeature, VarLenFeature = _tfrecord.FixedLenFeature, _tfrecord.VarLenFeature
string = "string"             # dtype tag of byte-string features (host-side: records, parsed features)
parse_single_example, parse_example, decode_raw = _tfrecord.parse_single_example, _tfrecord.parse_example, _tfrecord.decode_raw
for _n in ("Example", "Features", "Feature", "BytesList", "FloatList", "Int64List"):
    setattr(train, _n, getattr(_tfrecord, _n))

from .utils import gfile  # noqa: E402,F401


def _as_bytes(v, encoding="utf-8"):
    return v if isinstance(v, bytes) else str(v).encode(encoding)


def _as_text(v, encoding="utf-8"):
    return v.decode(encoding) if isinstance(v, bytes) else str(v)


compat = _types.SimpleNamespace(as_bytes=_as_bytes, as_text=_as_text, as_str=_as_text, as_str_any=lambda v: _as_text(v) if isinstance(v, bytes) else str(v))
VERSION = __version__


def _is_gpu_available(cuda_only=False, min_cuda_compute_capability=None):
    import torch as _t
    return _t.cuda.is_available() is True


test = _types.SimpleNamespace(is_gpu_available=_is_gpu_available, gpu_device_name=lambda: "/device:GPU:0" if _is_gpu_available() else "",
                              is_built_with_cuda=lambda: True)


def is_tensor(x) -> bool:
    """``tf.contrib.framework.is_tensor`` / ``tf.is_tensor``: a graph tensor or variable (not a numpy array / python value)."""
    from .framework.graph import Tensor as _T
    from .framework.variables import Variable as _V
    return isinstance(x, (_T, _V))


def reduce_logsumexp(x, axis=None, keepdims=False, name="ReduceLogSumExp", reduction_indices=None, keep_dims=None):
    """``log(sum(exp(x)))`` computed stably (the maximum is taken out first; its gradient is blocked, the result's is exact)."""
    if keep_dims is not None:
        keepdims = keep_dims
    ax = axis if axis is not None else reduction_indices
    x = convert_to_tensor(x)
    m = _ops.stop_gradient(_ops.reduce_max(x, axis=ax, keepdims=True))
    out = _ops.add(_ops.log(_ops.reduce_sum(_ops.exp(_ops.subtract(x, m)), axis=ax, keepdims=True)), m)
    if keepdims:
        return _ops.identity(out, name=name)
    return _ops.reduce_sum(out, axis=ax, name=name)          # the kept dimensions have size 1: summing drops them


def truncatediv(x, y, name="TruncateDiv"):
    """Integer division rounding towards zero (``floordiv`` rounds down)."""
    q = _ops.divide(cast(x, float64), cast(y, float64))
    return cast(where(_ops.greater_equal(q, 0.0) if hasattr(_ops, "greater_equal") else greater_equal(q, 0.0), floor(q), ceil(q)),
                convert_to_tensor(x).dtype, name=name)


def model_variables():
    return get_default_graph().get_collection("model_variables")


def moving_average_variables():
    return get_default_graph().get_collection("moving_average_variables")


def add_check_numerics_ops():
    """One op that checks every floating-point tensor of the graph built so far for NaN / Inf (run it next to the train op)."""
    g = get_default_graph()
    checks = []
    for n in list(g.nodes):
        if n.dtype in (float32, float64, float16, bfloat16) and n.op_type not in ("VariableV2", "Placeholder", "Assign", "CheckNumerics"):
            with g.control_dependencies(None):
                checks.append(check_numerics(n, "%s:0" % n.name))
    return group(*checks, name="check_numerics_all")


def timestamp(name="Timestamp"):
    """Seconds since the epoch at the time the op runs (float64 scalar)."""
    return py_func(lambda: __import__("numpy").float64(__import__("time").time()), [], float64, name=name)


initializers = _types.SimpleNamespace(zeros=zeros_initializer, ones=ones_initializer, constant=constant_initializer,
                                      random_normal=random_normal_initializer, truncated_normal=truncated_normal_initializer,
                                      random_uniform=random_uniform_initializer, glorot_uniform=glorot_uniform_initializer,
                                      variance_scaling=variance_scaling_initializer, global_variables=global_variables_initializer,
                                      local_variables=local_variables_initializer, variables=variables_initializer)

losses = _types.SimpleNamespace(mean_squared_error=_mean_squared_error, softmax_cross_entropy=_softmax_cross_entropy,
                                sparse_softmax_cross_entropy=_sparse_softmax_cross_entropy)
del _n

__all__ = [n for n in dir() if not n.startswith("_") and n not in {
    "Any", "Callable", "Dict", "List", "Optional", "Sequence", "Union", "annotations", "np", "torch", "F"}]
