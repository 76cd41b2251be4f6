"""A minimal ``tf.data``: the input-pipeline spelling TF-1.x programs use instead of ``feed_dict``
(``Dataset.from_tensor_slices(...).shuffle(...).repeat().batch(...)`` -> ``make_one_shot_iterator().get_next()``).

The reference feeds ``mnist.train.next_batch`` through placeholders (``/root/reference/distributed_mnist.py:149-152``); this is the
same role for programs written the other way.  Pipelines are host-side and lazy (Python generators over numpy arrays); the
``get_next`` tensors are produced by ONE stateful graph op that pulls the next element from the iterator kept in the executing
task's resource store -- so between-graph replicas each advance their own iterator, and the end of the data raises
``OutOfRangeError`` like TF.  The bulk path for the fabric engine is ``utils/input_pipeline.py`` (whole epochs in pinned memory)."""
from __future__ import annotations

import itertools
import threading
from typing import Any, Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..framework import errors
from ..framework.graph import get_default_graph
from ..framework.ops import _node, as_dtype, register_kernel

_COUNTER = itertools.count()


def _flatten(x) -> List[Any]:
    if isinstance(x, (tuple, list)):
        out = []
        for v in x:
            out += _flatten(v)
        return out
    if isinstance(x, dict):
        out = []
        for k in sorted(x):
            out += _flatten(x[k])
        return out
    return [x]


def _unflatten(structure, flat: Iterator):
    if isinstance(structure, (tuple, list)):
        return type(structure)(_unflatten(s, flat) for s in structure)
    if isinstance(structure, dict):
        return {k: _unflatten(structure[k], flat) for k in sorted(structure)}
    return next(flat)


def _map_structure(fn, x):
    if isinstance(x, (tuple, list)):
        return type(x)(_map_structure(fn, v) for v in x)
    if isinstance(x, dict):
        return {k: _map_structure(fn, v) for k, v in x.items()}
    return fn(x)


class _Spec:
    """dtype + static shape (``None`` = unknown dimension) of one component of a dataset's elements."""
    __slots__ = ("dtype", "shape")

    def __init__(self, dtype, shape):
        self.dtype, self.shape = np.dtype(dtype), tuple(shape)


def _spec_of(a) -> _Spec:
    a = np.asarray(a)
    return _Spec(a.dtype, a.shape)


class Dataset:
    """``make()`` returns a fresh python iterator over the elements (numpy arrays, or tuples / dicts of them); ``structure`` has the
    elements' nesting with a :class:`_Spec` per component."""

    def __init__(self, make: Callable[[], Iterator], structure):
        self._make, self._structure = make, structure

    # -- sources ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def from_tensor_slices(tensors) -> "Dataset":
        arrs = _map_structure(np.asarray, tensors)
        flat = _flatten(arrs)
        n = flat[0].shape[0]
        if any(a.shape[0] != n for a in flat):
            raise ValueError("from_tensor_slices: components differ in their first dimension")
        return Dataset(lambda: (_map_structure(lambda a, i=i: a[i], arrs) for i in range(n)),
                       _map_structure(lambda a: _Spec(a.dtype, a.shape[1:]), arrs))

    @staticmethod
    def from_tensors(tensors) -> "Dataset":
        arrs = _map_structure(np.asarray, tensors)
        return Dataset(lambda: iter([arrs]), _map_structure(_spec_of, arrs))

    @staticmethod
    def range(*args) -> "Dataset":
        return Dataset(lambda: (np.asarray(i, np.int64) for i in range(*args)), _Spec(np.int64, ()))

    # -- transformations --------------------------------------------------------------------------------------------------------
    def map(self, fn: Callable, num_parallel_calls=None) -> "Dataset":
        def apply(e):
            out = fn(*e) if isinstance(e, tuple) else fn(e)
            return _map_structure(np.asarray, out)

        def make():
            return (apply(e) for e in self._make())
        probe = next(iter(make()), None)             # the function's output signature, from the first element
        return Dataset(make, _map_structure(_spec_of, probe) if probe is not None else self._structure)

    def filter(self, pred: Callable) -> "Dataset":
        return Dataset(lambda: (e for e in self._make() if bool(pred(*e) if isinstance(e, tuple) else pred(e))), self._structure)

    def shuffle(self, buffer_size: int, seed: Optional[int] = None, reshuffle_each_iteration: bool = True) -> "Dataset":
        """TF's streaming shuffle: keep ``buffer_size`` elements, emit a random one, refill (a buffer >= the data set is a full
        permutation)."""
        epoch = itertools.count()

        def make():
            e = next(epoch) if reshuffle_each_iteration else 0
            rng = np.random.RandomState(None if seed is None else (int(seed) + e) % (2 ** 32))
            buf: List[Any] = []
            for item in self._make():
                if len(buf) < buffer_size:
                    buf.append(item)
                    continue
                j = int(rng.randint(0, len(buf)))
                out, buf[j] = buf[j], item
                yield out
            rng.shuffle(buf)
            yield from buf
        return Dataset(make, self._structure)

    def repeat(self, count: Optional[int] = None) -> "Dataset":
        def make():
            n = 0
            while count is None or count < 0 or n < count:
                empty = True
                for e in self._make():
                    empty = False
                    yield e
                if empty:
                    return
                n += 1
        return Dataset(make, self._structure)

    def batch(self, batch_size: int, drop_remainder: bool = False) -> "Dataset":
        def stack(items):
            flats = [_flatten(i) for i in items]
            cols = [np.stack([f[c] for f in flats]) for c in range(len(flats[0]))]
            return _unflatten(items[0], iter(cols))

        def make():
            cur = []
            for e in self._make():
                cur.append(e)
                if len(cur) == batch_size:
                    yield stack(cur)
                    cur = []
            if cur and not drop_remainder:
                yield stack(cur)
        return Dataset(make, _map_structure(lambda sp: _Spec(sp.dtype, (batch_size if drop_remainder else None,) + sp.shape), self._structure))

    def take(self, count: int) -> "Dataset":
        return Dataset(lambda: itertools.islice(self._make(), count), self._structure)

    def skip(self, count: int) -> "Dataset":
        return Dataset(lambda: itertools.islice(self._make(), count, None), self._structure)

    def prefetch(self, buffer_size: int) -> "Dataset":
        """Elements are produced by a helper thread, up to ``buffer_size`` ahead of the consumer."""
        import queue

        def make():
            q: "queue.Queue" = queue.Queue(maxsize=max(1, int(buffer_size)))
            end = object()

            def work():
                try:
                    for e in self._make():
                        q.put(e)
                    q.put(end)
                except BaseException as ex:      # noqa: BLE001 - re-raised in the consumer
                    q.put(ex)
            threading.Thread(target=work, name="dtf-data-prefetch", daemon=True).start()
            while True:
                e = q.get()
                if e is end:
                    return
                if isinstance(e, BaseException):
                    raise e
                yield e
        return Dataset(make, self._structure)

    # -- iterators -----------------------------------------------------------------------------------------------------------------
    def make_one_shot_iterator(self) -> "Iterator_":
        return Iterator_(self, one_shot=True)

    def make_initializable_iterator(self) -> "Iterator_":
        return Iterator_(self, one_shot=False)


class Iterator_:
    """``get_next()`` -> tensors (same nesting as the elements); ``initializer`` (re)starts the pipeline on the executing task."""

    def __init__(self, dataset: Dataset, one_shot: bool):
        self._ds, self._one_shot = dataset, one_shot
        self._name = "dtf_iterator_%d" % next(_COUNTER)
        self.initializer = _node("IteratorInit", (), {"iterator": self}, self._name + "/init") if not one_shot else None

    def _state(self, ctx):
        return ctx.store.get_resource(self._name, lambda: {"it": None, "lock": threading.Lock()})

    def get_next(self, name: str = "IteratorGetNext"):
        flat = _flatten(self._ds._structure)
        node = _node("IteratorGetNext", (), {"iterator": self, "n": len(flat)}, name, None, None)
        outs = [_node("TupleItemRaw", (node,), {"index": i}, "%s_%d" % (name, i), torch.from_numpy(np.zeros(1, sp.dtype)).dtype, sp.shape)
                for i, sp in enumerate(flat)]
        return _unflatten(self._ds._structure, iter(outs))


@register_kernel("IteratorInit", stateful=True)
def _k_iter_init(ctx, node):
    it = node.attrs["iterator"]
    st = it._state(ctx)
    with st["lock"]:
        st["it"] = it._ds._make()
    return None


@register_kernel("IteratorGetNext", stateful=True)
def _k_iter_next(ctx, node):
    it = node.attrs["iterator"]
    st = it._state(ctx)
    with st["lock"]:
        if st["it"] is None:
            if not it._one_shot:
                raise errors.FailedPreconditionError("GetNext() failed because the iterator has not been initialized: run "
                                                     "iterator.initializer first")
            st["it"] = it._ds._make()
        try:
            e = next(st["it"])
        except StopIteration:
            raise errors.OutOfRangeError("End of sequence")
    dev = ctx.torch_device(node)
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) if isinstance(a, np.ndarray) and a.ndim
                 else torch.as_tensor(np.asarray(a)).to(dev) for a in _flatten(e))
