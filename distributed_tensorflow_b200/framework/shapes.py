"""Static shapes with unknown (``None``) dimensions.

``standalone.py:49,56`` reads ``input_tensor.get_shape()[1]`` to size weight
matrices, so builders propagate best-effort static shapes.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

__all__ = ["TensorShape", "Dimension", "broadcast_shape", "matmul_shape", "reduce_shape"]


class Dimension:
    __slots__ = ("value",)

    def __init__(self, value: Optional[int]):
        self.value = None if value is None else int(value)

    def __int__(self) -> int:
        if self.value is None:
            raise ValueError("dimension is unknown")
        return self.value

    __index__ = __int__

    def __eq__(self, other) -> bool:
        o = other.value if isinstance(other, Dimension) else other
        return self.value == o

    def __hash__(self) -> int:
        return hash(self.value)

    def __repr__(self) -> str:
        return "Dimension(%s)" % ("?" if self.value is None else self.value)

    def __mul__(self, other):
        o = other.value if isinstance(other, Dimension) else other
        return Dimension(None if self.value is None or o is None else self.value * o)

    __rmul__ = __mul__


class TensorShape:
    def __init__(self, dims: Optional[Iterable[Optional[int]]]):
        if isinstance(dims, TensorShape):
            dims = dims._dims
        self._dims: Optional[Tuple[Optional[int], ...]] = None if dims is None else tuple(
            (d.value if isinstance(d, Dimension) else (None if d is None else int(d))) for d in dims)

    @property
    def ndims(self) -> Optional[int]:
        return None if self._dims is None else len(self._dims)

    @property
    def dims(self) -> Optional[List[Dimension]]:
        return None if self._dims is None else [Dimension(d) for d in self._dims]

    def as_list(self) -> List[Optional[int]]:
        if self._dims is None:
            raise ValueError("as_list() on an unknown shape")
        return list(self._dims)

    def is_fully_defined(self) -> bool:
        return self._dims is not None and all(d is not None for d in self._dims)

    def num_elements(self) -> Optional[int]:
        if not self.is_fully_defined():
            return None
        n = 1
        for d in self._dims:
            n *= d
        return n

    def __len__(self) -> int:
        if self._dims is None:
            raise ValueError("len() of an unknown shape")
        return len(self._dims)

    def __getitem__(self, i):
        if self._dims is None:
            return Dimension(None)
        if isinstance(i, slice):
            return TensorShape(self._dims[i])
        return Dimension(self._dims[i])

    def __iter__(self):
        if self._dims is None:
            raise ValueError("iterating an unknown shape")
        return iter(Dimension(d) for d in self._dims)

    def __eq__(self, other) -> bool:
        o = other._dims if isinstance(other, TensorShape) else (None if other is None else tuple(other))
        return self._dims == o

    def __repr__(self) -> str:
        if self._dims is None:
            return "TensorShape(None)"
        return "TensorShape([%s])" % ", ".join("?" if d is None else str(d) for d in self._dims)


def broadcast_shape(a: Optional[Sequence], b: Optional[Sequence]) -> Optional[Tuple]:
    if a is None or b is None:
        return None
    out = []
    for i in range(1, max(len(a), len(b)) + 1):
        da = a[-i] if i <= len(a) else 1
        db = b[-i] if i <= len(b) else 1
        if da == 1:
            out.append(db)
        elif db == 1:
            out.append(da)
        elif da is None or db is None:
            out.append(da if db is None else db)
        else:
            out.append(da)
    return tuple(reversed(out))


def matmul_shape(a: Optional[Sequence], b: Optional[Sequence], ta: bool = False, tb: bool = False) -> Optional[Tuple]:
    if a is None or b is None or len(a) != 2 or len(b) != 2:
        return None
    m = a[1] if ta else a[0]
    n = b[0] if tb else b[1]
    return (m, n)


def reduce_shape(shape: Optional[Sequence], axis, keepdims: bool) -> Optional[Tuple]:
    if shape is None:
        return None
    if axis is None:
        return tuple(1 for _ in shape) if keepdims else ()
    axes = [axis] if isinstance(axis, int) else list(axis)
    axes = [a % len(shape) for a in axes]
    out = []
    for i, d in enumerate(shape):
        if i in axes:
            if keepdims:
                out.append(1)
        else:
            out.append(d)
    return tuple(out)
