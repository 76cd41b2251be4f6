"""Random fills for the graph's random ops and initialisers (reference: ``tf.truncated_normal`` /
``tf.random_normal`` / ``tf.random_uniform`` in ``/root/reference/distributed_mnist.py:98-105``,
``example_between_graph.py:50-51``).

ONE stream definition (``csrc/philox.h``: Philox4x32-10, counter-addressed) with three implementations of the same
arithmetic, so that a seeded op draws the same values wherever it is placed (TF's guarantee):

* CUDA tensors: ``philox_fill_kernel`` (``csrc/elementwise.cu``) writes the variable in HBM directly;
* host tensors: ``dtf_cpu_philox_fill`` (``csrc/runtime/cpu_kernels.cpp``);
* without the native runtime: the numpy formulation below (also the tests' oracle).

The first CUDA fill of a process is cross-checked against the host implementation (1024 values per kind); a mismatch is
reported loudly, recorded in ``SELF_TEST`` and the process keeps drawing on the host and copying (the values stay
correct)."""
from __future__ import annotations

import logging
import math
import threading
from typing import Dict, Optional, Sequence

import numpy as np
import torch

UNIFORM, NORMAL, TRUNCATED_NORMAL = 0, 1, 2
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
SELF_TEST: Dict[str, object] = {"state": "not run"}
_SELF_LOCK = threading.Lock()
_log = logging.getLogger("dtf")


def philox_words_numpy(key: int, ctr_lo: int, nblk: int, ctr_hi: int = 0) -> np.ndarray:
    """``[nblk, 4]`` uint32 words of blocks ``ctr_lo .. ctr_lo + nblk - 1`` (vectorised Philox4x32-10)."""
    ctr = (np.arange(nblk, dtype=np.uint64) + np.uint64(ctr_lo & (2 ** 64 - 1)))
    c0, c1 = ctr & _MASK32, ctr >> np.uint64(32)
    c2 = np.full(nblk, ctr_hi & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(nblk, (ctr_hi >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(_M0) * c0, np.uint64(_M1) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & _MASK32, n2, p0 & _MASK32
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _erfinv_central(x: np.ndarray) -> np.ndarray:
    f = np.float32
    w = (-np.log((f(1) - x) * (f(1) + x)) - f(2.5)).astype(np.float32)
    p = np.full_like(w, 2.81022636e-08)
    for c in (3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164, 0.246640727,
              1.50140941):
        p = f(c) + p * w
    return (p * x).astype(np.float32)


def philox_fill_numpy(n: int, kind: int, p0: float, p1: float, key: int, offset: int, stream_id: int = 0) -> np.ndarray:
    """The fill of ``csrc/philox.h`` in numpy float32 arithmetic (oracle + fallback)."""
    f = np.float32
    nblk = (n + 3) // 4
    w = philox_words_numpy(key, offset, nblk, stream_id)
    k = (w >> np.uint32(8)).astype(np.float32)
    s = f(5.9604644775390625e-8)
    p0, p1 = f(p0), f(p1)
    if kind == UNIFORM:
        out = p0 + (p1 - p0) * (k * s)
    elif kind == NORMAL:
        u1 = (k[:, 0::2] + f(1)) * s
        u2 = k[:, 1::2] * s
        r = np.sqrt(f(-2) * np.log(u1)).astype(np.float32)
        t = (f(6.283185307179586) * u2).astype(np.float32)
        out = np.empty((nblk, 4), np.float32)
        out[:, 0::2] = p0 + p1 * (r * np.cos(t).astype(np.float32))
        out[:, 1::2] = p0 + p1 * (r * np.sin(t).astype(np.float32))
    elif kind == TRUNCATED_NORMAL:
        u = (k + f(0.5)) * s
        z = (f(2) * u - f(1)) * f(0.9544997361036416)
        out = p0 + p1 * (f(1.4142135623730951) * _erfinv_central(z.astype(np.float32)))
    else:
        raise ValueError("unknown random kind %r" % (kind,))
    return np.ascontiguousarray(out.astype(np.float32).reshape(-1)[:n])


def _fill_host(n: int, kind: int, p0: float, p1: float, key: int, offset: int, stream_id: int) -> torch.Tensor:
    from ..utils import native_runtime
    lib = native_runtime.load()
    if lib is not None and hasattr(lib, "dtf_cpu_philox_fill"):
        out = torch.empty(n, dtype=torch.float32)
        rc = lib.dtf_cpu_philox_fill(out.data_ptr(), n, key & (2 ** 64 - 1), offset & (2 ** 64 - 1), stream_id & (2 ** 64 - 1),
                                     int(kind), float(p0), float(p1))
        if rc != 0:
            raise RuntimeError("dtf_cpu_philox_fill failed with code %d" % rc)
        return out
    return torch.from_numpy(philox_fill_numpy(n, kind, p0, p1, key, offset, stream_id))


def _device_fill_checked(device: torch.device) -> bool:
    """First CUDA use in this process: the kernel's values against the host implementation's, every kind."""
    with _SELF_LOCK:
        if SELF_TEST["state"] != "not run":
            return SELF_TEST["state"] == "passed"
        from . import cuda_lib
        try:
            worst = 0.0
            for kind, a, b in ((UNIFORM, -1.5, 2.0), (NORMAL, 0.25, 1.5), (TRUNCATED_NORMAL, -0.5, 0.75)):
                dev = cuda_lib.philox_fill(torch.empty(1021, dtype=torch.float32, device=device), kind, a, b, 0x1234567887654321, 77, 3)
                host = _fill_host(1021, kind, a, b, 0x1234567887654321, 77, 3)
                worst = max(worst, float((dev.cpu() - host).abs().max()))
            ok = worst <= 1e-4 and math.isfinite(worst)
            SELF_TEST.update(state="passed" if ok else "failed", max_abs_diff=worst)
        except Exception as e:          # noqa: BLE001 - a broken kernel must not take variable initialisation down with it
            ok = False
            SELF_TEST.update(state="failed", error=repr(e)[:300])
        if not ok:
            _log.error("philox_fill_kernel disagrees with the host implementation (%s): random ops of this process are drawn on "
                       "the host and copied to the device", SELF_TEST)
        return ok


def philox_fill(shape: Sequence[int], kind: int, p0: float, p1: float, key: int, offset: int, device: Optional[torch.device] = None,
                stream_id: int = 0) -> torch.Tensor:
    """fp32 tensor of ``shape`` on ``device`` drawn from stream ``(key, stream_id)`` starting at block ``offset``."""
    shape = tuple(int(d) for d in shape)
    n = 1
    for d in shape:
        n *= d
    device = torch.device(device) if device is not None else torch.device("cpu")
    from . import cuda_lib
    if device.type == "cuda" and _device_fill_checked(device):         # cuda_lib.load() raises when the library is missing
        return cuda_lib.philox_fill(torch.empty(n, dtype=torch.float32, device=device), kind, p0, p1, key, offset, stream_id).view(shape)
    if device.type == "cpu" and cuda_lib.EMULATION:
        return cuda_lib.philox_fill(torch.empty(n, dtype=torch.float32), kind, p0, p1, key, offset, stream_id).view(shape)
    t = _fill_host(n, kind, p0, p1, key, offset, stream_id).view(shape)
    return t.to(device) if device.type != "cpu" else t


def blocks_used(shape: Sequence[int]) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return (n + 3) // 4
