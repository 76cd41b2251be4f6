"""Host input pipeline for the end-to-end training loop: shuffled epochs of batches in pinned host memory, the next epoch
prepared in the background.

Reference role: ``batch_xs, batch_ys = mnist.train.next_batch(FLAGS.batch_size)`` before every ``mon_sess.run``
(``/root/reference/distributed_mnist.py:149-150``): TF's ``DataSet`` reshuffles the split at every epoch boundary and hands out
consecutive windows.  Here an epoch is materialised ONCE as ``[nb, B, ...]`` arrays -- a row gather by the epoch's permutation,
done natively without the GIL (``csrc/runtime/cpu_kernels.cpp: dtf_gather_rows``) into page-locked memory -- so that
``PSTrainEngine.train_loop`` can copy batch after batch to the device straight from it, and the gather of epoch e+1 runs on a
helper thread while epoch e trains."""
from __future__ import annotations

import threading
from typing import Optional, Tuple

import numpy as np
import torch

from . import native_runtime


def gather_rows(src: np.ndarray, idx: np.ndarray, out: np.ndarray, threads: int = 4) -> np.ndarray:
    """``out[i] = src[idx[i]]`` over the leading dimension (contiguous arrays of the same dtype and row shape)."""
    assert src.flags.c_contiguous and out.flags.c_contiguous and src.dtype == out.dtype and src.shape[1:] == out.shape[1:]
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    assert out.shape[0] == idx.shape[0]
    lib = native_runtime.load()
    if lib is not None and hasattr(lib, "dtf_gather_rows"):
        row = int(np.prod(src.shape[1:], dtype=np.int64)) * src.dtype.itemsize
        rc = lib.dtf_gather_rows(src.ctypes.data, src.shape[0], idx.ctypes.data, idx.shape[0], row, out.ctypes.data, int(threads))
        if rc == -2:
            raise IndexError("gather_rows: index out of range")
        if rc != 0:
            raise RuntimeError("dtf_gather_rows failed with code %d" % rc)
        return out
    np.take(src, idx, axis=0, out=out)
    return out


class EpochBatcher:
    """Epochs of ``(x [nb, B, ...], y [nb, B, ...])`` torch tensors over host arrays ``images`` / ``labels``.

    ``shuffle``: a fresh permutation per epoch from ``RandomState(seed + epoch)`` (every task that builds the batcher with the
    same seed sees the same epochs, so worker w of W can take batches w, w + W, ... as its shard).  The remainder of the split
    that does not fill a batch is dropped for that epoch (another permutation brings it back).  ``pin``: page-lock the epoch
    buffers (default: when CUDA is available).  Two buffer sets alternate; :meth:`next_epoch` returns the prepared one and starts
    the gather of the following epoch into the other set on a helper thread -- the tensors of epoch e stay valid until the next
    call of ``next_epoch`` (which starts overwriting them with epoch e + 2)."""

    def __init__(self, images: np.ndarray, labels: np.ndarray, batch: int, shuffle: bool = True, seed: int = 0,
                 pin: Optional[bool] = None, background: bool = True, threads: int = 4):
        assert images.shape[0] == labels.shape[0] and batch >= 1 and images.shape[0] >= batch
        self._x = np.ascontiguousarray(images)
        self._y = np.ascontiguousarray(labels)
        self.batch, self.shuffle, self.seed, self.threads = int(batch), bool(shuffle), int(seed), int(threads)
        self.num_batches = self._x.shape[0] // self.batch
        n = self.num_batches * self.batch
        pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._bufs = []
        for _ in range(2):
            bx = torch.empty((n,) + self._x.shape[1:], dtype=torch.from_numpy(self._x[:1]).dtype)
            by = torch.empty((n,) + self._y.shape[1:], dtype=torch.from_numpy(self._y[:1]).dtype)
            if pin:
                bx, by = bx.pin_memory(), by.pin_memory()
            self._bufs.append((bx, by))
        self.pinned = pin
        self.epoch = 0                       # the epoch the NEXT call of next_epoch() returns
        self._background = bool(background)
        self._pending: Optional[threading.Thread] = None
        self._error: Optional[BaseException] = None
        self._start(0)

    def permutation(self, epoch: int) -> np.ndarray:
        n = self.num_batches * self.batch
        if not self.shuffle:
            return np.arange(n, dtype=np.int64)
        return np.random.RandomState((self.seed + epoch) % (2 ** 32)).permutation(self._x.shape[0])[:n].astype(np.int64)

    def _fill(self, epoch: int) -> None:
        try:
            bx, by = self._bufs[epoch % 2]
            perm = self.permutation(epoch)
            gather_rows(self._x, perm, bx.numpy(), self.threads)
            gather_rows(self._y, perm, by.numpy(), self.threads)
        except BaseException as e:      # noqa: BLE001 - surfaced by next_epoch()
            self._error = e

    def _start(self, epoch: int) -> None:
        if self._background:
            self._pending = threading.Thread(target=self._fill, args=(epoch,), name="dtf-epoch-batcher", daemon=True)
            self._pending.start()
        else:
            self._fill(epoch)

    def next_epoch(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._pending is not None:
            self._pending.join()
            self._pending = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err
        e = self.epoch
        bx, by = self._bufs[e % 2]
        self.epoch = e + 1
        self._start(e + 1)
        nb, B = self.num_batches, self.batch
        return bx.view((nb, B) + tuple(bx.shape[1:])), by.view((nb, B) + tuple(by.shape[1:]))
