"""utils/input_pipeline.py: the native row gather and the epoch batcher behind ``PSTrainEngine.train_loop`` (reference role:
``mnist.train.next_batch`` per step, distributed_mnist.py:149)."""
import numpy as np
import pytest
import torch

from distributed_tensorflow_b200.utils import input_pipeline as ip
from distributed_tensorflow_b200.utils import native_runtime


def test_gather_rows_native_matches_numpy_and_checks_indices():
    rng = np.random.RandomState(0)
    src = rng.rand(5000, 784).astype(np.float32)
    idx = rng.permutation(5000)[:3000]
    out = np.empty((3000, 784), np.float32)
    ip.gather_rows(src, idx, out, threads=4)                     # > 1 MiB: the threaded path when the native runtime is there
    assert np.array_equal(out, src[idx])
    small = np.empty((7, 784), np.float32)
    assert np.array_equal(ip.gather_rows(src, idx[:7], small, threads=4), src[idx[:7]])
    lab = rng.rand(5000, 10).astype(np.float64)
    out64 = np.empty((3000, 10), np.float64)
    assert np.array_equal(ip.gather_rows(lab, idx, out64), lab[idx])
    if native_runtime.load() is not None:
        with pytest.raises(IndexError):
            ip.gather_rows(src, np.array([0, 5000]), np.empty((2, 784), np.float32))


@pytest.mark.parametrize("background", [True, False])
def test_epoch_batcher_shuffles_per_epoch_covers_every_row_and_is_reproducible(background):
    n, B = 1030, 100
    x = np.arange(n, dtype=np.float32)[:, None] * np.ones((1, 8), np.float32)
    y = np.eye(10, dtype=np.float32)[np.arange(n) % 10]
    b = ip.EpochBatcher(x, y, B, seed=5, pin=False, background=background)
    assert b.num_batches == 10
    e0x, e0y = b.next_epoch()
    assert tuple(e0x.shape) == (10, B, 8) and tuple(e0y.shape) == (10, B, 10)
    rows0 = e0x[:, :, 0].reshape(-1).numpy().astype(np.int64).copy()
    assert len(set(rows0.tolist())) == 1000 and rows0.max() < n          # a permutation prefix: no row twice
    assert np.array_equal(e0y.reshape(-1, 10).numpy(), y[rows0])         # labels travel with their images
    e0_copy = e0x.clone()
    W = 4                                                                # worker shards: batches w, w + W, ... are disjoint and cover the epoch
    shards = [set(e0x[w::W, :, 0].reshape(-1).numpy().astype(np.int64).tolist()) for w in range(W)]
    assert sum(len(s) for s in shards) == 1000 and len(set().union(*shards)) == 1000
    e1x, _ = b.next_epoch()                                              # (epoch 0's buffers are being refilled with epoch 2 from here on)
    rows1 = e1x[:, :, 0].reshape(-1).numpy().astype(np.int64).copy()
    e1_copy = e1x.clone()
    assert not np.array_equal(rows0, rows1)                              # reshuffled
    b2 = ip.EpochBatcher(x, y, B, seed=5, pin=False, background=background)
    again0 = b2.next_epoch()[0].clone()
    again1 = b2.next_epoch()[0].clone()
    assert torch.equal(again0, e0_copy) and torch.equal(again1, e1_copy)
    assert np.array_equal(b.permutation(0)[:1000], rows0)
    # no shuffle: the split in order
    seq = ip.EpochBatcher(x, y, B, shuffle=False, pin=False, background=background)
    assert np.array_equal(seq.next_epoch()[0][:, :, 0].reshape(-1).numpy(), np.arange(1000, dtype=np.float32))


def test_epoch_batcher_feeds_a_training_loop_on_cpu():
    """The consumer contract of ``train_loop`` (contiguous [nb, B, D] fp32 tensors, batch b at x[b]) with a CPU model."""
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    xs, ys = synthetic_mnist(2000, seed=3)
    b = ip.EpochBatcher(xs, ys, 100, seed=1, pin=False)
    w = torch.zeros(784, 10, requires_grad=True)
    losses = []
    for _ in range(3):
        ex, ey = b.next_epoch()
        assert ex.is_contiguous() and ex.dtype == torch.float32 and ex[3].data_ptr() == ex.data_ptr() + 3 * 100 * 784 * 4
        for i in range(b.num_batches):
            loss = torch.nn.functional.cross_entropy(ex[i] @ w, ey[i].argmax(1))
            g, = torch.autograd.grad(loss, w)
            with torch.no_grad():
                w -= 0.5 * g
            losses.append(float(loss.detach()))
    assert losses[-1] < 0.5 * losses[0]
