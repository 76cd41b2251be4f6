"""K13 / K9 on the CPU tiers: the Philox stream (csrc/philox.h) in its three implementations -- numpy oracle, the native CPU
fill (csrc/runtime/cpu_kernels.cpp) and ``philox_fill_kernel`` (csrc/elementwise.cu under the host emulation of the CUDA block
model) -- and the ew_* element-wise / reduction kernels forward and backward through ``ops/native.py``, plus the graph-level
behaviour of the random ops (seeded, stateful, placement-independent) and of a linear-regression program on the kernels.
Reference: ``tf.truncated_normal`` etc. (distributed_mnist.py:98-105), ``weight * x + biase`` / ``tf.square`` /
``tf.reduce_mean`` (example_between_graph.py:55-60)."""
import math
import shutil

import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as tf
from distributed_tensorflow_b200.ops import cuda_lib, native, random_ops
from distributed_tensorflow_b200.utils import native_runtime


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cuda_lib.enable_emulation(str(tmp_path_factory.mktemp("emu_lib")))
    yield
    cuda_lib.disable_emulation()


def test_philox_known_answers():
    """Random123's published vectors for philox4x32-10 (counter, key -> four words)."""
    kat = [((0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffffffffffff, 0xffffffffffffffff), 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x85a308d3243f6a88, 0x0370734413198a2e), 0x299f31d0a4093822, (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    lib = native_runtime.load()
    for (lo, hi), key, want in kat:
        assert tuple(int(w) for w in random_ops.philox_words_numpy(key, lo, 1, hi)[0]) == want
        if lib is not None:
            out = np.zeros(4, np.uint32)
            assert lib.dtf_cpu_philox_words(out.ctypes.data, 1, key, lo, hi) == 0
            assert tuple(int(w) for w in out) == want
    # consecutive blocks = consecutive counters
    two = random_ops.philox_words_numpy(7, 41, 2)
    assert (two[1] == random_ops.philox_words_numpy(7, 42, 1)[0]).all()


@pytest.mark.parametrize("kind,p0,p1", [(0, -1.0, 2.0), (1, 0.5, 2.0), (2, 0.0, 0.1)])
def test_fills_agree_across_implementations_and_have_the_right_distribution(emulated, kind, p0, p1):
    n, key, off, sid = 100003, 0x9e3779b97f4a7c15, 12345678901, 2
    ref = random_ops.philox_fill_numpy(n, kind, p0, p1, key, off, sid)
    if native_runtime.load() is not None:
        host = random_ops._fill_host(n, kind, p0, p1, key, off, sid).numpy()
        assert np.abs(host - ref).max() < 5e-6 * max(1.0, abs(p0) + 4 * abs(p1))
    dev = cuda_lib.philox_fill(torch.empty(n), kind, p0, p1, key, off, sid).numpy()           # the kernel source, emulated
    assert np.abs(dev - ref).max() < 5e-6 * max(1.0, abs(p0) + 4 * abs(p1))
    # element i depends only on (key, offset + i // 4, i % 4): a fill that starts 5 blocks later continues the first one
    tail = cuda_lib.philox_fill(torch.empty(n - 20), kind, p0, p1, key, off + 5, sid).numpy()
    assert np.array_equal(tail, dev[20:])
    if kind == 0:
        assert ref.min() >= p0 and ref.max() < p1 and abs(ref.mean() - (p0 + p1) / 2) < 0.01 and abs(ref.std() - (p1 - p0) / math.sqrt(12)) < 0.01
    elif kind == 1:
        assert abs(ref.mean() - p0) < 0.03 and abs(ref.std() - p1) < 0.03
        z = (ref - p0) / p1
        assert 0.66 < (np.abs(z) < 1).mean() < 0.70 and (np.abs(z) > 4).mean() < 3e-4
    else:
        z = (ref - p0) / p1
        assert np.abs(z).max() <= 2.0 and abs(z.mean()) < 0.01 and abs(z.std() - 0.8796) < 0.01     # TF's truncated normal
        exact = math.sqrt(2) * torch.erfinv((2 * torch.from_numpy(((random_ops.philox_words_numpy(key, off, (n + 3) // 4, sid) >> 8)
                                                                   .astype(np.float64).reshape(-1)[:n] + 0.5) * 2.0 ** -24) - 1)
                                            * 0.9544997361036416)
        assert float((torch.from_numpy(z).double() - exact).abs().max()) < 2e-5               # the polynomial erfinv


def test_elementwise_kernels_forward_backward_match_torch(emulated):
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(6, 5, 7, generator=g) + 0.5, torch.rand(6, 5, 7, generator=g) + 0.5
    v, m, s = torch.rand(7, generator=g) + 0.5, torch.rand(5, 7, generator=g) + 0.5, torch.rand(1, generator=g) + 0.5
    n0 = cuda_lib.launch_count()
    for op, fn in native._TORCH_BINARY.items():
        for x, y in ((a, b), (a, v), (v, a), (a, m), (m, a), (a, s), (s, a), (v, s)):
            xs, ys = x.clone().requires_grad_(), y.clone().requires_grad_()
            out = native.binary(op, xs, ys)
            assert isinstance(out.grad_fn, torch.autograd.function.BackwardCFunction), op      # really our autograd function
            w = torch.rand(out.shape, generator=g)
            gx, gy = torch.autograd.grad((out * w).sum(), (xs, ys))
            xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
            ref = fn(xr, yr)
            rx, ry = torch.autograd.grad((ref * w).sum(), (xr, yr))
            torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(gx, rx, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(gy, ry, rtol=1e-4, atol=1e-5)
    for op, fn in native._TORCH_UNARY.items():
        x = (a - 1.0) if op in ("relu", "tanh", "sigmoid", "neg", "square") else a
        xs, xr = x.clone().requires_grad_(), x.clone().requires_grad_()
        out, ref = native.unary(op, xs), fn(xr)
        w = torch.rand(out.shape, generator=g)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.autograd.grad((out * w).sum(), xs)[0], torch.autograd.grad((ref * w).sum(), xr)[0], rtol=1e-4,
                                   atol=1e-5)
    big = torch.rand(300, 257, generator=g) - 0.5                        # many blocks: shuffles + shared memory + atomics
    for mean in (False, True):
        xs, xr = big.clone().requires_grad_(), big.clone().requires_grad_()
        out, ref = native.reduce_all(xs, mean), (xr.mean() if mean else xr.sum())
        assert out.shape == ()
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(torch.autograd.grad(out * 3.0, xs)[0], torch.autograd.grad(ref * 3.0, xr)[0])
    assert cuda_lib.launch_count() - n0 > 150
    # general broadcasts / other dtypes are left to torch
    col = torch.rand(6, 1, 7)
    assert not isinstance(native.binary("add", a.requires_grad_(), col).grad_fn, torch.autograd.function.BackwardCFunction)
    assert native.binary("add", torch.ones(3, dtype=torch.int64), torch.ones(3, dtype=torch.int64)).dtype == torch.int64
    assert cuda_lib.ew_broadcast_mode((4, 1), (4, 3), (4, 3)) is None and cuda_lib.ew_broadcast_mode((1, 3), (4, 3), (4, 3)) == (2, 3)


def _init_values(seed, op_seed=None, runs=1):
    g = tf.Graph()
    with g.as_default():
        if seed is not None:
            tf.set_random_seed(seed)
        w = tf.get_variable("w", [50, 20], initializer=tf.truncated_normal_initializer(stddev=0.1, seed=op_seed))
        u = tf.get_variable("u", [50, 20], initializer=tf.random_uniform_initializer(-1, 1, seed=op_seed))
        r = tf.random_normal([33], seed=op_seed)
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            return sess.run([w, u]) + [sess.run(r) for _ in range(runs)]


def test_graph_random_ops_are_seeded_stateful_and_identical_under_the_kernels(emulated, monkeypatch):
    w1, u1, r1a, r1b = _init_values(5, runs=2)                        # emulation on: philox_fill_kernel draws the values
    assert np.abs(w1).max() <= 0.2 + 1e-6 and abs(w1.std() - 0.088) < 0.01 and not np.array_equal(w1, u1)
    assert not np.array_equal(r1a, r1b)                              # a seeded op continues its stream
    monkeypatch.setattr(cuda_lib, "EMULATION", False)               # same program on the CPU tier's implementation
    w2, u2, r2a, r2b = _init_values(5, runs=2)
    for x, y in ((w1, w2), (u1, u2), (r1a, r2a), (r1b, r2b)):
        np.testing.assert_allclose(x, y, atol=2e-6)
    w3 = _init_values(6)[0]
    assert not np.allclose(w1, w3)
    a, b = _init_values(None)[0], _init_values(None)[0]              # unseeded: entropy
    assert not np.array_equal(a, b)
    np.testing.assert_allclose(_init_values(None, op_seed=9)[0], _init_values(None, op_seed=9)[0], atol=2e-6)


def test_linear_regression_program_trains_on_the_kernels(emulated):
    """example_between_graph.py's model (y = weight * x + biase, mean squared error, SGD) executed by the ew_* kernels under
    the emulation against the same program on torch: same trajectory."""
    def run():
        g = tf.Graph()
        with g.as_default():
            tf.set_random_seed(1)
            x, y_ = tf.placeholder(tf.float32, [None]), tf.placeholder(tf.float32, [None])
            weight = tf.get_variable("weight", [1], tf.float32, initializer=tf.random_normal_initializer())
            biase = tf.get_variable("biase", [1], tf.float32, initializer=tf.random_normal_initializer())
            loss = tf.reduce_mean(tf.square(y_ - (weight * x + biase)))
            train = tf.train.GradientDescentOptimizer(0.1).minimize(loss)
            rng = np.random.RandomState(0)
            out = []
            with tf.Session() as sess:
                sess.run(tf.global_variables_initializer())
                for _ in range(60):
                    xs = rng.rand(16).astype(np.float32)
                    out.append(sess.run([train, loss, weight, biase], {x: xs, y_: 2 * xs + 10})[1:])
            return out
    n0 = cuda_lib.launch_count()
    emu = run()
    assert cuda_lib.launch_count() - n0 > 60 * 8                     # forward + backward element-wise kernels every step
    cuda_lib.EMULATION = False
    try:
        ref = run()
    finally:
        cuda_lib.EMULATION = True
    for (l1, w1, b1), (l2, w2, b2) in zip(emu, ref):
        np.testing.assert_allclose([l1, w1[0], b1[0]], [l2, w2[0], b2[0]], rtol=2e-4, atol=1e-5)
    assert emu[-1][0] < emu[0][0] * 0.05


def test_concat_and_scatter_kernel(emulated):
    """K10: the gather-scatter kernel (copy_parts_kernel) behind ConcatV2 on /gpu and behind the one-launch scatter of a batch."""
    g = torch.Generator().manual_seed(2)
    for shapes, axis in (([(4, 3), (2, 3)], 0), ([(2, 1), (2, 1)], 0), ([(3, 2, 5), (3, 7, 5), (3, 1, 5)], 1), ([(3, 4, 2), (3, 4, 6)], -1),
                         ([(5,), (1,), (9,)], 0)):
        xs = [torch.rand(s, generator=g).requires_grad_() for s in shapes]
        rs = [x.detach().clone().requires_grad_() for x in xs]
        n0 = cuda_lib.launch_count()
        out, ref = native.concat(xs, axis), torch.cat(rs, dim=axis)
        assert cuda_lib.launch_count() == n0 + 1
        torch.testing.assert_close(out, ref, rtol=0, atol=0)
        w = torch.rand(ref.shape, generator=g)
        for a, b in zip(torch.autograd.grad((out * w).sum(), xs), torch.autograd.grad((ref * w).sum(), rs)):
            torch.testing.assert_close(a, b, rtol=0, atol=0)
    # 17 parts, mixed dtypes, mismatched shapes: torch
    many = [torch.rand(1, 2) for _ in range(17)]
    n0 = cuda_lib.launch_count()
    assert native.concat(many, 0).shape == (17, 2) and cuda_lib.launch_count() == n0
    assert native.concat([torch.ones(2, dtype=torch.int64), torch.ones(3, dtype=torch.int64)], 0).dtype == torch.int64
    # scatter of a global batch across per-worker staging buffers (in-graph replication), one launch
    x = torch.rand(10, 3, 4, generator=g)
    outs = [torch.zeros(2, 3, 4), torch.zeros(5, 3, 4), torch.zeros(3, 3, 4)]
    n0 = cuda_lib.launch_count()
    cuda_lib.scatter_rows(x, outs)
    assert cuda_lib.launch_count() == n0 + 1
    torch.testing.assert_close(torch.cat(outs, 0), x, rtol=0, atol=0)
    # the graph op: the in-graph example's split -> per-part work -> concat
    gr = tf.Graph()
    with gr.as_default():
        a = tf.constant(np.arange(12, dtype=np.float32).reshape(4, 3))
        parts = tf.split(a, 2)
        y = tf.concat([tf.matmul(p, tf.constant([[1.0], [1.0], [1.0]])) for p in parts], 0)
        with tf.Session() as sess:
            assert sess.run(y).tolist() == [[3.0], [12.0], [21.0], [30.0]]


def test_first_use_self_tests_run_clean_under_the_emulation(emulated, monkeypatch):
    """The checks a process runs before it trusts the new kernels on hardware, executed here against the emulated kernels;
    and their failure path: a wrong kernel is reported, recorded and the ops fall back to torch."""
    monkeypatch.setitem(native.EW_SELF_TEST, "state", "not run")
    assert native._ew_self_test(torch.device("cpu")) and native.EW_SELF_TEST["state"] == "passed"
    assert native.EW_SELF_TEST["max_abs_diff"] < 1e-4
    monkeypatch.setitem(random_ops.SELF_TEST, "state", "not run")
    assert random_ops._device_fill_checked(torch.device("cpu")) and random_ops.SELF_TEST["state"] == "passed"
    # failure path
    monkeypatch.setitem(native.EW_SELF_TEST, "state", "not run")
    monkeypatch.setattr(cuda_lib, "ew_affine", lambda x, alpha, beta=0.0, out_shape=None: torch.zeros_like(x))
    assert not native._ew_self_test(torch.device("cpu")) and native.EW_SELF_TEST["state"] == "failed"
    monkeypatch.setattr(cuda_lib, "EMULATION", False)
    assert not native._ew_ok(torch.ones(3))
    monkeypatch.setitem(random_ops.SELF_TEST, "state", "not run")
    monkeypatch.setattr(cuda_lib, "philox_fill", lambda out, *a, **k: out.zero_())
    assert not random_ops._device_fill_checked(torch.device("cpu")) and random_ops.SELF_TEST["state"] == "failed"
    monkeypatch.setitem(native.EW_SELF_TEST, "state", "passed")
    monkeypatch.setitem(random_ops.SELF_TEST, "state", "not run")


def test_concat_scatter_and_broadcast_kernels_on_random_shapes(emulated):
    """Random ranks / axes / part lengths (zero-length parts, non-contiguous inputs) for the gather-scatter kernel, random
    broadcast patterns for the binary kernel: always what torch computes."""
    rng = np.random.RandomState(7)
    g = torch.Generator().manual_seed(7)
    for _ in range(40):
        nd = int(rng.randint(1, 5))
        shape = [int(rng.randint(1, 6)) for _ in range(nd)]
        ax = int(rng.randint(-nd, nd))
        parts = []
        for _ in range(int(rng.randint(2, 7))):
            s = list(shape)
            s[ax] = int(rng.randint(0, 5))
            t = torch.rand(s, generator=g)
            if nd >= 2 and rng.randint(0, 3) == 0:
                t = t.transpose(0, nd - 1).contiguous().transpose(0, nd - 1)      # same values, non-contiguous layout
            parts.append(t)
        got = cuda_lib.concat(parts, ax)
        assert got is not None
        torch.testing.assert_close(got, torch.cat(parts, dim=ax), rtol=0, atol=0)
    for _ in range(40):
        nd = int(rng.randint(1, 4))
        out_shape = [int(rng.randint(1, 6)) for _ in range(nd)]
        def operand():
            kind = int(rng.randint(0, 3))
            if kind == 0:
                return torch.rand(out_shape, generator=g) + 0.5
            if kind == 1:
                return torch.rand([1] * int(rng.randint(0, nd + 1)) or [1], generator=g) + 0.5
            k = int(rng.randint(1, nd + 1))
            return torch.rand([1] * int(rng.randint(0, 2)) + out_shape[nd - k:], generator=g) + 0.5
        a, b = operand(), operand()
        op = list(native._TORCH_BINARY)[int(rng.randint(0, 5))]
        try:
            want = native._TORCH_BINARY[op](a, b)
        except RuntimeError:
            continue
        got = native.binary(op, a, b)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
