"""K13 / K9 kernels on hardware (written after the round's GPU budget was spent: the first hardware run is the driver's; the
file sorts last).  CPU coverage of the same sources: tests/test_random_and_elementwise_kernels.py (host emulation)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.cuda.set_device(0)


@pytest.mark.parametrize("kind,p0,p1", [(0, -1.0, 2.0), (1, 0.5, 2.0), (2, 0.0, 0.1)])
def test_philox_fill_kernel_matches_the_host_stream(kind, p0, p1):
    _need_gpu()
    from distributed_tensorflow_b200.ops import cuda_lib, random_ops
    n, key, off, sid = 1_000_003, 0x9e3779b97f4a7c15, 12345678901, 2
    n0 = cuda_lib.launch_count()
    dev = cuda_lib.philox_fill(torch.empty(n, device="cuda"), kind, p0, p1, key, off, sid)
    assert cuda_lib.launch_count() == n0 + 1
    ref = random_ops.philox_fill_numpy(n, kind, p0, p1, key, off, sid)
    assert np.abs(dev.cpu().numpy() - ref).max() < 1e-5 * max(1.0, abs(p0) + 4 * abs(p1))
    t = random_ops.philox_fill((1000, 17), kind, p0, p1, key, off, torch.device("cuda", 0), sid)
    assert t.is_cuda and random_ops.SELF_TEST["state"] == "passed", random_ops.SELF_TEST
    assert np.abs(t.cpu().numpy().reshape(-1) - ref[:17000]).max() < 1e-5 * max(1.0, abs(p0) + 4 * abs(p1))


def test_elementwise_kernels_match_torch_on_device():
    _need_gpu()
    from distributed_tensorflow_b200.ops import cuda_lib, native
    g = torch.Generator().manual_seed(1)
    a, b = (torch.rand(64, 50, 70, generator=g) + 0.5).cuda(), (torch.rand(64, 50, 70, generator=g) + 0.5).cuda()
    v, s = (torch.rand(70, generator=g) + 0.5).cuda(), (torch.rand(1, generator=g) + 0.5).cuda()
    n0 = cuda_lib.launch_count()
    for op, fn in native._TORCH_BINARY.items():
        for x, y in ((a, b), (a, v), (v, a), (a, s), (s, a)):
            xs, ys = x.clone().requires_grad_(), y.clone().requires_grad_()
            out = native.binary(op, xs, ys)
            w = torch.rand(out.shape, generator=g).cuda()
            gx, gy = torch.autograd.grad((out * w).sum(), (xs, ys))
            xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
            ref = fn(xr, yr)
            rx, ry = torch.autograd.grad((ref * w).sum(), (xr, yr))
            torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(gx, rx, rtol=2e-4, atol=1e-3)
            torch.testing.assert_close(gy, ry, rtol=2e-4, atol=1e-3)
    for op, fn in native._TORCH_UNARY.items():
        x = (a - 1.0) if op in ("relu", "tanh", "sigmoid", "neg", "square") else a
        xs, xr = x.clone().requires_grad_(), x.clone().requires_grad_()
        out, ref = native.unary(op, xs), fn(xr)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.autograd.grad(out.sum(), xs)[0], torch.autograd.grad(ref.sum(), xr)[0], rtol=1e-4, atol=1e-5)
    big = (torch.rand(3000, 257, generator=g) - 0.5).cuda()
    for mean in (False, True):
        xs, xr = big.clone().requires_grad_(), big.clone().requires_grad_()
        out, ref = native.reduce_all(xs, mean), (xr.double().mean() if mean else xr.double().sum()).float()
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-3)
        torch.testing.assert_close(torch.autograd.grad(out * 3.0, xs)[0], torch.full_like(big, 3.0 / big.numel() if mean else 3.0))
    assert native.EW_SELF_TEST["state"] == "passed", native.EW_SELF_TEST
    assert cuda_lib.launch_count() - n0 > 100                          # the kernels ran (no torch fallback)


def test_seeded_initialisers_draw_the_same_values_on_gpu_and_cpu():
    _need_gpu()
    import distributed_tensorflow_b200 as tf

    def run(dev):
        g = tf.Graph()
        with g.as_default(), tf.device(dev):
            tf.set_random_seed(11)
            w = tf.get_variable("w", [784, 100], initializer=tf.truncated_normal_initializer(stddev=1.0 / 28))
            u = tf.get_variable("u", [100, 10], initializer=tf.random_uniform_initializer(-1, 1))
            with tf.Session() as sess:
                sess.run(tf.global_variables_initializer())
                return sess.run([w, u])
    (w0, u0), (w1, u1) = run("/cpu:0"), run("/gpu:0")
    np.testing.assert_allclose(w0, w1, atol=2e-6)
    np.testing.assert_allclose(u0, u1, atol=2e-6)


def test_concat_and_scatter_kernel_on_device_and_across_peers():
    _need_gpu()
    from distributed_tensorflow_b200.ops import cuda_lib, native
    g = torch.Generator().manual_seed(2)
    for shapes, axis in (([(400, 30), (200, 30)], 0), ([(30, 20, 50), (30, 70, 50), (30, 1, 50)], 1), ([(3, 4, 2), (3, 4, 6)], -1)):
        xs = [torch.rand(s, generator=g).cuda().requires_grad_() for s in shapes]
        n0 = cuda_lib.launch_count()
        out = native.concat(xs, axis)
        ref = torch.cat([x.detach() for x in xs], dim=axis)
        torch.testing.assert_close(out, ref, rtol=0, atol=0)
        grads = torch.autograd.grad(out.sum(), xs)
        assert all(float(gr.min()) == 1.0 == float(gr.max()) for gr in grads)
    assert native.EW_SELF_TEST["state"] == "passed" and cuda_lib.launch_count() > n0
    x = torch.rand(1000, 784, generator=g).cuda()
    outs = [torch.zeros(300, 784, device="cuda"), torch.zeros(700, 784, device="cuda")]
    cuda_lib.scatter_rows(x, outs)
    torch.testing.assert_close(torch.cat(outs, 0), x, rtol=0, atol=0)
    if torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1):
        lib = cuda_lib.load()
        with torch.cuda.device(0):
            lib.dtf_enable_peer(1)
        far = torch.zeros(700, 784, device="cuda:1")
        near = torch.zeros(300, 784, device="cuda:0")
        torch.cuda.synchronize(1)
        cuda_lib.scatter_rows(x, [near, far])                      # the second part is written over NVLink by GPU 0's kernel
        torch.cuda.synchronize(0)
        torch.testing.assert_close(far.cpu(), x[300:].cpu(), rtol=0, atol=0)
        torch.testing.assert_close(near, x[:300], rtol=0, atol=0)


def test_fused_graph_plans_on_gpu_track_the_cpu_tier():
    """The tower spelling (MatMul + scalar bias + ReLU, mean squared error) and the linear-regression program with the plan-time
    rewrites (framework/fusion.py) on /gpu:0 against the same programs on /cpu:0."""
    _need_gpu()
    import distributed_tensorflow_b200 as tf
    from distributed_tensorflow_b200.ops import cuda_lib

    def tower(dev):
        tf.reset_default_graph()
        with tf.device(dev):
            tf.set_random_seed(4)
            x, t = tf.placeholder(tf.float32, [None, 2]), tf.placeholder(tf.float32, [None, 1])
            h = x
            for i, d in enumerate((64, 32)):
                w = tf.get_variable("affine%d/w" % i, [h.get_shape()[1], d], initializer=tf.truncated_normal_initializer(0, 0.3))
                b = tf.get_variable("affine%d/b" % i, [], initializer=tf.zeros_initializer)
                h = tf.nn.relu(tf.matmul(h, w) + b)
            w = tf.get_variable("affine_last/w", [h.get_shape()[1], 1], initializer=tf.constant_initializer(value=0.1))
            b = tf.get_variable("affine_last/b", [], initializer=tf.zeros_initializer)
            loss = tf.reduce_mean(tf.square(tf.matmul(h, w) + b - t))
            train = tf.train.GradientDescentOptimizer(0.01).minimize(loss)
        rng = np.random.RandomState(0)
        out = []
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            for _ in range(10):
                xs = rng.rand(1000, 2).astype(np.float32)
                out.append(float(sess.run([train, loss], {x: xs, t: xs.sum(1, keepdims=True)})[1]))
        return out
    n0 = cuda_lib.launch_count()
    gpu = tower("/gpu:0")
    assert cuda_lib.launch_count() - n0 >= 10 * 6
    cpu = tower("/cpu:0")
    np.testing.assert_allclose(gpu, cpu, rtol=3e-2)
    assert gpu[-1] < gpu[0]
