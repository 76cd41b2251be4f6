"""Fused batch-norm kernels (csrc/nn_kernels.cu) vs the plain PyTorch fp32 formulation, forward and backward.

First hardware run: round 2 (28 of 29 kernel-level cases green on a B200; the 29th compared two bf16 paths with a
tolerance tighter than bf16 noise, see ``test_resnet18_loss_and_grads_with_fused_bn``).  The fused kernels are now the
default (``DTF_FUSED_NN=0`` switches back to the element-wise PyTorch formulation)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("rows,C", [(65536, 64), (1000, 128), (4096, 256), (37, 512), (64, 8), (5000, 132)])
@pytest.mark.parametrize("relu,with_res", [(False, False), (True, False), (True, True)])
def test_fused_bn_matches_reference(rows, C, relu, with_res):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib, native
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 1.5 + 0.3).cuda()
    scale, offset = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    res = torch.randn(rows, C, generator=g).cuda() if with_res else None
    dy = torch.randn(rows, C, generator=g).cuda()
    n0 = cuda_lib.launch_count()
    y, mean, rstd = cuda_lib.bn_forward(x, scale, offset, res, relu, 1e-5)
    dx, dscale, doffset, dres = cuda_lib.bn_backward(dy, y if relu else None, x, mean, rstd, scale, with_res)
    assert cuda_lib.launch_count() - n0 == 4
    xd = x.double()
    m = xd.mean(0)
    r = torch.rsqrt(xd.var(0, unbiased=False) + 1e-5)
    torch.testing.assert_close(mean.double(), m, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rstd.double(), r, rtol=1e-5, atol=1e-6)
    want = native.bn_train_reference(x, scale, offset, res, relu)
    torch.testing.assert_close(y, want, rtol=1e-4, atol=1e-4)
    wdx, wds, wdo, wdr = native.bn_backward_reference(dy.double(), want.double(), xd, m, r, scale.double(), relu)
    torch.testing.assert_close(dx.double(), wdx, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dscale.double(), wds, rtol=1e-3, atol=1e-2 * max(1.0, rows ** 0.5 / 16))
    torch.testing.assert_close(doffset.double(), wdo, rtol=1e-3, atol=1e-2 * max(1.0, rows ** 0.5 / 16))
    if with_res:
        torch.testing.assert_close(dres.double(), wdr, rtol=0, atol=0)
    # replayable: a second call with the same workspace/tickets gives the same statistics
    _, mean2, rstd2 = cuda_lib.bn_forward(x, scale, offset, res, relu, 1e-5)
    assert torch.equal(mean, mean2) and torch.equal(rstd, rstd2)


def test_resnet18_loss_and_grads_with_fused_bn(monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.models.resnet import resnet18_init, resnet18_loss
    from distributed_tensorflow_b200.ops import native
    init = {k: v.cuda().requires_grad_() for k, v in resnet18_init(seed=2).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 32, 32, 3, generator=g).cuda()
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=g), 10).float().cuda()

    from distributed_tensorflow_b200.ops import cuda_lib

    def run(fused):
        monkeypatch.setattr(native, "_FUSED_BN", fused)
        monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
        loss = resnet18_loss(init, x, y)
        grads = torch.autograd.grad(loss, list(init.values()))
        return float(loss), grads
    l0, g0 = run(False)
    l1, g1 = run(True)
    # Measured on hardware (tools/resnet_grad_check.py, batch 32): with bf16 GEMM operands BOTH paths sit ~0.3 norm-wise from
    # the pure-fp32 gradients of this random-init network (a rounding decision that flips early is amplified by 18
    # conv + batch-norm layers), so the two paths differ from each other by a similar amount: loss tight, gradients within
    # that measured noise level
    native._FORCE_EAGER = True
    try:
        lref = float(resnet18_loss(init, x, y))
        gref = torch.autograd.grad(resnet18_loss(init, x, y), list(init.values()))
    finally:
        native._FORCE_EAGER = False
    assert abs(l0 - l1) < 2e-3 * max(1.0, abs(l0)) and abs(l1 - lref) < 2e-2 * abs(lref)
    e0 = [float((a - r).norm() / (r.norm() + 1e-12)) for a, r in zip(g0, gref)]
    e1 = [float((b - r).norm() / (r.norm() + 1e-12)) for b, r in zip(g1, gref)]
    assert max(e1) < 0.6 and sorted(e1)[len(e1) // 2] < 1.25 * sorted(e0)[len(e0) // 2] + 0.02, (max(e0), max(e1))


@pytest.mark.parametrize("shape,k,stride", [((8, 32, 32, 64), 3, 1), ((4, 16, 16, 128), 3, 2), ((2, 9, 7, 256), 1, 2), ((3, 8, 8, 8), 3, 1)])
def test_vector_im2col_col2im_match_scalar_kernels(shape, k, stride, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib
    g = torch.Generator().manual_seed(k + stride)
    x = torch.randn(*shape, generator=g).cuda()
    pads = (1, 1, 1, 1) if k == 3 else (0, 0, 0, 0)
    outs = []
    for fused in (False, True):
        monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
        cols, (n, ho, wo) = cuda_lib.im2col_nhwc(x, k, k, (stride, stride), pads)
        gc = torch.randn(cols.shape, generator=torch.Generator().manual_seed(1)).cuda()
        gx = cuda_lib.col2im_nhwc(gc, x.shape, k, k, (stride, stride), pads)
        outs.append((cols, gx))
    assert torch.equal(outs[0][0], outs[1][0])
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape,k,stride", [((8, 32, 32, 64), 3, 2), ((4, 16, 16, 128), 2, 2), ((2, 7, 9, 8), 3, 1)])
def test_pooling_kernels_match_pytorch(shape, k, stride, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib, native
    monkeypatch.setattr(cuda_lib, "FUSED_NN", True)
    g = torch.Generator().manual_seed(k)
    x = torch.randn(*shape, generator=g).cuda()

    def both(fn):
        outs = []
        for fused in (True, False):
            monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
            leaf = x.clone().requires_grad_()
            y = fn(leaf)
            (gx,) = torch.autograd.grad((y * y).sum(), leaf)
            outs.append((y.detach(), gx))
        return outs
    (y1, g1), (y0, g0) = both(lambda t: native.max_pool_nhwc(t, (1, k, k, 1), (1, stride, stride, 1), "SAME"))
    assert torch.equal(y1, y0)
    torch.testing.assert_close(g1, g0, rtol=1e-6, atol=1e-6)
    (y1, g1), (y0, g0) = both(native.global_avg_pool)
    torch.testing.assert_close(y1, y0, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(g1, g0, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("shape,k,cout", [((8, 32, 32, 64), 3, 64), ((4, 16, 16, 128), 3, 128), ((2, 8, 8, 16), 2, 24)])
def test_stride1_data_gradient_as_convolution_matches_col2im_path(shape, k, cout, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib, native
    g = torch.Generator().manual_seed(k)
    x = torch.randn(*shape, generator=g).cuda()
    w = (torch.randn(k, k, shape[-1], cout, generator=g) * 0.1).cuda()
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(cuda_lib, "FUSED_NN", fused)
        xl, wl = x.clone().requires_grad_(), w.clone().requires_grad_()
        y = native.conv2d_nhwc(xl, wl, (1, 1, 1, 1), "SAME")
        outs.append(torch.autograd.grad((y * y).sum(), [xl, wl]))
    for a, b in zip(*outs):
        assert float((a - b).norm() / (b.norm() + 1e-12)) < 1e-2
