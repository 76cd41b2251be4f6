"""Closed-form batch-norm backward used by the fused kernels (csrc/nn_kernels.cu) vs autograd of the plain formulation,
on CPU.  The CUDA kernels themselves are checked against the same references in tests/test_gpu_nn_fused.py."""
import pytest
import torch

from distributed_tensorflow_b200.ops import native


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_bn_backward_closed_form_matches_autograd(relu, with_res):
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(6, 5, 5, 8, generator=g) * 2 + 0.5).double().requires_grad_()
    scale = (torch.rand(8, generator=g) + 0.5).double().requires_grad_()
    offset = torch.randn(8, generator=g).double().requires_grad_()
    res = torch.randn(6, 5, 5, 8, generator=g).double().requires_grad_() if with_res else None
    eps = 1e-5

    # plain formulation in float64 (bn_train_reference casts to float32: redo it here in double for a tight check)
    dims = (0, 1, 2)
    mean = x.mean(dim=dims, keepdim=True)
    var = (x - mean).pow(2).mean(dim=dims, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps) * scale + offset
    if res is not None:
        y = y + res
    if relu:
        y = torch.relu(y)
    dy = torch.randn(y.shape, generator=g).double()
    grads = torch.autograd.grad(y, [x, scale, offset] + ([res] if with_res else []), dy)

    x2 = x.detach().reshape(-1, 8)
    m2 = x2.mean(0)
    rstd = torch.rsqrt(x2.var(0, unbiased=False) + eps)
    dx, dscale, doffset, dres = native.bn_backward_reference(dy.reshape(-1, 8), y.detach().reshape(-1, 8), x2, m2, rstd,
                                                             scale.detach(), relu)
    torch.testing.assert_close(dx.reshape(x.shape), grads[0], rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(dscale, grads[1], rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(doffset, grads[2], rtol=1e-9, atol=1e-9)
    if with_res:
        torch.testing.assert_close(dres.reshape(x.shape), grads[3], rtol=1e-9, atol=1e-9)


def test_batch_norm_train_reference_matches_functional_batch_norm():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 6, 6, 16, generator=g)
    scale, offset = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g)
    res = torch.randn(4, 6, 6, 16, generator=g)
    got = native.batch_norm_train(x, scale, offset, residual=res, relu=True)
    want = torch.nn.functional.batch_norm(x.permute(0, 3, 1, 2), None, None, scale, offset, training=True, eps=1e-5)
    want = torch.relu(want.permute(0, 2, 3, 1) + res)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
