"""Implicit-GEMM convolution (csrc/gemm_tcgen05.cu ``conv`` modes: 4-D TMA boxes over the NHWC activation, no patch
matrix) vs ``torch.nn.functional.conv2d`` in fp32 on the same bf16-rounded operands -- forward, data gradient (the same
kernel over dY with the flipped filter) and weight gradient (patches^T . dY, split-K over the pixels)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu]

# (images, H, W, Cin, Cout, k): the four ResNet-18 stages at batch 64 / 8, a widening 3x3, a 1x1, an even kernel
SHAPES = [(64, 32, 32, 64, 64, 3), (8, 32, 32, 64, 64, 3), (64, 16, 16, 128, 128, 3), (64, 8, 8, 256, 256, 3),
          (64, 4, 4, 512, 512, 3), (8, 4, 4, 512, 512, 3), (4, 8, 8, 64, 128, 3), (8, 16, 16, 64, 64, 1),
          (4, 16, 16, 128, 64, 2)]


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("n,h,w,cin,cout,k", SHAPES)
def test_implicit_conv_forward_and_gradients(n, h, w, cin, cout, k, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib, native
    g = torch.Generator().manual_seed(n * 1000 + h * 10 + k)
    x = torch.randn(n, h, w, cin, generator=g).cuda().requires_grad_()
    wt = (torch.randn(k, k, cin, cout, generator=g) * (2.0 / (k * k * cin)) ** 0.5).cuda().requires_grad_()
    dy = torch.randn(n, h, w, cout, generator=g).cuda()
    calls = []
    real = cuda_lib.conv_igemm
    monkeypatch.setattr(cuda_lib, "conv_igemm", lambda *a, **kw: (calls.append(kw.get("wgrad", False)), real(*a, **kw))[1])
    y = native.conv2d_nhwc(x, wt, (1, 1, 1, 1), "SAME")
    y.backward(dy)
    assert calls == [False, True, False], calls          # forward, wgrad, dgrad all took the implicit path
    # oracle: fp32 convolution of the bf16-rounded operands (what the tensor cores multiply), SAME padding as TF splits it
    r = lambda t: t.detach().bfloat16().float()
    xr, wr = r(x).requires_grad_(), r(wt).requires_grad_()
    pt, pb = native._same_pad(h, k, 1)
    pl, pr = native._same_pad(w, k, 1)
    with torch.backends.cudnn.flags(allow_tf32=False):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            want = F.conv2d(F.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb)), wr.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
            # backward oracle: dY rounded to bf16 too (both gradient GEMMs read the rounded dY)
            want.backward(r(dy))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
    assert y.shape == want.shape
    assert _rel(y, want) < 2e-3, _rel(y, want)
    assert _rel(wt.grad, wr.grad) < 2e-3, _rel(wt.grad, wr.grad)
    # dX multiplies the bf16-rounded FLIPPED filter: same rounding as the oracle's filter
    assert _rel(x.grad, xr.grad) < 2e-3, _rel(x.grad, xr.grad)
    # element-wise too (a mis-addressed tap would hide in a norm only if it were tiny): borders included
    torch.testing.assert_close(y, want, rtol=2e-2, atol=2e-2 * float(want.abs().max()))
    torch.testing.assert_close(x.grad, xr.grad, rtol=2e-2, atol=2e-2 * float(xr.grad.abs().max()))
    torch.testing.assert_close(wt.grad, wr.grad, rtol=2e-2, atol=2e-2 * float(wr.grad.abs().max()))


def test_shapes_outside_the_implicit_path_fall_back_to_the_patch_matrix(monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib, native
    assert not cuda_lib.implicit_conv_ok(64, 32, 32, 3) and not cuda_lib.implicit_conv_ok(64, 32, 32, 64, (2, 2))
    assert not cuda_lib.implicit_conv_ok(2, 4, 4, 512) and cuda_lib.implicit_conv_ok(8, 4, 4, 512)
    calls = []
    real = cuda_lib.conv_igemm
    monkeypatch.setattr(cuda_lib, "conv_igemm", lambda *a, **kw: (calls.append(1), real(*a, **kw))[1])
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 16, 16, 64, generator=g).cuda()
    wt = (torch.randn(3, 3, 64, 128, generator=g) * 0.05).cuda()
    y = native.conv2d_nhwc(x, wt, (1, 2, 2, 1), "SAME")              # strided: gather + GEMM
    assert not calls and y.shape == (4, 8, 8, 128)
    r = lambda t: t.bfloat16().float()
    want = F.conv2d(F.pad(r(x).permute(0, 3, 1, 2), (0, 1, 0, 1)), r(wt).permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    assert _rel(y, want) < 5e-3
