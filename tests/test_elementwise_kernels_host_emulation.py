"""csrc/elementwise.cu under the host emulation (tests/emu/host_emu.h; warp shuffles = per-warp exchange + barrier): the
non-tensor-core kernels of the graph tier -- fused softmax-cross-entropy fwd+bwd with the reference's clip semantics (K2/K3),
optimizer applies (K5/K6), argmax (K8), column sums / ReLU backward (K4 bias gradients), tower mean (K12), conversions,
scalar im2col / col2im (the 3-channel stem) and the CUDA-core reference GEMM -- against plain PyTorch, from the same source
and launchers the GPU tier checks in tests/test_gpu_kernels.py."""
import ctypes
import math
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributed_tensorflow_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libew_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-DDTF_HOST_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + CSRC,
                    "-x", "c++", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(CSRC, "elementwise.cu")], check=True)
    lib = ctypes.CDLL(so)
    vp, ll, i, f = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
    lib.dtf_convert_f32_bf16.argtypes = [vp, ll, vp, ll, ll, ll, ll, vp]
    lib.dtf_convert_u8_bf16.argtypes = [vp, vp, ll, f, vp]
    lib.dtf_softmax_xent.argtypes = [vp, ll, vp, ll, i, i, f, vp, vp, vp, ll, vp, ll, i, vp, ll, f, vp]
    lib.dtf_relu_grad.argtypes = [vp, vp, vp, ll, vp]
    lib.dtf_colsum.argtypes = [vp, ll, i, i, vp, vp]
    lib.dtf_argmax_rows.argtypes = [vp, ll, i, i, vp, vp]
    lib.dtf_mean_of_n.argtypes = [vp, i, vp, ll, vp]
    lib.dtf_optimizer_apply.argtypes = [vp, vp, vp, vp, vp, ll, i, f, f, i, f, f, f, f, vp]
    lib.dtf_im2col_nhwc.argtypes = [vp, vp] + [i] * 12 + [ll, vp]
    lib.dtf_col2im_nhwc.argtypes = [vp, ll, vp] + [i] * 12 + [vp]
    lib.dtf_gemm_ref.argtypes = [vp, vp, vp, i, i, i, ll, ll, ll, i, i, vp, i, f, vp]
    return lib


def _p(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("rows,cols", [(100, 10), (7, 40), (3, 1)])
def test_softmax_xent_clipped_batch_sum_and_gradient(emu, rows, cols):
    """The reference's loss: -sum(labels * log(clip(softmax, 1e-10, 1))) with the clip gating the gradient; one row is
    driven into the clip (a huge negative logit under a non-zero label)."""
    g = torch.Generator().manual_seed(rows + cols)
    logits = torch.randn(rows, cols, generator=g) * 3
    labels = torch.nn.functional.one_hot(torch.randint(0, cols, (rows,), generator=g), cols).float()
    if cols > 1:
        logits[0] = 0.0
        logits[0, 1] = -60.0                   # softmax ~ 1e-26 < 1e-10
        labels[0] = 0.0
        labels[0, 1] = 1.0
    clip = 1e-10
    loss_sum, loss_rows = torch.zeros(1), torch.zeros(rows)
    dl, probs = torch.empty(rows, cols), torch.empty(rows, cols)
    ldb = (cols + 7) // 8 * 8
    dl16 = torch.full((rows, ldb), 9.0, dtype=torch.bfloat16)
    assert emu.dtf_softmax_xent(_p(logits), cols, _p(labels), cols, rows, cols, clip, _p(loss_sum), _p(loss_rows), _p(dl), cols,
                                _p(dl16), ldb, ldb, _p(probs), cols, 1.0, None) == 0
    z = logits.double().requires_grad_()
    y = torch.softmax(z, -1)
    want = -(labels.double() * torch.log(torch.clamp(y, clip, 1.0))).sum()
    (gref,) = torch.autograd.grad(want, z)
    torch.testing.assert_close(loss_sum[0].double(), want.detach(), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(loss_rows.double().sum(), want.detach(), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(probs.double(), y.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(dl.double(), gref, rtol=1e-4, atol=1e-6)
    if cols > 1:
        assert float(dl[0].abs().max()) == 0.0                    # clipped row: no gradient at all
    assert torch.equal(dl16[:, :cols], dl.bfloat16()) and float(dl16[:, cols:].float().abs().sum()) == 0.0
    # clip_min = 0: per-row -sum(labels * log_softmax), gradient scaled by grad_scale
    soft = torch.softmax(torch.randn(rows, cols, generator=g), -1)
    assert emu.dtf_softmax_xent(_p(logits), cols, _p(soft), cols, rows, cols, 0.0, None, _p(loss_rows), _p(dl), cols, None, 0, 0,
                                None, 0, 0.5, None) == 0
    z = logits.double().requires_grad_()
    rows_want = -(soft.double() * torch.log_softmax(z, -1)).sum(-1)
    (gref,) = torch.autograd.grad(rows_want.sum(), z)
    torch.testing.assert_close(loss_rows.double(), rows_want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dl.double(), 0.5 * gref, rtol=1e-4, atol=1e-6)


def test_colsum_relu_grad_argmax_mean_of_n(emu):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 70, generator=g)
    out = torch.empty(70)
    assert emu.dtf_colsum(_p(x), 70, 37, 70, _p(out), None) == 0
    torch.testing.assert_close(out, x.sum(0), rtol=1e-5, atol=1e-5)
    y, gr = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    o = torch.empty(1000)
    assert emu.dtf_relu_grad(_p(gr), _p(y), _p(o), 1000, None) == 0
    assert torch.equal(o, gr * (y > 0))
    a = torch.randn(13, 45, generator=g)
    a[2, 7] = a[2, 30] = 99.0                                         # tie: the lower index wins (tf.argmax)
    idx = torch.empty(13, dtype=torch.int64)
    assert emu.dtf_argmax_rows(_p(a), 45, 13, 45, _p(idx), None) == 0
    assert torch.equal(idx, a.argmax(1)) and int(idx[2]) == 7
    ins = [torch.randn(515, generator=g) for _ in range(3)]
    ptrs = torch.tensor([t.data_ptr() for t in ins], dtype=torch.int64)
    m = torch.empty(515)
    assert emu.dtf_mean_of_n(_p(ptrs), 3, _p(m), 515, None) == 0
    torch.testing.assert_close(m, (ins[0] + ins[1] + ins[2]) / 3)


def test_optimizer_applies_with_bf16_shadow(emu):
    g = torch.Generator().manual_seed(1)
    n = 777
    for kind in (0, 1, 2):
        var, grad = torch.randn(n, generator=g), torch.randn(n, generator=g)
        m, v = torch.rand(n, generator=g), torch.rand(n, generator=g)
        v0, m0, var0 = v.clone(), m.clone(), var.clone()
        shadow = torch.zeros(n, dtype=torch.bfloat16)
        assert emu.dtf_optimizer_apply(_p(var), _p(m), _p(v), _p(grad), _p(shadow), n, kind, 0.05, 0.9, 1 if kind == 1 else 0,
                                       0.9, 0.999, 1e-8, 0.5, None) == 0
        ge = 0.5 * grad                                              # grad_scale
        if kind == 0:
            want = var0 - 0.05 * ge
        elif kind == 1:
            acc = 0.9 * m0 + ge
            want = var0 - (0.05 * ge + 0.05 * 0.9 * acc)             # nesterov
            torch.testing.assert_close(m, acc)
        else:
            mm, vv = 0.9 * m0 + 0.1 * ge, 0.999 * v0 + 0.001 * ge * ge
            want = var0 - 0.05 * mm / (vv.sqrt() + 1e-8)
            torch.testing.assert_close(m, mm)
            torch.testing.assert_close(v, vv)
        torch.testing.assert_close(var, want, rtol=1e-5, atol=1e-6)
        assert torch.equal(shadow, var.bfloat16())


def test_conversions_scalar_im2col_col2im_and_reference_gemm(emu):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(100, 784, generator=g)
    o = torch.zeros(100, 784, dtype=torch.bfloat16)
    assert emu.dtf_convert_f32_bf16(_p(x), 784, _p(o), 784, 100, 784, 784, None) == 0          # vector path
    assert torch.equal(o, x.bfloat16())
    xs = torch.randn(5, 13, generator=g)
    o2 = torch.full((5, 16), 3.0, dtype=torch.bfloat16)
    assert emu.dtf_convert_f32_bf16(_p(xs), 13, _p(o2), 16, 5, 13, 16, None) == 0              # scalar path, zero-padded
    assert torch.equal(o2[:, :13], xs.bfloat16()) and float(o2[:, 13:].float().abs().sum()) == 0.0
    u = torch.randint(0, 256, (1000,), generator=g, dtype=torch.uint8)
    o3 = torch.zeros(1000, dtype=torch.bfloat16)
    assert emu.dtf_convert_u8_bf16(_p(u), _p(o3), 1000, 1.0 / 255.0, None) == 0
    assert torch.equal(o3, (u.float() * (1.0 / 255.0)).bfloat16())
    # 3-channel stem convolution lowering (K = 27 padded to 32), and its adjoint
    img = torch.randn(2, 6, 5, 3, generator=g)
    n, h, w, c, k = 2, 6, 5, 3, 3
    cols = torch.full((n * h * w, 32), 5.0, dtype=torch.bfloat16)
    assert emu.dtf_im2col_nhwc(_p(img), _p(cols), n, h, w, c, k, k, 1, 1, 1, 1, h, w, 32, None) == 0
    xp = torch.nn.functional.pad(img, (0, 0, 1, 1, 1, 1))
    want = torch.cat([xp[:, ky:ky + h, kx:kx + w, :] for ky in range(k) for kx in range(k)], dim=-1).reshape(n * h * w, 27)
    assert torch.equal(cols[:, :27], want.bfloat16()) and float(cols[:, 27:].float().abs().sum()) == 0.0
    gc = torch.randn(n * h * w, 27, generator=g)
    gx = torch.empty(n, h, w, c)
    assert emu.dtf_col2im_nhwc(_p(gc), 27, _p(gx), n, h, w, c, k, k, 1, 1, 1, 1, h, w, None) == 0
    xr = img.clone().requires_grad_()
    xpr = torch.nn.functional.pad(xr, (0, 0, 1, 1, 1, 1))
    ref = torch.cat([xpr[:, ky:ky + h, kx:kx + w, :] for ky in range(k) for kx in range(k)], dim=-1).reshape(n * h * w, 27)
    (gref,) = torch.autograd.grad(ref, xr, gc)
    torch.testing.assert_close(gx, gref, rtol=1e-5, atol=1e-5)
    # reference GEMM: C = alpha * A[M,K] . B[N,K]^T + bias, ReLU
    A, B = torch.randn(9, 20, generator=g).bfloat16(), torch.randn(11, 20, generator=g).bfloat16()
    bias, C = torch.randn(11, generator=g), torch.empty(9, 11)
    assert emu.dtf_gemm_ref(_p(A), _p(B), _p(C), 9, 11, 20, 20, 20, 11, 0, 0, _p(bias), 1, 0.5, None) == 0
    torch.testing.assert_close(C, torch.relu(0.5 * (A.float() @ B.float().t()) + bias), rtol=1e-5, atol=1e-5)
