"""csrc/nn_kernels.cu compiled by g++ against tests/emu/host_emu.h (threads + barriers emulate a thread block, blocks run
one after another) and checked against the PyTorch references: the SAME kernel source, launch configuration and
launcher code the GPU runs -- indexing, the shared-memory reduction tree, the last-block ticket logic and its reset --
without a GPU.  Memory-model races and performance are what only the hardware run (tests/test_gpu_nn_fused.py) can show."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from distributed_tensorflow_b200.ops import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "distributed_tensorflow_b200", "csrc", "nn_kernels.cu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libnn_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-DDTF_HOST_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-x", "c++",
                    "-shared", "-fPIC", "-pthread", "-o", so, SRC], check=True)
    lib = ctypes.CDLL(so)
    vp, ll, i, f = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
    lib.dtf_bn_row_splits.argtypes = [ll, i]
    lib.dtf_bn_workspace_floats.argtypes = [ll, i]
    lib.dtf_bn_workspace_floats.restype = ll
    lib.dtf_bn_reduce.argtypes = [i] + [vp] * 5 + [ll, i] + [vp] * 4 + [f, vp]
    lib.dtf_bn_apply.argtypes = [vp] * 7 + [ll, i, i, vp]
    lib.dtf_bn_bwd_apply.argtypes = [vp] * 10 + [ll, i, vp]
    return lib


def _p(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("rows,C", [(200, 132), (37, 64), (64, 8), (130, 256)])
@pytest.mark.parametrize("relu,with_res", [(False, False), (True, True)])
def test_emulated_bn_kernels_match_references(emu, rows, C, relu, with_res):
    g = torch.Generator().manual_seed(rows * 7 + C)
    x = torch.randn(rows, C, generator=g) * 1.5 + 0.3
    scale, offset = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(rows, C, generator=g) if with_res else None
    dy = torch.randn(rows, C, generator=g)
    ws = torch.full((int(emu.dtf_bn_workspace_floats(rows, C)),), float("nan"))
    tickets = torch.zeros(64, dtype=torch.int32)
    mean, rstd, y = torch.empty(C), torch.empty(C), torch.empty(rows, C)
    for _ in range(2):                       # twice: the ticket reset makes the kernel replayable
        assert emu.dtf_bn_reduce(0, _p(x), None, None, None, None, rows, C, _p(ws), _p(tickets), _p(mean), _p(rstd), 1e-5, None) == 0
        assert int(tickets.abs().sum()) == 0
    assert emu.dtf_bn_apply(_p(x), _p(res), _p(y), _p(mean), _p(rstd), _p(scale), _p(offset), rows, C, int(relu), None) == 0
    xd = x.double()
    m, r = xd.mean(0), torch.rsqrt(xd.var(0, unbiased=False) + 1e-5)
    torch.testing.assert_close(mean.double(), m, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rstd.double(), r, rtol=1e-5, atol=1e-6)
    want = native.bn_train_reference(x, scale, offset, res, relu)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)

    doffset, dscale, dx = torch.empty(C), torch.empty(C), torch.empty(rows, C)
    dres = torch.empty(rows, C) if with_res else None
    mask = y if relu else None
    assert emu.dtf_bn_reduce(1, _p(x), _p(dy), _p(mask), _p(mean), _p(rstd), rows, C, _p(ws), _p(tickets), _p(doffset), _p(dscale),
                             0.0, None) == 0
    assert emu.dtf_bn_bwd_apply(_p(dy), _p(mask), _p(x), _p(mean), _p(rstd), _p(scale), _p(doffset), _p(dscale), _p(dx), _p(dres),
                                rows, C, None) == 0
    wdx, wds, wdo, wdr = native.bn_backward_reference(dy.double(), want.double(), xd, m, r, scale.double(), relu)
    torch.testing.assert_close(doffset.double(), wdo, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dscale.double(), wds, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dx.double(), wdx, rtol=1e-4, atol=1e-4)
    if with_res:
        assert torch.equal(dres, (dy * (want > 0)) if relu else dy)


def test_emulated_launchers_reject_bad_arguments(emu):
    x = torch.zeros(8, 6)
    assert emu.dtf_bn_reduce(0, _p(x), None, None, None, None, 8, 6, None, None, None, None, 1e-5, None) == -1     # C % 4
    x = torch.zeros(9, 8)
    off = x.view(-1)[1:65].view(8, 8)                                                                       # 4-byte offset
    assert emu.dtf_bn_apply(_p(off), None, _p(x), None, None, None, None, 8, 8, 0, None) == -1


def _im2col_ref(x, kh, kw, sh, sw, pt, pb, pl, pr):
    n, h, w, c = x.shape
    xp = torch.nn.functional.pad(x, (0, 0, pl, pr, pt, pb))
    ho, wo = (h + pt + pb - kh) // sh + 1, (w + pl + pr - kw) // sw + 1
    cols = torch.empty(n, ho, wo, kh * kw * c)
    for ky in range(kh):
        for kx in range(kw):
            cols[..., (ky * kw + kx) * c:(ky * kw + kx + 1) * c] = xp[:, ky:ky + (ho - 1) * sh + 1:sh, kx:kx + (wo - 1) * sw + 1:sw, :]
    return cols.reshape(n * ho * wo, kh * kw * c), ho, wo


@pytest.mark.parametrize("shape,k,stride,pads", [((2, 6, 6, 8), 3, 1, (1, 1, 1, 1)), ((1, 7, 5, 16), 3, 2, (1, 1, 1, 1)),
                                                 ((2, 4, 4, 8), 1, 2, (0, 0, 0, 0)), ((1, 8, 8, 8), 3, 2, (0, 1, 0, 1))])
def test_emulated_vector_im2col_col2im(emu, shape, k, stride, pads):
    vp, ll, i = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emu.dtf_im2col_nhwc_vec8.argtypes = [vp, vp] + [i] * 12 + [ll, vp]
    emu.dtf_col2im_nhwc_vec4.argtypes = [vp, ll, vp] + [i] * 12 + [vp]
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = torch.randn(*shape, generator=g)
    n, h, w, c = shape
    pt, pb, pl, pr = pads
    want, ho, wo = _im2col_ref(x, k, k, stride, stride, pt, pb, pl, pr)
    ld = k * k * c
    cols = torch.full((n * ho * wo, ld), 7.0, dtype=torch.bfloat16)
    assert emu.dtf_im2col_nhwc_vec8(_p(x), _p(cols), n, h, w, c, k, k, stride, stride, pt, pl, ho, wo, ld, None) == 0
    assert torch.equal(cols, want.bfloat16())              # same round-to-nearest-even as the hardware conversion
    # col2im is the adjoint of im2col: <im2col(x), G> == <x, col2im(G)>; and equals autograd's gradient
    gcols = torch.randn(n * ho * wo, ld, generator=g)
    gx = torch.full(shape, float("nan"))
    assert emu.dtf_col2im_nhwc_vec4(_p(gcols), ld, _p(gx), n, h, w, c, k, k, stride, stride, pt, pl, ho, wo, None) == 0
    xr = x.clone().requires_grad_()
    ref, _, _ = _im2col_ref(xr, k, k, stride, stride, pt, pb, pl, pr)
    (gref,) = torch.autograd.grad(ref, xr, gcols)
    torch.testing.assert_close(gx, gref, rtol=1e-5, atol=1e-5)
    # ineligible shapes are refused (the caller falls back to the scalar kernels)
    x3 = torch.zeros(1, 4, 4, 3)
    assert emu.dtf_im2col_nhwc_vec8(_p(x3), _p(cols), 1, 4, 4, 3, 3, 3, 1, 1, 1, 1, 4, 4, 32, None) == -1


@pytest.mark.parametrize("shape,k,stride,pads", [((2, 8, 8, 8), 3, 2, (0, 1, 0, 1)), ((1, 7, 5, 4), 2, 2, (0, 1, 0, 1)),
                                                 ((2, 6, 6, 12), 3, 1, (1, 1, 1, 1))])
def test_emulated_pooling_kernels(emu, shape, k, stride, pads):
    vp, i = ctypes.c_void_p, ctypes.c_int
    emu.dtf_maxpool_nhwc_fwd.argtypes = [vp, vp, vp] + [i] * 12 + [vp]
    emu.dtf_maxpool_nhwc_bwd.argtypes = [vp, vp, vp] + [i] * 12 + [vp]
    emu.dtf_global_avgpool_nhwc.argtypes = [vp, vp, i, i, i, i, vp]
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    n, h, w, c = shape
    pt, pb, pl, pr = pads
    ho, wo = (h + pt + pb - k) // stride + 1, (w + pl + pr - k) // stride + 1
    y = torch.empty(n, ho, wo, c)
    arg = torch.empty(n, ho, wo, c, dtype=torch.uint8)
    assert emu.dtf_maxpool_nhwc_fwd(_p(x), _p(y), _p(arg), n, h, w, c, k, k, stride, stride, pt, pl, ho, wo, None) == 0
    xr = x.clone().requires_grad_()
    xp = torch.nn.functional.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float("-inf"))
    ref = torch.nn.functional.max_pool2d(xp, k, stride).permute(0, 2, 3, 1)
    assert torch.equal(y, ref.detach())
    dy = torch.randn(n, ho, wo, c, generator=g)
    dx = torch.full(shape, float("nan"))
    assert emu.dtf_maxpool_nhwc_bwd(_p(dy), _p(arg), _p(dx), n, h, w, c, k, k, stride, stride, pt, pl, ho, wo, None) == 0
    (gref,) = torch.autograd.grad(ref, xr, dy)
    torch.testing.assert_close(dx, gref, rtol=1e-6, atol=1e-6)
    # global average pooling and its gradient
    out = torch.empty(n, c)
    assert emu.dtf_global_avgpool_nhwc(_p(x), _p(out), n, h * w, c, 0, None) == 0
    torch.testing.assert_close(out, x.mean(dim=(1, 2)), rtol=1e-5, atol=1e-6)
    gd = torch.randn(n, c, generator=g)
    gx = torch.empty(shape)
    assert emu.dtf_global_avgpool_nhwc(_p(gd), _p(gx), n, h * w, c, 1, None) == 0
    torch.testing.assert_close(gx, (gd / (h * w))[:, None, None, :].expand(shape), rtol=1e-6, atol=1e-7)
    assert emu.dtf_maxpool_nhwc_fwd(_p(x), _p(y), _p(arg), n, h, w, 6, k, k, stride, stride, pt, pl, ho, wo, None) == -1   # C % 4


def test_kernels_do_not_depend_on_block_order():
    """The emulator runs blocks one after another; ``DTF_EMU_BLOCK_ORDER=reverse`` walks the grid backwards.  Everything
    except ps_apply (where block 0 deciding first is the design, and all CTAs are co-resident on the hardware) must give the
    same answers: the nn / element-wise suites and the head / push / fabric / staging tests are re-run in reverse order."""
    import sys
    env = dict(os.environ, DTF_EMU_BLOCK_ORDER="reverse")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(here, "test_nn_kernels_host_emulation.py"),
                        os.path.join(here, "test_elementwise_kernels_host_emulation.py"),
                        os.path.join(here, "test_ps_kernels_host_emulation.py"),
                        "-k", "(emulated or softmax or colsum or optimizer or conversions or head or push_grad or fabric or stage) and not block_order"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
