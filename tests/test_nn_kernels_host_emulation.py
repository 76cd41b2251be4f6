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
