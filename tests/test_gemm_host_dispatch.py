"""Host half of csrc/gemm_tcgen05.cu (dtf_gemm_bf16): argument checks, BLOCK_N / stage / kernel selection, tensor-map boxes,
grid, dynamic shared memory and cluster size -- compiled by g++ against tests/emu/gemm_host_stubs.h, which records the tensor
maps and the launch instead of encoding / issuing them.  The kernels themselves are hardware-only; this pins the dispatch
(including the round-1 bug: the K-major B box of a CTA pair must cover HALF of the tile's N rows)."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch

from distributed_tensorflow_b200.ops.cuda_lib import GemmArgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributed_tensorflow_b200", "csrc")


class Rec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("gx", ctypes.c_uint), ("gy", ctypes.c_uint), ("gz", ctypes.c_uint), ("smem", ctypes.c_longlong),
                ("cluster", ctypes.c_int), ("block_n", ctypes.c_int), ("stages", ctypes.c_int), ("num_kb", ctypes.c_int),
                ("kb_per_split", ctypes.c_int), ("atomic", ctypes.c_int), ("tiles_m", ctypes.c_int), ("tiles_n", ctypes.c_int),
                ("a_rows", ctypes.c_longlong), ("a_cols", ctypes.c_longlong), ("a_ld", ctypes.c_longlong),
                ("a_box_cols", ctypes.c_int), ("a_box_rows", ctypes.c_int),
                ("b_rows", ctypes.c_longlong), ("b_cols", ctypes.c_longlong), ("b_ld", ctypes.c_longlong),
                ("b_box_cols", ctypes.c_int), ("b_box_rows", ctypes.c_int)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libgemm_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-DDTF_HOST_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + CSRC, "-x", "c++",
                    "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(CSRC, "gemm_tcgen05.cu")], check=True)
    lib = ctypes.CDLL(so)
    lib.dtf_gemm_bf16.argtypes = [ctypes.POINTER(GemmArgs), ctypes.c_void_p]
    lib.dtf_emu_gemm_last.argtypes = [ctypes.POINTER(Rec)]
    return lib


_buf = torch.zeros(64, dtype=torch.bfloat16)          # a 16-byte aligned address; nothing is dereferenced


def _args(M, N, K, a_mn=0, b_mn=0, lda=None, ldb=None, **kw):
    g = GemmArgs()
    g.a = g.b = g.c = _buf.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.a_mn, g.b_mn = a_mn, b_mn
    g.lda = lda if lda is not None else (M if a_mn else K)
    g.ldb = ldb if ldb is not None else (N if b_mn else K)
    g.ldc, g.alpha, g.splits = N, 1.0, 1
    for k, v in kw.items():
        setattr(g, k, v)
    return g


def _run(host, g):
    rc = host.dtf_gemm_bf16(ctypes.byref(g), None)
    r = Rec()
    host.dtf_emu_gemm_last(ctypes.byref(r))
    return rc, r


def test_mnist_forward_gemm_tile_kernel(host):
    # F1 of the MNIST step: x[100, 784] . W1[784, 100] (W1 MN-major from the ps replica, pitch 104), BLOCK_N 64
    rc, r = _run(host, _args(100, 100, 784, b_mn=1, ldb=104, block_n_override=64))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy, r.gz) == (1, 2, 1) and r.cluster == 1
    assert (r.block_n, r.num_kb, r.kb_per_split, r.stages, r.atomic) == (64, 13, 13, 8, 0)
    assert r.smem == 8 * (128 * 64 * 2 + 64 * 64 * 2) + 1024
    assert (r.a_rows, r.a_cols, r.a_ld, r.a_box_cols, r.a_box_rows) == (100, 784, 784, 64, 128)       # A K-major: 128-row box
    assert (r.b_rows, r.b_cols, r.b_ld, r.b_box_cols, r.b_box_rows) == (784, 100, 104, 64, 64)        # B MN-major: [K, N], 64 x 64
    # dW1 = x^T . dh: both operands MN-major
    rc, r = _run(host, _args(784, 100, 100, a_mn=1, b_mn=1, lda=784, ldb=104, block_n_override=64))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy) == (7, 2) and (r.a_rows, r.a_cols, r.a_box_rows) == (100, 784, 64)
    assert r.num_kb == 2 and r.stages == 2


def test_large_square_gemm_uses_cta_pairs_with_half_tile_b_box(host):
    rc, r = _run(host, _args(4096, 4096, 4096))
    assert rc == 0 and r.kind == 2 and r.cluster == 2 and (r.gx, r.gy, r.gz) == (148, 1, 1)
    assert r.block_n == 256 and (r.tiles_m, r.tiles_n) == (16, 16)                                   # 256-row tiles per pair
    assert (r.b_box_cols, r.b_box_rows) == (64, 128)            # each CTA of the pair loads HALF of the 256 N rows
    assert r.stages == 6 and r.smem == 6 * (128 * 64 * 2 + 128 * 64 * 2) + 1024
    # forced 1-CTA persistent kernel: whole-tile B box, fewer and larger stages
    rc, r = _run(host, _args(4096, 4096, 4096, persistent=1))
    assert rc == 0 and r.kind == 1 and r.cluster == 1 and r.gx == 148 and r.b_box_rows == 256 and (r.tiles_m, r.tiles_n) == (32, 16)
    assert r.stages == 4 and r.smem == 4 * (128 * 64 * 2 + 256 * 64 * 2) + 1024
    # auto mode without pairs; never persistent
    rc, r = _run(host, _args(4096, 4096, 4096, cta_pair=-1))
    assert rc == 0 and r.kind == 1 and r.b_box_rows == 256
    rc, r = _run(host, _args(4096, 4096, 4096, persistent=-1))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy, r.gz) == (32, 16, 1) and r.stages == 4
    # MN-major B: pairs need BLOCK_N % 128 == 0 and keep the 64 x 64 box
    rc, r = _run(host, _args(4096, 4096, 4096, b_mn=1))
    assert rc == 0 and r.kind == 2 and (r.b_box_cols, r.b_box_rows) == (64, 64)
    # a small grid stays on the tile kernel; forcing pairs on a single 128-row block falls back to one CTA per tile
    rc, r = _run(host, _args(1024, 1024, 512))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy) == (8, 4)
    rc, r = _run(host, _args(128, 1024, 512, persistent=2))
    assert rc == 0 and r.kind == 1


def test_split_k_fused_signals_and_argument_errors(host):
    rc, r = _run(host, _args(100, 100, 784, b_mn=1, ldb=104, block_n_override=64, splits=4))
    assert rc == 0 and r.kind == 0 and r.gz == 4 and (r.kb_per_split, r.atomic, r.stages) == (4, 1, 4)
    # a fused wait (token acquire) keeps the tile kernel whatever the tile count
    flag = torch.zeros(1, dtype=torch.int64)
    rc, r = _run(host, _args(8192, 8192, 512, wait_flag=flag.data_ptr()))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy) == (64, 32)
    assert _run(host, _args(0, 8, 8))[0] == -2
    assert _run(host, _args(64, 64, 100, lda=100, ldb=100))[0] == -3                 # pitch not a multiple of 8 elements
    g = _args(64, 64, 64)
    g.a = _buf.data_ptr() + 2
    assert _run(host, g)[0] == -4                                                      # 16-byte alignment
    assert _run(host, _args(64, 64, 256, splits=2, relu=1))[0] == -5
    assert _run(host, _args(64, 96, 64, b_mn=1, ldb=96, block_n_override=48))[0] == -6  # MN-major B needs BLOCK_N % 64 == 0


def _conv(mode, n, h, w, c, cout, k=3, splits=1):
    pixels, kdim = n * h * w, k * k * c
    M, K = (pixels, kdim) if mode == 1 else (kdim, pixels)
    return _args(M, cout, K, a_mn=int(mode == 2), b_mn=1, lda=c, ldb=cout, splits=splits, conv=mode, cv_n=n, cv_h=h, cv_w=w,
                 cv_c=c, cv_kh=k, cv_kw=k, cv_pt=k // 2, cv_pl=k // 2)


def test_implicit_conv_dispatch_boxes_stages_and_grid(host):
    # fprop of a ResNet stage-0 convolution: 128-pixel boxes (4 rows of one 32-wide image), one K block per tap.  512 tiles of
    # 9 K blocks: the PERSISTENT 1-CTA kernel (never CTA pairs), one CTA per SM, 8 x 24 KB stages
    rc, r = _run(host, _conv(1, 64, 32, 32, 64, 64))
    assert rc == 0 and r.kind == 1 and (r.gx, r.gy, r.gz) == (148, 1, 1) and (r.block_n, r.num_kb, r.stages) == (64, 9, 8)
    assert (r.tiles_m, r.tiles_n) == (512, 1) and r.smem == 8 * (16384 + 8192) + 1024
    assert (r.a_rows, r.a_cols, r.a_box_cols, r.a_box_rows) == (65536, 64, 64, 128)
    assert (r.b_rows, r.b_cols, r.b_box_cols, r.b_box_rows) == (576, 64, 64, 64)
    # the same product split in two (tile kernel, 1024 CTAs): 4 x 24 KB stages so two CTAs share an SM
    rc, r = _run(host, _conv(1, 64, 32, 32, 64, 64, splits=2))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy, r.gz) == (512, 1, 2) and r.stages == 4 and r.smem == 4 * (16384 + 8192) + 1024
    # one wave or less (128 tiles): the deep pipeline stays (TMA-latency bound)
    rc, r = _run(host, _conv(1, 64, 16, 16, 128, 128))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy, r.gz) == (128, 1, 1) and (r.block_n, r.num_kb, r.stages) == (128, 18, 6)
    # 4x4 feature maps: a box spans 8 whole images; 512 channels = 8 chunks per tap; split-K over the taps x chunks
    rc, r = _run(host, _conv(1, 64, 4, 4, 512, 512, splits=9))
    assert rc == 0 and (r.gx, r.gy, r.gz) == (8, 2, 9) and (r.num_kb, r.kb_per_split, r.atomic) == (72, 8, 1) and r.a_box_rows == 128
    # wgrad: M = taps x channels (the 10th 64-channel chunk of the last tile is an all-out-of-bounds box), K = pixels in
    # 64-pixel boxes, split over the CTAs
    rc, r = _run(host, _conv(2, 64, 32, 32, 64, 64, splits=29))
    assert rc == 0 and r.kind == 0 and (r.gx, r.gy) == (5, 1) and r.num_kb == 1024 and r.a_box_rows == 64 and r.atomic == 1
    assert r.gz * r.kb_per_split >= 1024 and (r.gz - 1) * r.kb_per_split < 1024
    # rejected: 3 input channels, widths that do not tile, pixel counts that leave a partial box, fp32 operands
    assert _run(host, _conv(1, 64, 32, 32, 3, 64))[0] == -3           # (a 6-byte pixel pitch is not even a legal TMA stride)
    assert _run(host, _conv(1, 64, 32, 32, 32, 64))[0] == -8
    assert _run(host, _conv(1, 4, 24, 24, 64, 64))[0] == -8
    assert _run(host, _conv(1, 2, 4, 4, 512, 512))[0] == -8
    g = _conv(1, 64, 32, 32, 64, 64)
    g.tf32 = 1
    g.lda = g.ldb = 64
    assert _run(host, g)[0] == -8
