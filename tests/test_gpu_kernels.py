"""Kernel-correctness tier: each sm_100a kernel vs a plain PyTorch fp32 reference of the same op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

cuda = torch.cuda.is_available()


@pytest.fixture(scope="module")
def lib():
    if not cuda:
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib
    cuda_lib.load()          # must load on a GPU box: no silent fallback
    old = cuda_lib.set_matmul_precision("bf16")      # the bf16-rounded references below; the tf32 tests pass precision=
    yield cuda_lib
    cuda_lib.set_matmul_precision(old)


def _ref_mm(a, b, ta, tb):
    a16, b16 = a.bfloat16().float(), b.bfloat16().float()
    return (a16.t() if ta else a16) @ (b16.t() if tb else b16)


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (100, 100, 784), (784, 100, 100), (100, 10, 100), (300, 520, 200),
                                   (1, 8, 24)])
def test_tcgen05_gemm_all_operand_majors(lib, M, N, K, ta, tb):
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((K, M) if ta else (M, K), device="cuda")
    b = torch.randn((N, K) if tb else (K, N), device="cuda")
    got = lib.gemm(a, b, ta, tb)
    ref = _ref_mm(a, b, ta, tb)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (100, 100, 784), (784, 100, 100), (100, 10, 100), (300, 520, 200),
                                   (1, 8, 24), (256, 96, 36)])
def test_tcgen05_gemm_tf32_all_operand_majors_vs_pure_fp32(lib, M, N, K, ta, tb):
    """fp32 operands read in place by TMA and multiplied as TF32 (tcgen05.mma.kind::tf32), all four operand-major
    combinations, against a float64 matmul of the UNROUNDED fp32 inputs: the only error is TF32's 10-bit mantissa
    (2^-11 per operand), i.e. ~1e-3 norm-wise -- an order of magnitude closer to the fp32 reference model than bf16."""
    torch.manual_seed(M * 5 + N * 3 + K)
    a = torch.randn((K, M) if ta else (M, K), device="cuda")
    b = torch.randn((N, K) if tb else (K, N), device="cuda")
    got = lib.gemm(a, b, ta, tb, precision="tf32")
    ref = ((a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double()))
    rel = float((got.double() - ref).norm() / ref.norm())
    assert rel < 1.5e-3, rel
    bf = lib.gemm(a, b, ta, tb, precision="bf16")
    rel_bf = float((bf.double() - ref).norm() / ref.norm())
    assert rel < rel_bf, (rel, rel_bf)               # and it IS closer to fp32 than the bf16 path


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tcgen05_gemm_tf32_persistent_and_epilogues(lib, mode):
    torch.manual_seed(11)
    M, N, K = 2048, 512, 320
    a, b, bias = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda"), torch.randn(N, device="cuda")
    got = lib.gemm(a, b, bias=bias, relu=True, precision="tf32", persistent=mode, block_n=256 if mode == 2 else 0)
    ref = torch.relu(a.double() @ b.double() + bias.double())
    assert float((got.double() - ref).norm() / ref.norm()) < 1.5e-3
    got_t = lib.gemm(a.t().contiguous(), b.t().contiguous(), True, True, precision="tf32", persistent=mode)
    ref_t = a.double() @ b.double()
    assert float((got_t.double() - ref_t).norm() / ref_t.norm()) < 1.5e-3


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K,bn", [(1000, 520, 200, 0), (2048, 1536, 512, 256), (4100, 96, 72, 0), (512, 512, 2048, 128)])
def test_persistent_gemm_double_buffered_tmem(lib, M, N, K, bn, ta, tb, mode):
    """Persistent kernel (tile loop per SM, two TMEM accumulator buffers, epilogue overlapped with the next tile's
    mainloop), forced on so that small grids exercise several tiles per CTA as well.  mode 2 = CTA pairs
    (tcgen05 cta_group::2: 256-row tiles, each CTA loads half of B; falls back to 1-CTA when the tile shape
    does not split)."""
    if tb is False and bn == 0 and N % 64:
        bn = 0                                       # MN-major B picks a multiple of 64 itself
    torch.manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), device="cuda")
    b = torch.randn((N, K) if tb else (K, N), device="cuda")
    bias = torch.randn(N, device="cuda")
    got = lib.gemm(a, b, ta, tb, bias=bias, relu=True, persistent=mode, block_n=bn)
    ref = torch.relu(_ref_mm(a, b, ta, tb) + bias)
    torch.testing.assert_close(got, ref, rtol=3e-4, atol=4e-3)
    # auto-dispatch (persistent when tiles > SMs) agrees with the single-tile-per-CTA kernel
    plain = lib.gemm(a, b, ta, tb, bias=bias, relu=True, persistent=-1, block_n=bn)
    torch.testing.assert_close(got, plain, rtol=0, atol=0)


def test_gemm_fused_bias_relu_and_bf16_out(lib):
    torch.manual_seed(0)
    a, b, bias = torch.randn(100, 784, device="cuda"), torch.randn(784, 100, device="cuda"), torch.randn(100, device="cuda")
    got = lib.gemm(a, b, bias=bias, relu=True)
    ref = torch.relu(_ref_mm(a, b, False, False) + bias)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=3e-3)
    got16 = lib.gemm(a, b, bias=bias, relu=True, out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    torch.testing.assert_close(got16.float(), ref, rtol=1e-2, atol=5e-2)


def test_gemm_split_k_matches(lib):
    torch.manual_seed(1)
    a, b = torch.randn(100, 784, device="cuda"), torch.randn(784, 100, device="cuda")
    torch.testing.assert_close(lib.gemm(a, b, splits=7), _ref_mm(a, b, False, False), rtol=2e-4, atol=3e-3)


def test_gemm_mask_colsum_and_signal_epilogue(lib):
    """ReLU-backward mask, bias-gradient column sums and the arrival counter, all from the GEMM epilogue."""
    torch.manual_seed(2)
    M, N, K = 100, 100, 16
    dy = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda")              # [N,K] row-major == K-major B
    h = torch.randn(M, N, device="cuda")
    a16, lda = lib.to_bf16_padded(dy)
    b16, ldb = lib.to_bf16_padded(w)
    h16, ldm = lib.to_bf16_padded(h)
    c = torch.zeros(M, N, device="cuda")
    cs = torch.zeros(N, device="cuda")
    sig = torch.zeros(1, dtype=torch.int64, device="cuda")
    lib.gemm_raw(a16, lda, b16, ldb, c, N, M, N, K, a_mn=False, b_mn=False, mask=h16, ldmask=ldm, colsum=cs,
                 signal=sig.data_ptr())
    ref = _ref_mm(dy, w, False, True) * (h16[:, :N].float() > 0)
    torch.testing.assert_close(c, ref, rtol=2e-4, atol=2e-3)
    torch.testing.assert_close(cs, ref.sum(0), rtol=1e-3, atol=1e-2)
    assert int(sig.item()) == 1                      # one CTA -> one arrival


def test_gemm_wait_flag_already_satisfied_and_timeout_sets_err(lib):
    a, b = torch.randn(64, 64, device="cuda"), torch.randn(64, 64, device="cuda")
    a16, lda = lib.to_bf16_padded(a)
    b16, ldb = lib.to_bf16_padded(b)
    c = torch.empty(64, 64, device="cuda")
    flag = torch.tensor([5], dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    # b is [N, K] row-major == K-major B operand (b_mn=False)
    lib.gemm_raw(a16, lda, b16, ldb, c, 64, 64, 64, 64, False, False, wait_flag=flag.data_ptr(), wait_target=5,
                 err=err.data_ptr())
    torch.testing.assert_close(c, _ref_mm(a, b, False, True), rtol=2e-4, atol=2e-3)
    assert int(err.item()) == 0
    lib.gemm_raw(a16, lda, b16, ldb, c, 64, 64, 64, 64, False, False, wait_flag=flag.data_ptr(), wait_target=6,
                 err=err.data_ptr(), timeout_ns=20_000_000)
    torch.cuda.synchronize()
    assert int(err.item()) == 1                      # bounded wait, no hang


@pytest.mark.parametrize("clip", [1e-10, 0.0])
def test_softmax_xent_fwd_bwd(lib, clip):
    torch.manual_seed(3)
    z = (torch.randn(100, 10, device="cuda") * 4)
    z[0] = torch.tensor([60.0] + [0.0] * 9, device="cuda")      # forces probabilities under the clip
    y = torch.eye(10, device="cuda")[torch.randint(0, 10, (100,), device="cuda")]
    y[0] = torch.eye(10, device="cuda")[3]
    loss, dl = lib.softmax_xent_fwd_bwd(z, y, clip, reduce_sum=True)
    zr = z.clone().requires_grad_(True)
    if clip > 0:
        ref = -(y * torch.log(torch.clamp(torch.softmax(zr, -1), clip, 1.0))).sum()
    else:
        ref = -(y * torch.log_softmax(zr, -1)).sum()
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    torch.testing.assert_close(dl, zr.grad, rtol=1e-4, atol=1e-5)
    rows, _ = lib.softmax_xent_fwd_bwd(z, y, clip, reduce_sum=False)
    assert rows.shape == (100,) and abs(float(rows.sum()) - float(ref)) < 1e-3 * abs(float(ref))


def test_optimizer_apply_kernels(lib):
    from distributed_tensorflow_b200.train.optimizer import adam_reference_step
    torch.manual_seed(4)
    w0, g = torch.randn(1001, device="cuda"), torch.randn(1001, device="cuda")
    w = w0.clone()
    lib.apply_sgd_(w, g, 0.1)
    torch.testing.assert_close(w, w0 - 0.1 * g)
    w, acc = w0.clone(), torch.zeros_like(w0)
    for t in range(3):
        lib.apply_momentum_(w, acc, g, 0.1, 0.9)
    ra, rw = torch.zeros_like(w0), w0.clone()
    for t in range(3):
        ra = 0.9 * ra + g
        rw = rw - 0.1 * ra
    torch.testing.assert_close(w, rw, rtol=1e-5, atol=1e-6)
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    rw, rm, rv = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    shadow = torch.empty(1001, dtype=torch.bfloat16, device="cuda")
    for t in range(1, 4):
        lr_t = 0.01 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
        lib.apply_adam_(w, m, v, g, lr_t, 0.9, 0.999, 1e-8, shadow=shadow)
        rw, rm, rv = adam_reference_step(rw, rm, rv, g, t, lr=0.01)
    torch.testing.assert_close(w, rw, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(shadow.float(), w.bfloat16().float())


def test_small_kernels(lib):
    torch.manual_seed(5)
    x = torch.randn(77, 130, device="cuda")
    torch.testing.assert_close(lib.colsum(x), x.sum(0), rtol=1e-5, atol=1e-4)
    assert bool((lib.argmax_rows(x) == x.argmax(1)).all())
    g, y = torch.randn(50, 20, device="cuda"), torch.randn(50, 20, device="cuda")
    torch.testing.assert_close(lib.relu_grad(g, y), g * (y > 0))
    t16, ld = lib.to_bf16_padded(torch.randn(5, 13, device="cuda"))
    assert ld == 16 and t16.shape == (5, 16) and bool((t16[:, 13:] == 0).all())


def test_conv_lowering_matches_torch(lib):
    from distributed_tensorflow_b200.ops import native
    torch.manual_seed(6)
    x = torch.randn(2, 9, 9, 5, device="cuda", requires_grad=True)
    w = torch.randn(3, 3, 5, 7, device="cuda", requires_grad=True)
    y = native.conv2d_nhwc(x, w, (1, 2, 2, 1), "SAME")
    xr = x.detach().bfloat16().float().requires_grad_(True)
    wr = w.detach().bfloat16().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), stride=2, padding=1).permute(0, 2, 3, 1)
    torch.testing.assert_close(y, yr, rtol=2e-3, atol=2e-2)
    gy = torch.randn_like(y)
    y.backward(gy)
    yr.backward(gy)
    torch.testing.assert_close(w.grad, wr.grad, rtol=2e-2, atol=1e-1)
    torch.testing.assert_close(x.grad, xr.grad, rtol=2e-2, atol=1e-1)


def test_graph_api_runs_on_native_kernels_and_matches_cpu(lib):
    """The public graph API placed on /gpu:0 goes through the sm_100a kernels and tracks the CPU oracle."""
    import distributed_tensorflow_b200 as dtf

    def build(dev):
        with dtf.device(dev):
            gs = dtf.train.get_or_create_global_step()
            w1 = dtf.Variable(dtf.truncated_normal([784, 32], stddev=1 / 28, seed=1), name="hid_w")
            b1 = dtf.Variable(dtf.zeros([32]), name="hid_b")
            w2 = dtf.Variable(dtf.truncated_normal([32, 10], stddev=0.2, seed=2), name="sm_w")
            b2 = dtf.Variable(dtf.zeros([10]), name="sm_b")
            x, y_ = dtf.placeholder(dtf.float32, [None, 784]), dtf.placeholder(dtf.float32, [None, 10])
            hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, w1, b1))
            y = dtf.nn.softmax(dtf.nn.xw_plus_b(hid, w2, b2))
            loss = -dtf.reduce_sum(y_ * dtf.log(dtf.clip_by_value(y, 1e-10, 1.0)))
            train = dtf.train.GradientDescentOptimizer(0.001).minimize(loss, global_step=gs)
        return x, y_, loss, train
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    xs, ys = synthetic_mnist(300, seed=9)
    out = {}
    for dev in ("/cpu:0", "/gpu:0"):
        dtf.reset_default_graph()
        x, y_, loss, train = build(dev)
        n0 = lib.launch_count()
        with dtf.Session() as sess:
            sess.run(dtf.global_variables_initializer())
            out[dev] = [float(sess.run([train, loss], {x: xs[i * 100:(i + 1) * 100], y_: ys[i * 100:(i + 1) * 100]})[1])
                        for i in range(3)]
        if dev == "/gpu:0":
            assert lib.launch_count() - n0 >= 3 * 6          # GEMMs + conversions really ran natively
    np.testing.assert_allclose(out["/gpu:0"], out["/cpu:0"], rtol=3e-2)
