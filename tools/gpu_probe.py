"""GPU bring-up probe: runs each kernel check in its own subprocess (a trap or hang in one case does
not poison the rest) and writes gpurun_out/probe.json.  Usage: python tools/gpu_probe.py [case ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _rel_err(got, ref):
    import torch
    return float((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-6))


def _gemm_case(M, N, K, ta, tb, bias=False, relu=False, splits=1, out_bf16=False):
    import torch
    from distributed_tensorflow_b200.ops import cuda_lib
    torch.manual_seed(0)
    dev = "cuda:0"
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    bv = torch.randn(N, device=dev) if bias else None
    got = cuda_lib.gemm(a, b, ta, tb, bias=bv, relu=relu, splits=splits,
                        out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    torch.cuda.synchronize()
    a16, b16 = a.bfloat16().float(), b.bfloat16().float()
    ref = (a16.t() if ta else a16) @ (b16.t() if tb else b16)
    if bias:
        ref = ref + bv
    if relu:
        ref = torch.relu(ref)
    ref2 = cuda_lib.gemm_ref(a, b, ta, tb, bias=bv, relu=relu)
    torch.cuda.synchronize()
    return {"rel_err_vs_torch": _rel_err(got, ref), "ref_kernel_vs_torch": _rel_err(ref2, ref),
            "shape": [M, N, K], "ta": ta, "tb": tb}


@case
def gemm_nt_small():      # A K-major, B K-major ([N,K]) -- the plain DeepGEMM-style layout
    return _gemm_case(128, 64, 64, False, True)


@case
def gemm_nt_k256():
    return _gemm_case(128, 128, 256, False, True)


@case
def gemm_nn_small():      # B MN-major (TF weight layout [K,N])
    return _gemm_case(128, 64, 64, False, False)


@case
def gemm_nn_n128():
    return _gemm_case(128, 128, 128, False, False)


@case
def gemm_tn_small():      # A MN-major (x^T . dy)
    return _gemm_case(128, 64, 64, True, False)


@case
def gemm_tt_small():
    return _gemm_case(128, 64, 128, True, True)


@case
def gemm_mnist_fwd():     # [100,784].[784,100] + bias + relu
    return _gemm_case(100, 100, 784, False, False, bias=True, relu=True)


@case
def gemm_mnist_dw1():     # x^T[784,100] . dh[100,100]
    return _gemm_case(784, 100, 100, True, False)


@case
def gemm_mnist_dx():      # dy[100,10] . W^T
    return _gemm_case(100, 100, 10, False, True)


@case
def gemm_splitk():
    return _gemm_case(100, 100, 784, False, False, splits=7)


@case
def gemm_big_tiles():
    return _gemm_case(1000, 700, 520, False, False, bias=True)


@case
def gemm_bf16_out():
    return _gemm_case(256, 256, 256, False, True, out_bf16=True)


@case
def gemm_perf_4096():
    import torch
    from distributed_tensorflow_b200.ops import cuda_lib
    dev = "cuda:0"
    M = N = K = 4096
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    c = torch.empty(M, N, device=dev)
    out = {}
    for bn in (128, 256):
        for _ in range(3):
            cuda_lib.gemm_raw(a, K, b, K, c, N, M, N, K, a_mn=False, b_mn=False, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cuda_lib.gemm_raw(a, K, b, K, c, N, M, N, K, a_mn=False, b_mn=False, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out["bn%d_ms" % bn] = ms
        out["bn%d_tflops" % bn] = 2.0 * M * N * K / ms / 1e9
    ref = a.float() @ b.float().t()
    out["rel_err"] = _rel_err(c, ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.matmul(a, b.t())
    e0.record()
    for _ in range(10):
        torch.matmul(a, b.t())
    e1.record()
    torch.cuda.synchronize()
    out["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9
    return out


@case
def xent_and_apply():
    import torch
    from distributed_tensorflow_b200.ops import cuda_lib, native
    dev = "cuda:0"
    torch.manual_seed(1)
    z = torch.randn(100, 10, device=dev) * 3
    y = torch.eye(10, device=dev)[torch.randint(0, 10, (100,), device=dev)]
    loss, dl = cuda_lib.softmax_xent_fwd_bwd(z, y, 1e-10, True)
    zr = z.clone().requires_grad_(True)
    lr = -(y * torch.log(torch.clamp(torch.softmax(zr, -1), 1e-10, 1.0))).sum()
    lr.backward()
    out = {"loss_err": abs(float(loss) - float(lr)) / abs(float(lr)), "dl_err": _rel_err(dl, zr.grad)}
    w = torch.randn(1000, device=dev)
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    g = torch.randn_like(w)
    from distributed_tensorflow_b200.train.optimizer import adam_reference_step
    rw, rm, rv = adam_reference_step(w.clone(), m.clone(), v.clone(), g, 1, lr=0.01)
    lr_t = 0.01 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    cuda_lib.apply_adam_(w, m, v, g, lr_t, 0.9, 0.999, 1e-8)
    out["adam_err"] = _rel_err(w, rw)
    x = torch.randn(64, 50, device=dev)
    out["colsum_err"] = _rel_err(cuda_lib.colsum(x), x.sum(0))
    out["argmax_ok"] = bool((cuda_lib.argmax_rows(x) == x.argmax(1)).all())
    return out


@case
def graph_mode_mlp_gpu():
    """The public graph API on cuda:0: every matmul/xent/apply goes through the sm_100a kernels."""
    import numpy as np
    import torch
    import distributed_tensorflow_b200 as dtf
    from distributed_tensorflow_b200.ops import cuda_lib
    with dtf.device("/gpu:0"):
        gs = dtf.train.get_or_create_global_step()
        hid_w = dtf.Variable(dtf.truncated_normal([784, 100], stddev=1.0 / 28, seed=1), name="hid_w")
        hid_b = dtf.Variable(dtf.zeros([100]), name="hid_b")
        sm_w = dtf.Variable(dtf.truncated_normal([100, 10], stddev=0.1, seed=2), name="sm_w")
        sm_b = dtf.Variable(dtf.zeros([10]), name="sm_b")
        x = dtf.placeholder(dtf.float32, [None, 784])
        y_ = dtf.placeholder(dtf.float32, [None, 10])
        hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
        logits = dtf.nn.xw_plus_b(hid, sm_w, sm_b)
        loss = dtf.nn.clipped_softmax_xent_sum(logits, y_)
        train = dtf.train.AdamOptimizer(0.01).minimize(loss, global_step=gs)
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    xs, ys = synthetic_mnist(2000, seed=3)
    losses = []
    n0 = cuda_lib.launch_count()
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        for i in range(40):
            s = (i * 100) % 1900
            _, l = sess.run([train, loss], {x: xs[s:s + 100], y_: ys[s:s + 100]})
            losses.append(float(l))
    return {"first": losses[0], "last": losses[-1], "launches": cuda_lib.launch_count() - n0,
            "decreased": losses[-1] < 0.5 * losses[0]}


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        name = sys.argv[2]
        try:
            res = CASES[name]()
            print("PROBE_RESULT " + json.dumps({"ok": True, "res": res}))
        except Exception as e:  # noqa: BLE001
            import traceback
            print("PROBE_RESULT " + json.dumps({"ok": False, "err": repr(e), "tb": traceback.format_exc()[-1500:]}))
        return
    names = sys.argv[1:] or list(CASES)
    results = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
            if line:
                results[n] = json.loads(line[-1][len("PROBE_RESULT "):])
            else:
                results[n] = {"ok": False, "err": "no result (rc=%d)" % r.returncode,
                              "stderr": r.stderr[-1200:], "stdout": r.stdout[-600:]}
        except subprocess.TimeoutExpired:
            results[n] = {"ok": False, "err": "timeout"}
        results[n]["secs"] = round(time.time() - t0, 1)
        print(n, json.dumps(results[n])[:600], flush=True)
    with open(os.path.join(out_dir, "probe.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
