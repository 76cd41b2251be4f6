"""Bandwidth of the fused NN kernels (csrc/nn_kernels.cu) vs the element-wise PyTorch formulation they replace, on the
ResNet-18 (CIFAR, batch 64) activation shapes: fused training BN (+ residual + ReLU) forward / backward, channel-vectorised
im2col / col2im vs the scalar kernels, NHWC max pooling, global average pooling.

CUDA events on the launching stream, 3 warm-ups, 10 timed calls per point, three rotating buffer sets (the stage-0 tensors
are 16.8 MB each, well below the 126 MB L2: rotation keeps consecutive calls from re-reading the same lines).  GB/s counts
the bytes the FUSED kernel must move (reads + writes), so the eager number is "effective".  Writes gpurun_out/nn_perf.json.
    python tools/nn_perf.py            # one GPU
    python tools/nn_perf.py --ncu      # a single pass over the fused kernels only (run under ncu --set full -k regex:bn_|im2col|pool)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.ops import cuda_lib, native  # noqa: E402

DEV = os.environ.get("DTF_NN_PERF_DEVICE", "cuda")      # "cpu": dry run of this script on tiny shapes under the kernel emulation
ACTS = [(64 * 32 * 32, 64), (64 * 16 * 16, 128), (64 * 8 * 8, 256), (64 * 4 * 4, 512)]       # [rows, C] per stage
CONVS = [((64, 32, 32, 64), 3, 1), ((64, 16, 16, 128), 3, 1), ((64, 8, 8, 256), 3, 1), ((64, 4, 4, 512), 3, 1)]


def time_it(fn, sets, iters=10):
    if DEV == "cpu":
        import time
        t0 = time.time()
        fn(*sets[0])
        return (time.time() - t0) * 1e3
    for i in range(3):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    global ACTS, CONVS
    ncu = "--ncu" in sys.argv
    if DEV == "cpu":
        cuda_lib.enable_emulation()
        ACTS, CONVS = [(96, 8), (40, 16)], [((2, 6, 6, 8), 3, 1)]
    else:
        torch.cuda.set_device(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out = {"peaks": peaks, "bn": [], "lowering": [], "pool": []}
    for rows, C in ACTS:
        sets = [(torch.randn(rows, C, device=DEV), torch.randn(rows, C, device=DEV), torch.randn(rows, C, device=DEV))
                for _ in range(3)]
        scale, offset = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
        saved = {}

        def fused_fwd(x, res, dy):
            saved["y"], saved["mean"], saved["rstd"] = cuda_lib.bn_forward(x, scale, offset, res, True, 1e-5)

        def fused_bwd(x, res, dy):
            cuda_lib.bn_backward(dy, saved["y"], x, saved["mean"], saved["rstd"], scale, True)

        def eager_fwd(x, res, dy):
            native.bn_train_reference(x, scale, offset, res, True)

        def eager_fwd_bwd(x, res, dy):
            xl, rl = x.detach().requires_grad_(), res.detach().requires_grad_()
            sl, ol = scale.detach().requires_grad_(), offset.detach().requires_grad_()
            torch.autograd.grad(native.bn_train_reference(xl, sl, ol, rl, True), [xl, sl, ol, rl], dy)
        if ncu:
            fused_fwd(*sets[0])
            fused_bwd(*sets[0])
            continue
        t = {"fused_fwd_ms": time_it(fused_fwd, sets), "eager_fwd_ms": time_it(eager_fwd, sets)}
        fused_fwd(*sets[0])
        t["fused_bwd_ms"] = time_it(fused_bwd, [sets[0]])
        t["eager_fwd_bwd_ms"] = time_it(eager_fwd_bwd, sets)
        nbytes = rows * C * 4
        row = {"rows": rows, "C": C, **{k: round(v, 4) for k, v in t.items()},
               "fused_fwd_gbps": round(4 * nbytes / t["fused_fwd_ms"] / 1e6, 1),      # stats: read x; apply: read x, res, write y
               "fused_bwd_gbps": round(8 * nbytes / t["fused_bwd_ms"] / 1e6, 1)}      # sums: dy, x, y; apply: dy, y, x -> dx, dres
        out["bn"].append(row)
        print(row, flush=True)
    for shape, k, stride in CONVS:
        sets = [(torch.randn(*shape, device=DEV),) for _ in range(3)]
        res = {}
        for fused in (False, True):
            cuda_lib.FUSED_NN = fused
            cols, _ = cuda_lib.im2col_nhwc(sets[0][0], k, k, (stride, stride), (1, 1, 1, 1))
            gcols = torch.randn(cols.shape, device=DEV)
            if ncu and not fused:
                continue
            if ncu:
                cuda_lib.col2im_nhwc(gcols, shape, k, k, (stride, stride), (1, 1, 1, 1))
                continue
            tag = "vec" if fused else "scalar"
            res["im2col_%s_ms" % tag] = round(time_it(lambda x: cuda_lib.im2col_nhwc(x, k, k, (stride, stride), (1, 1, 1, 1)), sets), 4)
            res["col2im_%s_ms" % tag] = round(time_it(lambda x: cuda_lib.col2im_nhwc(gcols, shape, k, k, (stride, stride), (1, 1, 1, 1)), sets), 4)
        if not ncu:
            n, h, w, c = shape
            res.update({"shape": shape, "k": k, "cols_mb": round(n * h * w * k * k * c * 2 / 1e6, 1)})
            res["im2col_vec_gbps"] = round((n * h * w * c * 4 + n * h * w * k * k * c * 2) / res["im2col_vec_ms"] / 1e6, 1)
            out["lowering"].append(res)
            print(res, flush=True)
    cuda_lib.FUSED_NN = True
    for shape in ([(64, 32, 32, 64), (64, 16, 16, 128)] if DEV != "cpu" else [(2, 6, 6, 8)]):
        sets = [(torch.randn(*shape, device=DEV),) for _ in range(3)]
        if ncu:
            native.max_pool_nhwc(sets[0][0], (1, 3, 3, 1), (1, 2, 2, 1), "SAME")
            native.global_avg_pool(sets[0][0])
            continue
        row = {"shape": shape}
        for fused in (True, False):
            cuda_lib.FUSED_NN = fused
            tag = "ours" if fused else "eager"
            row["maxpool_%s_ms" % tag] = round(time_it(lambda x: native.max_pool_nhwc(x, (1, 3, 3, 1), (1, 2, 2, 1), "SAME"), sets), 4)
            row["gap_%s_ms" % tag] = round(time_it(lambda x: native.global_avg_pool(x), sets), 4)
        out["pool"].append(row)
        print(row, flush=True)
    if not ncu:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nn_perf.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
