"""tcgen05 GEMM throughput: single-tile-per-CTA kernel vs persistent double-buffered kernel vs cuBLAS (bf16 -> fp32).

CUDA events on the launching stream, 3 warm-ups, 10 timed launches per point; operands + output of the large
shapes (>= 96 MB at 4096^3 with fp32 C) are comparable to the 126 MB L2 and are rotated across 3 buffer sets so
consecutive launches do not re-read the same lines.  Writes gpurun_out/gemm_perf.json.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.ops import cuda_lib  # noqa: E402

SHAPES = [(4096, 4096, 4096), (8192, 8192, 4096), (2048, 2048, 2048), (131072, 64, 576), (32768, 128, 1152),
          (8192, 256, 2304), (2048, 512, 4608)]       # square roofline points + ResNet-18 conv-as-GEMM shapes (batch 128, 32x32)


def time_it(fn, sets, iters=10):
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
    torch.cuda.set_device(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out = {"peaks": peaks, "points": []}
    for (M, N, K) in SHAPES:
        sets = []
        for _ in range(3):
            a = torch.randn(M, K, device="cuda").bfloat16()
            b = torch.randn(N, K, device="cuda").bfloat16()
            c = torch.empty(M, N, device="cuda")
            sets.append((a, b, c))
        flops = 2.0 * M * N * K
        row = {"M": M, "N": N, "K": K}
        bns = [bn for bn in (64, 128, 256) if bn <= max(64, N)]
        for mode, pers in (("tile", -1), ("persistent", 1), ("pair", 2)):
            best = None
            for bn in bns:
                def run(a, b, c, bn=bn, pers=pers):
                    cuda_lib.gemm_raw(a, K, b, K, c, N, M, N, K, a_mn=False, b_mn=False, block_n=bn, persistent=pers)
                ms = time_it(run, sets)
                row["%s_bn%d_tflops" % (mode, bn)] = round(flops / ms / 1e9, 1)
                best = max(best or 0.0, flops / ms / 1e9)
            row[mode + "_best_tflops"] = round(best, 1)
        a, b, c = sets[0]
        ref = a.float() @ b.float().t() if M * N <= 4096 * 4096 else None
        if ref is not None:
            cuda_lib.gemm_raw(a, K, b, K, c, N, M, N, K, a_mn=False, b_mn=False, persistent=2)
            row["pair_rel_err"] = float((c - ref).abs().max() / ref.abs().max())
        ms = time_it(lambda a, b, c: torch.matmul(a, b.t()), sets)
        row["cublas_bf16_out_tflops"] = round(flops / ms / 1e9, 1)
        out["points"].append(row)
        print(json.dumps(row), flush=True)
        del sets
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_perf.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
