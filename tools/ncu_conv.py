"""Driver for `ncu -k regex:gemm_bf16_tcgen05_kernel`: the implicit-GEMM convolution products of a ResNet-18 stage-0 layer
(64 images of 32x32x64, 3x3, SAME): fprop, wgrad (split-K over the pixels), dgrad -- and their wall-clock (CUDA events, warm)
next to the gather + GEMM pair they replace."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.ops import cuda_lib  # noqa: E402

torch.cuda.set_device(0)
out = {}
for (n, h, w, c, co) in [(64, 32, 32, 64, 64), (64, 16, 16, 128, 128), (64, 8, 8, 256, 256), (64, 4, 4, 512, 512)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, w, c, generator=g).cuda()
    wt = (torch.randn(3, 3, c, co, generator=g) * 0.05).cuda()
    dy = torch.randn(n * h * w, co, generator=g).cuda()
    x16 = x.bfloat16().contiguous()
    dy16 = dy.bfloat16().contiguous()
    w2 = wt.reshape(9 * c, co)

    def timed(fn, iters=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    def explicit():
        cols, _ = cuda_lib.im2col_nhwc(x, 3, 3, (1, 1), (1, 1, 1, 1))
        return cuda_lib.gemm(cols, w2, False, False, precision="bf16"), cols
    _, cols = explicit()
    flops = 2.0 * n * h * w * 9 * c * co
    r = {"fprop_implicit_us": timed(lambda: cuda_lib.conv_igemm(x16, w2, 3, 3, 1, 1)),
         "fprop_gather_plus_gemm_us": timed(lambda: explicit()),
         "wgrad_implicit_us": timed(lambda: cuda_lib.conv_igemm(x16, dy16, 3, 3, 1, 1, wgrad=True)),
         "wgrad_gemm_on_patch_matrix_us": timed(lambda: cuda_lib.gemm(cols, dy16, True, False, precision="bf16")),
         "dgrad_implicit_us": timed(lambda: cuda_lib.conv_igemm(dy16.view(n, h, w, co), w2[: 9 * co].contiguous() if c == co else w2, 3, 3, 1, 1))}
    r["fprop_implicit_tflops"] = flops / r["fprop_implicit_us"] / 1e6
    r["wgrad_implicit_tflops"] = flops / r["wgrad_implicit_us"] / 1e6
    out["%dx%dx%dx%d->%d" % (n, h, w, c, co)] = {k: round(v, 2) for k, v in r.items()}
print("CONV_PERF " + json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_perf.json"), "w"), indent=1)
