"""One launch of the persistent CTA-pair GEMM (4096^3, bf16 -> fp32) for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:persistent --launch-skip 2 -c 1 \
        -o gpurun_out/prof_gemm_pair python tools/ncu_gemm.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.ops import cuda_lib  # noqa: E402

torch.cuda.set_device(0)
M = N = K = 4096
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
c = torch.empty(M, N, device="cuda")
for _ in range(3):
    cuda_lib.gemm_raw(a, K, b, K, c, N, M, N, K, a_mn=False, b_mn=False, block_n=256, persistent=2)
torch.cuda.synchronize()
print("done")
