"""ResNet-18 gradient check on one GPU: the kernel path with the element-wise PyTorch glue (unfused), the kernel path with
the fused NN kernels of csrc/nn_kernels.cu, and a pure-PyTorch fp32 reference (DTF_FORCE_EAGER formulation: F.conv2d,
torch ops) -- per-variable norm-wise relative error of both kernel paths against the fp32 reference, so a real defect
(one layer far off) can be told from bf16 rounding noise (all layers off by a similar, small amount).

    python tools/resnet_grad_check.py [batch]      -> gpurun_out/resnet_grad_check.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from distributed_tensorflow_b200.models.resnet import resnet18_init, resnet18_loss  # noqa: E402
from distributed_tensorflow_b200.ops import cuda_lib, native  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    init = {k: v.cuda().requires_grad_() for k, v in resnet18_init(seed=2).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 32, 32, 3, generator=g).cuda()
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (B,), generator=g), 10).float().cuda()

    def run(mode):
        native._FORCE_EAGER = mode == "eager_fp32"
        native._FUSED_BN = mode == "fused"
        cuda_lib.FUSED_NN = mode == "fused"
        old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        loss = resnet18_loss(init, x, y)
        grads = torch.autograd.grad(loss, list(init.values()))
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        native._FORCE_EAGER = False
        return float(loss), [t.detach().double() for t in grads]
    ref_l, ref = run("eager_fp32")
    out = {"batch": B, "loss": {"eager_fp32": ref_l}, "rel_err": {}}
    for mode in ("unfused", "fused"):
        l, gr = run(mode)
        out["loss"][mode] = l
        for (k, _), a, b in zip(init.items(), gr, ref):
            out["rel_err"].setdefault(k, {})[mode] = float((a - b).norm() / (b.norm() + 1e-30))
    worst = {m: max(out["rel_err"].items(), key=lambda kv: kv[1][m]) for m in ("unfused", "fused")}
    out["worst"] = {m: [k, v[m]] for m, (k, v) in worst.items()}
    convs = [k for k in out["rel_err"] if k.endswith("conv") or "conv" in k.split("/")[-1]]
    out["conv_median"] = {m: sorted(out["rel_err"][k][m] for k in convs)[len(convs) // 2] for m in ("unfused", "fused")}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "resnet_grad_check.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("RESNET_GRAD_CHECK " + json.dumps({"loss": out["loss"], "worst": out["worst"], "conv_median": out["conv_median"]}))
    for k, v in list(out["rel_err"].items())[:8]:
        print(k, v)


if __name__ == "__main__":
    main()
