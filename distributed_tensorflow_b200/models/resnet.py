"""ResNet-18 under the parameter-server API (BASELINE.json config 5).

NHWC activations, HWIO filters (channels-last is the tensor-core layout).  Every convolution and the final
dense layer lower to the tcgen05 GEMM through ``ops.native`` (im2col gather kernel + ``gemm_bf16_tcgen05``;
backward = two more GEMMs + col2im), the loss is the fused softmax-cross-entropy kernel; batch-norm
(training mode, batch statistics) + residual add + ReLU go through ``native.batch_norm_train`` (fused kernels
of ``csrc/nn_kernels.cu`` behind ``DTF_FUSED_BN=1``, element-wise PyTorch glue otherwise); pooling is glue.  On CPU the same functions run
plain PyTorch, which is the oracle in the tests.

``resnet18_param_shapes(num_classes, stem)`` lists the variables in creation order (what the ps shards
round-robin); ``resnet18_loss(params, x, y)`` is the ``loss_fn`` for :class:`parallel.generic_engine.GenericPSEngine`;
``build_resnet18_graph`` builds the same network with the graph API (``dtf.nn.conv2d`` ...) for
``MonitoredTrainingSession`` use.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from ..ops import native

__all__ = ["resnet18_param_shapes", "resnet18_init", "resnet18_forward", "resnet18_loss", "build_resnet18_graph"]

_STAGES = [(64, 1), (128, 2), (256, 2), (512, 2)]      # (channels, stride of the first block); 2 basic blocks each


def resnet18_param_shapes(num_classes: int = 10, stem: str = "cifar", in_ch: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    k = 3 if stem == "cifar" else 7
    shapes: List[Tuple[str, Tuple[int, ...]]] = [("stem/conv", (k, k, in_ch, 64)), ("stem/bn_scale", (64,)), ("stem/bn_offset", (64,))]
    cin = 64
    for si, (c, stride) in enumerate(_STAGES):
        for bi in range(2):
            p = "stage%d/block%d" % (si, bi)
            shapes += [(p + "/conv1", (3, 3, cin, c)), (p + "/bn1_scale", (c,)), (p + "/bn1_offset", (c,)),
                       (p + "/conv2", (3, 3, c, c)), (p + "/bn2_scale", (c,)), (p + "/bn2_offset", (c,))]
            if bi == 0 and (stride != 1 or cin != c):
                shapes += [(p + "/down_conv", (1, 1, cin, c)), (p + "/down_bn_scale", (c,)), (p + "/down_bn_offset", (c,))]
            cin = c
    shapes += [("fc/w", (512, num_classes)), ("fc/b", (num_classes,))]
    return shapes


def resnet18_init(num_classes: int = 10, stem: str = "cifar", seed: int = 0, in_ch: int = 3) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in resnet18_param_shapes(num_classes, stem, in_ch):
        if name.endswith("_scale"):
            out[name] = torch.ones(shape)
        elif name.endswith("_offset") or name.endswith("/b"):
            out[name] = torch.zeros(shape)
        elif len(shape) == 4:
            fan_in = shape[0] * shape[1] * shape[2]
            out[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        else:
            out[name] = torch.randn(shape, generator=g) * math.sqrt(1.0 / shape[0])
    return out


def _bn(x: torch.Tensor, scale: torch.Tensor, offset: torch.Tensor, residual=None, relu: bool = False) -> torch.Tensor:
    """Training-mode BN (+ residual add + ReLU): ``native.batch_norm_train`` -- the fused kernels of
    ``csrc/nn_kernels.cu`` when enabled (``DTF_FUSED_BN=1``), the plain PyTorch formulation otherwise."""
    return native.batch_norm_train(x, scale, offset, residual=residual, relu=relu)


def resnet18_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, stem: str = "cifar") -> torch.Tensor:
    """x: [N, H, W, C] float32 -> logits [N, classes]."""
    if stem == "cifar":
        h = native.conv2d_nhwc(x, p["stem/conv"], (1, 1, 1, 1), "SAME")
    else:
        h = native.conv2d_nhwc(x, p["stem/conv"], (1, 2, 2, 1), "SAME")
    h = _bn(h, p["stem/bn_scale"], p["stem/bn_offset"], relu=True)
    if stem != "cifar":
        h = native.max_pool_nhwc(h, (1, 3, 3, 1), (1, 2, 2, 1), "SAME").contiguous()
    cin = 64
    for si, (c, stride) in enumerate(_STAGES):
        for bi in range(2):
            pre = "stage%d/block%d" % (si, bi)
            s = stride if bi == 0 else 1
            y = native.conv2d_nhwc(h, p[pre + "/conv1"], (1, s, s, 1), "SAME")
            y = _bn(y, p[pre + "/bn1_scale"], p[pre + "/bn1_offset"], relu=True)
            y = native.conv2d_nhwc(y, p[pre + "/conv2"], (1, 1, 1, 1), "SAME")
            if (pre + "/down_conv") in p:
                sc = native.conv2d_nhwc(h, p[pre + "/down_conv"], (1, s, s, 1), "SAME")
                sc = _bn(sc, p[pre + "/down_bn_scale"], p[pre + "/down_bn_offset"])
            else:
                sc = h
            h = _bn(y, p[pre + "/bn2_scale"], p[pre + "/bn2_offset"], residual=sc, relu=True)      # relu(bn(y) + shortcut)
            cin = c
    pooled = native.global_avg_pool(h)
    return native.linear(pooled.contiguous(), p["fc/w"], p["fc/b"])


def resnet18_loss(p: Dict[str, torch.Tensor], x: torch.Tensor, y_onehot: torch.Tensor, stem: str = "cifar") -> torch.Tensor:
    logits = resnet18_forward(p, x, stem)
    return native.softmax_xent(logits, y_onehot).mean()


def build_resnet18_graph(x, y_, num_classes: int = 10, stem: str = "cifar"):
    """Graph-API version (variables created in the same order -> same ps placement)."""
    import distributed_tensorflow_b200 as dtf

    def var(name, shape):
        if name.endswith("_scale"):
            init = dtf.ones_initializer()
        elif name.endswith("_offset") or name.endswith("/b"):
            init = dtf.zeros_initializer()
        else:
            init = dtf.variance_scaling_initializer(2.0, "fan_in")
        return dtf.get_variable(name, list(shape), initializer=init)
    p = {n: var(n, s) for n, s in resnet18_param_shapes(num_classes, stem)}

    def bn(t, pre):
        return dtf.nn.fused_batch_norm_train(t, p[pre + "_scale"], p[pre + "_offset"])
    s0 = [1, 1, 1, 1] if stem == "cifar" else [1, 2, 2, 1]
    h = dtf.nn.relu(bn(dtf.nn.conv2d(x, p["stem/conv"], s0, "SAME"), "stem/bn"))
    if stem != "cifar":
        h = dtf.nn.max_pool(h, [1, 3, 3, 1], [1, 2, 2, 1], "SAME")
    for si, (c, stride) in enumerate(_STAGES):
        for bi in range(2):
            pre = "stage%d/block%d" % (si, bi)
            s = stride if bi == 0 else 1
            y = dtf.nn.relu(bn(dtf.nn.conv2d(h, p[pre + "/conv1"], [1, s, s, 1], "SAME"), pre + "/bn1"))
            y = bn(dtf.nn.conv2d(y, p[pre + "/conv2"], [1, 1, 1, 1], "SAME"), pre + "/bn2")
            sc = bn(dtf.nn.conv2d(h, p[pre + "/down_conv"], [1, s, s, 1], "SAME"), pre + "/down_bn") \
                if (pre + "/down_conv") in p else h
            h = dtf.nn.relu(y + sc)
    pooled = dtf.reduce_mean(h, axis=[1, 2])
    logits = dtf.nn.xw_plus_b(pooled, p["fc/w"], p["fc/b"])
    loss = dtf.reduce_mean(dtf.nn.softmax_cross_entropy_with_logits(labels=y_, logits=logits))
    return logits, loss, p
