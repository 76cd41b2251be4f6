"""Dispatch layer between graph kernels and the sm_100a CUDA library.

CPU tensors run plain PyTorch (so the full API is testable without a GPU).
CUDA tensors run the hand-written kernels from ``csrc/`` through
``ops/cuda_lib.py`` wrapped in ``torch.autograd.Function`` so the graph
executor's reverse pass works unchanged:

* ``matmul`` / ``linear``  -> tcgen05 GEMM (bf16 operands from TMA, fp32 accumulate in TMEM,
  bias/ReLU epilogue) for forward *and* both backward GEMMs (K1, K4, K11);
* ``clipped_softmax_xent_sum`` / ``softmax_xent`` -> fused softmax + cross-entropy
  forward + dlogits (K2, K3);
* ``conv2d_nhwc`` -> im2col gather + tcgen05 GEMM (K14).

If a tensor is on a CUDA device and the library failed to load, these raise
(no silent eager fallback on a GPU box).  ``DTF_FORCE_EAGER=1`` switches the
CUDA path to cuBLAS/eager PyTorch for A/B debugging only.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

_FORCE_EAGER = os.environ.get("DTF_FORCE_EAGER", "0") == "1"


def _use_native(t: torch.Tensor) -> bool:
    if _FORCE_EAGER:
        return False
    if t.is_cuda:
        return True
    import sys
    mod = sys.modules.get(__package__ + ".cuda_lib")       # host tensors: only under the kernel emulation (tests)
    return bool(mod is not None and mod.EMULATION)


def _lib():
    from . import cuda_lib
    return cuda_lib


# ---------------------------------------------------------------------------
# GEMM family
# ---------------------------------------------------------------------------
class _MatMulFn(torch.autograd.Function):
    """C = op(A) @ op(B) on the tcgen05 GEMM; backward = two more tcgen05 GEMMs."""

    @staticmethod
    def forward(ctx, a, b, ta: bool, tb: bool):
        lib = _lib()
        ctx.save_for_backward(a, b)
        ctx.ta, ctx.tb = ta, tb
        return lib.gemm(a, b, ta, tb)

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        a, b = ctx.saved_tensors
        ta, tb = ctx.ta, ctx.tb
        g = g.contiguous()
        ga = gb = None
        if ctx.needs_input_grad[0]:
            # A' = op(A): dA' = g @ op(B)^T ; undo op on A
            ga = lib.gemm(b, g, tb, True) if ta else lib.gemm(g, b, False, not tb)
        if ctx.needs_input_grad[1]:
            gb = lib.gemm(g, a, True, ta) if tb else lib.gemm(a, g, not ta, False)
        return ga, gb, None, None


def matmul(a: torch.Tensor, b: torch.Tensor, ta: bool = False, tb: bool = False) -> torch.Tensor:
    if _use_native(a) and a.dim() == 2 and b.dim() == 2 and a.is_floating_point():
        return _MatMulFn.apply(a, b, ta, tb)
    x = a.t() if ta else a
    y = b.t() if tb else b
    return torch.matmul(x, y)


class _LinearFn(torch.autograd.Function):
    """y = act(x @ W + b), bias+ReLU applied in the GEMM epilogue straight out of TMEM."""

    @staticmethod
    def forward(ctx, x, w, b, relu: bool):
        lib = _lib()
        y = lib.gemm(x, w, False, False, bias=b, relu=relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        x, w, y = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = lib.relu_grad(g, y)
        gx = lib.gemm(g, w, False, True) if ctx.needs_input_grad[0] else None
        gw = lib.gemm(x, g, True, False) if ctx.needs_input_grad[1] else None
        gb = lib.colsum(g) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = False) -> torch.Tensor:
    if _use_native(x) and x.dim() == 2 and b is not None:
        return _LinearFn.apply(x, w, b, relu)
    y = torch.matmul(x, w)
    if b is not None:
        y = y + b
    return torch.relu(y) if relu else y


# ---------------------------------------------------------------------------
# softmax / cross-entropy
# ---------------------------------------------------------------------------
class _ClippedXentSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, clip_min: float):
        lib = _lib()
        loss, dlogits = lib.softmax_xent_fwd_bwd(logits, labels, clip_min, reduce_sum=True)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None


def clipped_softmax_xent_sum(logits: torch.Tensor, labels: torch.Tensor, clip_min: float = 1e-10) -> torch.Tensor:
    """``-sum(labels * log(clip(softmax(logits), clip_min, 1)))`` (reference distributed_mnist.py:112-113)."""
    if _use_native(logits) and logits.dim() == 2:
        return _ClippedXentSumFn.apply(logits, labels, float(clip_min))
    y = torch.softmax(logits.float(), dim=-1)
    return -(labels.float() * torch.log(torch.clamp(y, clip_min, 1.0))).sum()


class _SoftmaxXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lib = _lib()
        loss, dlogits = lib.softmax_xent_fwd_bwd(logits, labels, 0.0, reduce_sum=False)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g.unsqueeze(-1), None


def softmax_xent(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Per-row ``-sum(labels * log_softmax(logits))``."""
    if _use_native(logits) and logits.dim() == 2:
        return _SoftmaxXentFn.apply(logits, labels)
    return -(labels.float() * torch.log_softmax(logits.float(), dim=-1)).sum(dim=-1)


# ---------------------------------------------------------------------------
# batch normalisation (training mode, batch statistics) + residual add + ReLU
# ---------------------------------------------------------------------------
_FUSED_BN = os.environ.get("DTF_FUSED_NN", os.environ.get("DTF_FUSED_BN", "1")) == "1"      # csrc/nn_kernels.cu (validated on hardware in round 2)


def bn_train_reference(x, scale, offset, residual=None, relu=False, eps: float = 1e-5):
    """Plain PyTorch formulation (CPU oracle; CUDA path while DTF_FUSED_BN is off): statistics over every axis but
    the last, biased variance."""
    dims = tuple(range(x.dim() - 1))
    xf = x.float()
    mean = xf.mean(dim=dims, keepdim=True)
    var = (xf - mean).pow(2).mean(dim=dims, keepdim=True)
    y = (xf - mean) * torch.rsqrt(var + eps) * scale + offset
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y


def bn_backward_reference(dy, y, x2d, mean, rstd, scale, relu: bool):
    """The closed form the fused backward kernels implement, in PyTorch (tested against autograd on CPU):
    g = dy * [y > 0] (when ReLU was fused); doffset = sum g; dscale = sum g * xhat;
    dx = scale * rstd * (g - doffset / rows - xhat * dscale / rows); dresidual = g."""
    g = dy * (y > 0) if relu else dy
    xhat = (x2d - mean) * rstd
    doffset = g.sum(0)
    dscale = (g * xhat).sum(0)
    rows = x2d.shape[0]
    dx = scale * rstd * (g - doffset / rows - xhat * dscale / rows)
    return dx, dscale, doffset, g


class _FusedBNFn(torch.autograd.Function):
    """y = relu?(BN_train(x) (+ residual)) as two launches forward (statistics, apply) and two backward (sums, apply)."""

    @staticmethod
    def forward(ctx, x, scale, offset, residual, relu: bool, eps: float):
        lib = _lib()
        shape = x.shape
        x2 = x.float().contiguous().view(-1, shape[-1])
        r2 = residual.float().contiguous().view(-1, shape[-1]) if residual is not None else None
        y, mean, rstd = lib.bn_forward(x2, scale, offset, r2, relu, eps)
        ctx.save_for_backward(x2, mean, rstd, scale, y if relu else None)
        ctx.relu, ctx.has_res, ctx.shape = relu, residual is not None, shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        x2, mean, rstd, scale, y = ctx.saved_tensors
        g2 = g.contiguous().view(-1, ctx.shape[-1])
        want_res = ctx.has_res and ctx.needs_input_grad[3]
        dx, dscale, doffset, dres = lib.bn_backward(g2, y if ctx.relu else None, x2, mean, rstd, scale, want_res)
        return (dx.view(ctx.shape), dscale.view_as(scale), doffset.view_as(scale),
                dres.view(ctx.shape) if dres is not None else None, None, None)


def batch_norm_train(x: torch.Tensor, scale: torch.Tensor, offset: torch.Tensor, residual: Optional[torch.Tensor] = None,
                     relu: bool = False, eps: float = 1e-5) -> torch.Tensor:
    """Training-mode batch norm over the last (channel) axis, optionally fused with a residual add and ReLU."""
    if _use_native(x) and _FUSED_BN and x.shape[-1] % 4 == 0 and scale.dim() == 1:
        return _FusedBNFn.apply(x, scale, offset, residual, relu, float(eps))
    return bn_train_reference(x, scale, offset, residual, relu, eps)


# ---------------------------------------------------------------------------
# pooling (NHWC)
# ---------------------------------------------------------------------------
def _fused_nn() -> bool:
    return bool(_lib().FUSED_NN)


class _MaxPoolFn(torch.autograd.Function):
    """Window max with a byte argmax per element; the backward gathers (deterministic, no atomics)."""

    @staticmethod
    def forward(ctx, x, kh, kw, strides, pads):
        y, arg = _lib().maxpool_nhwc(x, kh, kw, strides, pads)
        ctx.save_for_backward(arg)
        ctx.meta = (tuple(x.shape), kh, kw, strides, pads)
        return y

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        xshape, kh, kw, strides, pads = ctx.meta
        return _lib().maxpool_nhwc_bwd(g, arg, xshape, kh, kw, strides, pads), None, None, None, None


def max_pool_nhwc(x: torch.Tensor, ksize: Sequence[int], strides: Sequence[int], padding: str) -> torch.Tensor:
    """``tf.nn.max_pool`` on NHWC data (padding counts as -inf)."""
    kh, kw = int(ksize[1]), int(ksize[2])
    sh, sw = int(strides[1]), int(strides[2])
    if padding.upper() == "SAME":
        pt, pb = _same_pad(x.shape[1], kh, sh)
        pl, pr = _same_pad(x.shape[2], kw, sw)
    else:
        pt = pb = pl = pr = 0
    if _use_native(x) and _fused_nn() and x.shape[-1] % 4 == 0 and kh * kw <= 255:
        return _MaxPoolFn.apply(x, kh, kw, (sh, sw), (pt, pb, pl, pr))
    xn = x.permute(0, 3, 1, 2)
    if pt or pb or pl or pr:
        xn = F.pad(xn, (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(xn, (kh, kw), (sh, sw)).permute(0, 2, 3, 1)


class _GlobalAvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.xshape = tuple(x.shape)
        return _lib().global_avgpool_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return _lib().global_avgpool_nhwc_bwd(g, ctx.xshape)


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """[N, H, W, C] -> [N, C] mean over the spatial axes (the pooling in front of ResNet's classifier)."""
    if _use_native(x) and _fused_nn() and x.dim() == 4 and x.shape[-1] % 4 == 0:
        return _GlobalAvgPoolFn.apply(x)
    return x.mean(dim=(1, 2))


# ---------------------------------------------------------------------------
# convolution (NHWC data, HWIO filter)
# ---------------------------------------------------------------------------
def _same_pad(size: int, k: int, s: int):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


class _ConvFn(torch.autograd.Function):
    """Convolution on the tcgen05 GEMM.

    Stride 1 with whole 64-channel chunks (every 3x3 convolution of a ResNet stage after its first): IMPLICIT GEMM -- the
    activation is rounded to bf16 once ([N, H, W, C], 2 B / element) and the GEMM's TMA producer reads shifted 4-D boxes of
    it, one per (tap, channel chunk); the patch matrix (kh*kw times larger) is never written or read, and the box's
    out-of-bounds zero fill is the padding.  dX = the same kernel over dY with the flipped, in/out-swapped filter; dW =
    patches(x)^T . dY with the pixels as the (split) K dimension.
    Otherwise (strided, 3-channel stem): NHWC patches gathered to a [N*Ho*Wo, kh*kw*Cin] bf16 matrix by a hand-written gather
    kernel, then the plain GEMM against the [kh*kw*Cin, Cout] filter."""

    @staticmethod
    def forward(ctx, x, w, strides, pads):
        lib = _lib()
        kh, kw, cin, cout = w.shape
        w2 = w.reshape(kh * kw * cin, cout)
        n, h, wd = x.shape[0], x.shape[1], x.shape[2]
        pt, pb, pl, pr = pads
        same = (h + pt + pb - kh + 1 == h) and (wd + pl + pr - kw + 1 == wd)
        if lib.IMPLICIT_CONV and not lib.EMULATION and same and lib.implicit_conv_ok(n, h, wd, cin, strides) \
                and lib.implicit_conv_ok(n, h, wd, cout) and cout % 8 == 0:
            x16 = lib.to_bf16_padded(x.reshape(n * h * wd, cin))[0].view(n, h, wd, cin)
            y = lib.conv_igemm(x16, w2, kh, kw, pt, pl)
            ctx.save_for_backward(x16, w2)
            ctx.meta = (x.shape, w.shape, strides, pads, (n, h, wd), True)
            return y.reshape(n, h, wd, cout)
        cols, (n, ho, wo) = lib.im2col_nhwc(x, kh, kw, strides, pads)
        y = lib.gemm(cols, w2, False, False, precision="bf16")      # conv model family: bf16 operands (BASELINE config 5)
        ctx.save_for_backward(cols, w2)
        ctx.meta = (x.shape, w.shape, strides, pads, (n, ho, wo), False)
        return y.reshape(n, ho, wo, cout)

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        cols, w2 = ctx.saved_tensors
        xshape, wshape, strides, pads, (n, ho, wo), implicit = ctx.meta
        kh, kw, cin, cout = wshape
        pt, pb, pl, pr = pads
        g2 = g.reshape(n * ho * wo, cout).contiguous()
        gx = gw = None
        if implicit:
            x16 = cols
            g16 = lib.to_bf16_padded(g2)[0]                      # one rounding of dY serves both products
            if ctx.needs_input_grad[1]:
                gw = lib.conv_igemm(x16, g16, kh, kw, pt, pl, wgrad=True).reshape(wshape)
            if ctx.needs_input_grad[0]:
                wf = w2.reshape(kh, kw, cin, cout).flip(0, 1).permute(0, 1, 3, 2).reshape(kh * kw * cout, cin)
                gx = lib.conv_igemm(g16.view(n, ho, wo, cout), wf.contiguous(), kh, kw, kh - 1 - pt, kw - 1 - pl).reshape(xshape)
            return gx, gw, None, None
        if ctx.needs_input_grad[1]:
            gw = lib.gemm(cols, g2, True, False, precision="bf16").reshape(wshape)
        if ctx.needs_input_grad[0]:
            if lib.FUSED_NN and strides == (1, 1) and cout % 8 == 0:
                # stride 1: dX is itself a convolution of dY with the spatially flipped, in/out-swapped filter --
                #   dX[b, iy, ix, ci] = sum_{ky', kx', co} dY[b, iy + ky' - (kh-1-pt), ix + kx' - (kw-1-pl), co] * w[kh-1-ky', kw-1-kx', ci, co]
                # so it reuses the im2col + GEMM pair: a bf16 patch matrix of dY (rows x kh*kw*Cout x 2 B) replaces the fp32
                # [rows, kh*kw*Cin] product + the col2im gather (half the HBM traffic, one launch less, output written once)
                wf = w2.reshape(kh, kw, cin, cout).flip(0, 1).permute(0, 1, 3, 2).reshape(kh * kw * cout, cin)
                gcols_in, _ = lib.im2col_nhwc(g2.reshape(n, ho, wo, cout), kh, kw, (1, 1), (kh - 1 - pt, kh - 1 - pb, kw - 1 - pl, kw - 1 - pr))
                gx = lib.gemm(gcols_in, wf.contiguous(), False, False, precision="bf16").reshape(xshape)
            else:
                gcols = lib.gemm(g2, w2, False, True, precision="bf16")
                gx = lib.col2im_nhwc(gcols, xshape, kh, kw, strides, pads)
        return gx, gw, None, None


def conv2d_nhwc(x: torch.Tensor, w: torch.Tensor, strides: Sequence[int], padding: str,
                data_format: str = "NHWC") -> torch.Tensor:
    if data_format != "NHWC":
        raise ValueError("only NHWC is supported (channels-last is the tensor-core layout)")
    sh, sw = int(strides[1]), int(strides[2])
    kh, kw = int(w.shape[0]), int(w.shape[1])
    if padding.upper() == "SAME":
        pt, pb = _same_pad(x.shape[1], kh, sh)
        pl, pr = _same_pad(x.shape[2], kw, sw)
    else:
        pt = pb = pl = pr = 0
    if _use_native(x):
        return _ConvFn.apply(x, w, (sh, sw), (pt, pb, pl, pr))
    xn = x.permute(0, 3, 1, 2)
    if pt or pb or pl or pr:
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=(sh, sw))
    return y.permute(0, 2, 3, 1)


# ---------------------------------------------------------------------------
# element-wise graph ops and full reductions (K9): csrc/elementwise.cu ew_* kernels
# ---------------------------------------------------------------------------
# Reference programs: `y = weight * x + biase`, `tf.square(y_ - y)`, `tf.reduce_mean(...)` (example_between_graph.py:55-60),
# `tf.reduce_sum(y_ * tf.log(...))` (distributed_mnist.py:113).  fp32 tensors on a CUDA device (or host tensors under the
# kernel emulation) run our kernels forward AND backward; other dtypes / general broadcasts stay with torch.
_TORCH_BINARY = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div,
                 "sqdiff": lambda a, b: (a - b) * (a - b)}
_TORCH_UNARY = {"neg": torch.neg, "square": lambda x: x * x, "sqrt": torch.sqrt, "exp": torch.exp, "log": torch.log,
                "sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu}
EW_SELF_TEST = {"state": "not run"}
_EW_ENABLED = os.environ.get("DTF_NATIVE_ELEMENTWISE", "1") == "1"


def _ew_self_test(device: torch.device) -> bool:
    """First CUDA use in a process: every ew_* kernel against torch on small tensors (full, scalar and trailing-vector
    operands, a reduction longer than one block).  A mismatch is logged, recorded in ``EW_SELF_TEST`` and switches these ops
    back to torch for the process -- a wrong kernel must not silently train a wrong model."""
    if EW_SELF_TEST["state"] != "not run":
        return EW_SELF_TEST["state"] == "passed"
    lib = _lib()
    EW_SELF_TEST["state"] = "running"
    try:
        g = torch.Generator().manual_seed(3)
        a = (torch.rand(37, 29, generator=g) + 0.5).to(device)
        b = (torch.rand(37, 29, generator=g) + 0.5).to(device)
        v = (torch.rand(29, generator=g) + 0.5).to(device)
        s = (torch.rand(1, generator=g) + 0.5).to(device)
        worst = 0.0
        for op, fn in _TORCH_BINARY.items():
            for x, y in ((a, b), (a, v), (v, a), (a, s), (s, a)):
                worst = max(worst, float((lib.ew_binary(op, x, y) - fn(x, y)).abs().max()))
        for op, fn in _TORCH_UNARY.items():
            worst = max(worst, float((lib.ew_unary(op, a) - fn(a)).abs().max()))
        big = (torch.rand(70001, generator=g) - 0.5).to(device)
        for got, want in ((lib.ew_reduce_sum(big, 0.5), 0.5 * big.double().sum()),
                          (lib.ew_reduce_sum(big, 1.0, square=True), (big.double() ** 2).sum())):
            # fp32 partial sums combined by atomics: judged relative to the magnitudes summed, scaled onto the 2e-3 bound
            worst = max(worst, 20.0 * abs(float(got) - float(want)) / (1.0 + float(big.double().abs().sum())))
        worst = max(worst, float((lib.ew_affine(a, -2.0, 0.25) - (-2.0 * a + 0.25)).abs().max()))
        worst = max(worst, float((lib.ew_affine(s, 3.0, out_shape=(5, 7)) - (3.0 * s).expand(5, 7)).abs().max()))
        parts = [a[:, :11].contiguous(), a[:, 11:12].contiguous(), a[:, 12:].contiguous()]
        worst = max(worst, float((lib.concat(parts, 1) - a).abs().max()))
        worst = max(worst, float((lib.concat([a, b], 0) - torch.cat([a, b], 0)).abs().max()))
        o1, o2 = torch.empty(30, 29, device=device), torch.empty(7, 29, device=device)
        lib.scatter_rows(a, [o1, o2])
        worst = max(worst, float((torch.cat([o1, o2], 0) - a).abs().max()))
        ok = worst <= 2e-3 and worst == worst
        EW_SELF_TEST.update(state="passed" if ok else "failed", max_abs_diff=worst)
    except Exception as e:          # noqa: BLE001
        ok = False
        EW_SELF_TEST.update(state="failed", error=repr(e)[:300])
    if not ok:
        import logging
        logging.getLogger("dtf").error("element-wise kernels failed their self-test (%s): these graph ops run on torch in this "
                                       "process", EW_SELF_TEST)
    return ok


def _ew_ok(*ts) -> bool:
    if not _EW_ENABLED:
        return False
    dev = None
    for t in ts:
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or not _use_native(t):
            return False
        if dev is not None and t.device != dev:
            return False
        dev = t.device
    if _lib().EMULATION:
        return True
    if EW_SELF_TEST["state"] == "running":
        return False
    return _ew_self_test(dev)


def _unbroadcast(g: torch.Tensor, shape, out_shape) -> torch.Tensor:
    """Sum the gradient of a broadcast operand back to the operand's shape (full / scalar / trailing-dims vector)."""
    lib = _lib()
    mode = lib.ew_broadcast_mode(shape, out_shape, out_shape)
    if mode[0] == 0:
        return g.reshape(shape)
    if mode[0] == 1:
        return lib.ew_reduce_sum(g).reshape(shape)
    return lib.colsum(g.reshape(-1, mode[1])).reshape(shape)


class _EwBinaryFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, op: str):
        out = _lib().ew_binary(op, a, b)
        ctx.op = op
        ctx.save_for_backward(a, b, out if op == "div" else a.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        a, b, out = ctx.saved_tensors
        op = ctx.op
        g = g.contiguous()
        oshape = tuple(g.shape)
        ga = gb = None
        if op == "add":
            ga, gb = g, g
        elif op == "sub":
            ga, gb = g, (lib.ew_affine(g, -1.0) if ctx.needs_input_grad[1] else None)
        elif op == "mul":
            ga = lib.ew_binary("mul", g, b) if ctx.needs_input_grad[0] else None
            gb = lib.ew_binary("mul", g, a) if ctx.needs_input_grad[1] else None
        elif op == "div":
            q = lib.ew_binary("div", g, b)
            ga = q
            gb = lib.ew_affine(lib.ew_binary("mul", q, out), -1.0) if ctx.needs_input_grad[1] else None
        else:                                   # sqdiff: d/da (a - b)^2 = 2 (a - b)
            ga = lib.ew_affine(lib.ew_binary("mul", g, lib.ew_binary("sub", a, b)), 2.0)
            gb = lib.ew_affine(ga, -1.0) if ctx.needs_input_grad[1] else None
        ga = _unbroadcast(ga, tuple(a.shape), oshape) if (ga is not None and ctx.needs_input_grad[0]) else None
        gb = _unbroadcast(gb, tuple(b.shape), oshape) if (gb is not None and ctx.needs_input_grad[1]) else None
        return ga, gb, None


def binary(op: str, a, b):
    """Graph kernels Add / Sub / Mul / RealDiv / SquaredDifference."""
    if _ew_ok(a, b):
        lib = _lib()
        try:
            oshape = torch.broadcast_shapes(a.shape, b.shape)
        except RuntimeError:
            oshape = None
        if oshape is not None and lib.ew_broadcast_mode(a.shape, b.shape, oshape) is not None \
                and lib.ew_broadcast_mode(b.shape, a.shape, oshape) is not None and len(oshape) > 0:
            return _EwBinaryFn.apply(a, b, op)
    return _TORCH_BINARY[op](a, b)


class _EwUnaryFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op: str):
        y = _lib().ew_unary(op, x)
        ctx.op = op
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        x, y = ctx.saved_tensors
        op = ctx.op
        g = g.contiguous()
        if op == "neg":
            return lib.ew_affine(g, -1.0), None
        if op == "square":
            return lib.ew_affine(lib.ew_binary("mul", g, x), 2.0), None
        if op == "sqrt":
            return lib.ew_affine(lib.ew_binary("div", g, y), 0.5), None
        if op == "exp":
            return lib.ew_binary("mul", g, y), None
        if op == "log":
            return lib.ew_binary("div", g, x), None
        if op == "sigmoid":
            return lib.ew_binary("mul", lib.ew_binary("mul", g, y), lib.ew_affine(y, -1.0, 1.0)), None
        if op == "tanh":
            return lib.ew_binary("mul", g, lib.ew_affine(lib.ew_unary("square", y), -1.0, 1.0)), None
        return lib.relu_grad(g, y), None


def unary(op: str, x):
    """Graph kernels Neg / Square / Sqrt / Exp / Log / Sigmoid / Tanh / Relu."""
    if _ew_ok(x) and x.dim() > 0:
        return _EwUnaryFn.apply(x, op)
    return _TORCH_UNARY[op](x)


class _EwReduceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale: float):
        ctx.scale, ctx.shape = scale, tuple(x.shape)
        return _lib().ew_reduce_sum(x, scale)

    @staticmethod
    def backward(ctx, g):
        return _lib().ew_affine(g.reshape(1).contiguous(), ctx.scale, out_shape=ctx.shape), None


def reduce_all(x, mean: bool):
    """``tf.reduce_sum(x)`` / ``tf.reduce_mean(x)`` over every element -> 0-d tensor."""
    if _ew_ok(x) and x.numel() > 0 and x.dim() > 0:
        return _EwReduceFn.apply(x, (1.0 / x.numel()) if mean else 1.0)
    return x.mean() if mean else x.sum()


class _ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axis: int, *xs):
        ctx.axis, ctx.lens = axis, [x.shape[axis] for x in xs]
        return _lib().concat(xs, axis)

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for l in ctx.lens:                       # the gradient of a concat is a split: views, no copy
            outs.append(g.narrow(ctx.axis, off, l))
            off += l
        return (None,) + tuple(outs)


def concat(xs, axis: int):
    """Graph kernel ConcatV2 (K10): one gather kernel for 2..16 fp32 parts on /gpu (``example_in_graph.py:58``,
    ``standalone.py:86``); anything else is ``torch.cat``."""
    xs = list(xs)
    if 2 <= len(xs) <= 16 and _ew_ok(*xs) and all(x.dim() == xs[0].dim() and x.dim() > 0 for x in xs):
        nd = xs[0].dim()
        ax = axis % nd
        ref = [d for i, d in enumerate(xs[0].shape) if i != ax]
        if all([d for i, d in enumerate(x.shape) if i != ax] == ref for x in xs):
            return _ConcatFn.apply(ax, *xs)
    return torch.cat(xs, dim=axis)


class _MseFn(torch.autograd.Function):
    """mean((a - b)^2) over every element: the difference is kept for the backward pass, the square is folded into the reduction."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib()
        d = lib.ew_binary("sub", a, b)
        ctx.save_for_backward(d)
        ctx.shapes = (tuple(a.shape), tuple(b.shape))
        return lib.ew_reduce_sum(d, 1.0 / d.numel(), square=True)

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        (d,) = ctx.saved_tensors
        gd = lib.ew_binary("mul", d, g.reshape(1).contiguous())            # d * dL (scalar operand)
        oshape = tuple(d.shape)
        ga = _unbroadcast(lib.ew_affine(gd, 2.0 / d.numel()), ctx.shapes[0], oshape) if ctx.needs_input_grad[0] else None
        gb = _unbroadcast(lib.ew_affine(gd, -2.0 / d.numel()), ctx.shapes[1], oshape) if ctx.needs_input_grad[1] else None
        return ga, gb


def mse(a, b):
    """``tf.reduce_mean(tf.square(a - b))`` (``example_between_graph.py:59``, ``standalone.py:61``) as one fused op."""
    if _ew_ok(a, b):
        lib = _lib()
        try:
            oshape = torch.broadcast_shapes(a.shape, b.shape)
        except RuntimeError:
            oshape = None
        if oshape is not None and len(oshape) > 0 and lib.ew_broadcast_mode(a.shape, b.shape, oshape) is not None \
                and lib.ew_broadcast_mode(b.shape, a.shape, oshape) is not None:
            return _MseFn.apply(a, b)
    d = a - b
    return (d * d).mean()
