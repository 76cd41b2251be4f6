"""Library baseline for BASELINE config 5: ResNet-18 (CIFAR stem) under the same ps workflow with stock PyTorch pieces --
cuDNN convolutions / batch norm (channels-last, bf16 autocast), autograd, ``torch.distributed`` NCCL broadcast (pull) and
reduce (push) of one flat fp32 buffer to rank 0, momentum SGD applied there.  Same shapes, batch, precision class and
timing protocol as ``bench.py --model resnet18`` (every step copies its batch from pinned host memory and reads the loss
back; CUDA events, barrier + synchronize on both sides, max over ranks)."""
from __future__ import annotations

import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_STAGES = [(64, 1), (128, 2), (256, 2), (512, 2)]


def _shapes(num_classes=10):
    out = [("stem/conv", (64, 3, 3, 3)), ("stem/bn_s", (64,)), ("stem/bn_o", (64,))]
    cin = 64
    for si, (c, stride) in enumerate(_STAGES):
        for bi in range(2):
            p = "s%d/b%d" % (si, bi)
            out += [(p + "/c1", (c, cin, 3, 3)), (p + "/s1", (c,)), (p + "/o1", (c,)), (p + "/c2", (c, c, 3, 3)), (p + "/s2", (c,)),
                    (p + "/o2", (c,))]
            if bi == 0 and (stride != 1 or cin != c):
                out += [(p + "/dc", (c, cin, 1, 1)), (p + "/ds", (c,)), (p + "/do", (c,))]
            cin = c
    return out + [("fc/w", (num_classes, 512)), ("fc/b", (num_classes,))]


def _forward(p, x):
    def bn(h, s, o):
        return F.batch_norm(h, None, None, p[s], p[o], training=True, eps=1e-5)
    h = F.relu(bn(F.conv2d(x, p["stem/conv"], padding=1), "stem/bn_s", "stem/bn_o"))
    for si, (c, stride) in enumerate(_STAGES):
        for bi in range(2):
            q = "s%d/b%d" % (si, bi)
            s = stride if bi == 0 else 1
            y = F.relu(bn(F.conv2d(h, p[q + "/c1"], stride=s, padding=1), q + "/s1", q + "/o1"))
            y = bn(F.conv2d(y, p[q + "/c2"], padding=1), q + "/s2", q + "/o2")
            sc = bn(F.conv2d(h, p[q + "/dc"], stride=s), q + "/ds", q + "/do") if (q + "/dc") in p else h
            h = F.relu(y + sc)
    return F.linear(h.mean(dim=(2, 3)), p["fc/w"], p["fc/b"])


def run_nccl_resnet(args, rank: int, world: int, local_rank: int, sampler_cls=None):
    if sampler_cls is None:
        from bench import ClockSampler as sampler_cls
    dev = torch.device("cuda", local_rank)
    N = args.gpus
    pow_ = (bool(getattr(args, "ps_on_workers", 1)) and not bool(getattr(args, "ps_only_task", 0))) or N == 1
    is_worker = pow_ or rank > 0
    num_workers = N if pow_ else N - 1
    B = args.batch if args.batch != 100 else 64
    shapes = _shapes()
    sizes = [math.prod(s) for _, s in shapes]
    P = sum(sizes)
    flat = torch.zeros(P, device=dev)
    g = torch.Generator().manual_seed(2)
    o = 0
    for (name, shp), n in zip(shapes, sizes):
        if len(shp) == 4:
            flat[o:o + n] = (torch.randn(shp, generator=g) * math.sqrt(2.0 / (shp[1] * shp[2] * shp[3]))).flatten().to(dev)
        elif name.endswith(("s1", "s2", "ds", "bn_s")):
            flat[o:o + n] = 1.0
        elif name == "fc/w":
            flat[o:o + n] = (torch.randn(shp, generator=g) / math.sqrt(512)).flatten().to(dev)
        o += n
    if world > 1:
        dist.broadcast(flat, src=0)
    grad = torch.zeros(P, device=dev)
    mom = torch.zeros(P, device=dev)
    lr = args.lr or 0.05

    def views(t, requires_grad=False):
        out, o = {}, 0
        for (name, shp), n in zip(shapes, sizes):
            v = t[o:o + n].view(shp)
            if len(shp) == 4:
                v = v.contiguous(memory_format=torch.channels_last)
            out[name] = v.detach().requires_grad_(requires_grad)
            o += n
        return out
    nb = 64
    gen = torch.Generator().manual_seed(5 + rank)
    hx = torch.randn(nb, B, 3, 32, 32, generator=gen).pin_memory() if is_worker else None
    hy = torch.randint(0, 10, (nb, B), generator=gen).pin_memory() if is_worker else None
    losses = []

    def step(i):
        if world > 1:
            dist.broadcast(flat, src=0)                                   # pull
        if is_worker:
            x = hx[i % nb].to(dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
            y = hy[i % nb].to(dev, non_blocking=True)
            p = views(flat, True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = F.cross_entropy(_forward(p, x).float(), y)
            gl = torch.autograd.grad(loss, [p[n] for n, _ in shapes])
            torch.cat([t.reshape(-1).float() for t in gl], out=grad)
            losses.append(float(loss))                                    # D2H read of the step's loss
        else:
            grad.zero_()
        if world > 1:
            dist.reduce(grad, dst=0, op=dist.ReduceOp.SUM)                # push
        if rank == 0:
            mom.mul_(0.9).add_(grad, alpha=1.0 / num_workers)
            flat.sub_(mom, alpha=lr)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    K, W = (args.steps if args.steps != 2000 else 20), max(3, min(args.warmup, 5))
    for i in range(W):
        step(i)
    sync()
    sampler = sampler_cls(local_rank)
    sampler.start()
    time.sleep(0.25)
    sync()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    sync()
    clocks = sampler.stop(t0, time.time())
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    return {"impl": "nccl-baseline (cuDNN conv / batch norm channels-last bf16 autocast + autograd + NCCL broadcast / reduce, eager)",
            "value": num_workers * B * K / (ms / 1e3), "unit": "samples/sec", "ms_per_step": ms / K, "steps": K, "clocks": clocks,
            "first_loss": losses[W] if len(losses) > W else None, "final_loss": losses[-1] if losses else None}
