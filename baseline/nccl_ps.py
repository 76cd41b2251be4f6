"""In-repo baseline: the same parameter-server workflow written with stock PyTorch pieces only
(``torch.distributed`` NCCL collectives, cuBLAS matmuls through ``torch.matmul`` in bf16, torch ops for the
loss and the optimizer, CUDA graphs when capture succeeds).  It is what "calling the libraries" gives on
this box, i.e. the number the hand-written fabric engine has to beat.

Protocol per step (sync replicas, ``replicas_to_aggregate = num_workers``):
  pull  : ``dist.broadcast(flat_params, src=ps)``                (C1)
  work  : forward/backward of the 784-H-10 MLP on the worker     (K1-K4; autograd-free manual backward)
  push  : ``dist.reduce(flat_grads, dst=ps, op=SUM)``            (C2; the ps contributes zeros)
  apply : mean + SGD/Adam on the ps, ``global_step += 1``        (K5-K7)
The blocking collectives play the role of the token barrier.  With one GPU the ps and the worker are the
same process and the collectives disappear.
"""
from __future__ import annotations

import math
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_nccl_baseline(args, rank: int, world: int, local_rank: int):
    from bench import ClockSampler
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    dev = torch.device("cuda", local_rank)
    D, H, C, B = 784, args.hidden, 10, args.batch
    N = args.gpus
    is_ps = rank == 0
    is_worker = (N == 1) or rank > 0
    num_workers = max(N - 1, 1)
    sizes = [D * H, H, H * C, C]
    P = sum(sizes)
    flat = torch.zeros(P, device=dev)
    grad = torch.zeros(P, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    if is_ps:
        w1 = torch.empty(D, H)
        torch.nn.init.trunc_normal_(w1, 0, 1 / 28, -2 / 28, 2 / 28, generator=g)
        w2 = torch.empty(H, C)
        sd = 1 / math.sqrt(H)
        torch.nn.init.trunc_normal_(w2, 0, sd, -2 * sd, 2 * sd, generator=g)
        flat.copy_(torch.cat([w1.flatten(), torch.zeros(H), w2.flatten(), torch.zeros(C)]))
    if world > 1:
        dist.broadcast(flat, src=0)

    def views(t):
        o = 0
        out = []
        for s, shp in zip(sizes, [(D, H), (H,), (H, C), (C,)]):
            out.append(t[o:o + s].view(shp))
            o += s
        return out
    pw1, pb1, pw2, pb2 = views(flat)
    gw1, gb1, gw2, gb2 = views(grad)
    lr = args.lr if args.lr is not None else (0.01 if args.optimizer == "adam" else 0.001)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    step_t = torch.zeros((), device=dev)
    gstep = torch.zeros((), dtype=torch.int64, device=dev)
    loss_buf = torch.zeros((), device=dev)

    images = labels = None
    if is_worker:
        xs, ys = synthetic_mnist(args.num_train, seed=1)
        images = torch.from_numpy(xs).to(dev)
        labels = torch.from_numpy(ys).to(dev)
    nb = args.num_train // B
    widx = max(rank - 1, 0)
    bidx = torch.zeros((), dtype=torch.int64, device=dev)       # device-side batch counter (graph-replayable)

    def worker_fwd_bwd():
        i = (bidx * num_workers + widx) % nb
        idx = i * B + torch.arange(B, device=dev)
        x = images.index_select(0, idx).bfloat16()
        y = labels.index_select(0, idx)
        w1b, w2b = pw1.bfloat16(), pw2.bfloat16()
        h = torch.relu((x @ w1b).float() + pb1)
        hb = h.bfloat16()
        logits = (hb @ w2b).float() + pb2
        p = torch.softmax(logits, -1)
        loss_buf.copy_(-(y * torch.log(torch.clamp(p, 1e-10, 1.0))).sum())
        dl = torch.where(p >= 1e-10, p * y.sum(-1, keepdim=True) - y, p * 0)
        dlb = dl.bfloat16()
        gw2.copy_((hb.t() @ dlb).float())
        gb2.copy_(dl.sum(0))
        dh = ((dlb @ w2b.t()).float() * (h > 0)).bfloat16()
        gb1.copy_(dh.float().sum(0))
        gw1.copy_((x.t() @ dh).float())
        bidx.add_(1)

    def ps_apply():
        gmean = grad / float(num_workers)
        if args.optimizer == "sgd":
            flat.sub_(gmean, alpha=lr)
        elif args.optimizer == "momentum":
            m.mul_(0.9).add_(gmean)
            flat.sub_(m, alpha=lr)
        else:
            step_t.add_(1)
            m.mul_(0.9).add_(gmean, alpha=0.1)
            v.mul_(0.999).addcmul_(gmean, gmean, value=0.001)
            lr_t = lr * torch.sqrt(1 - 0.999 ** step_t) / (1 - 0.9 ** step_t)
            flat.sub_(lr_t * m / (v.sqrt() + 1e-8))
        gstep.add_(1)

    def step():
        if world > 1:
            dist.broadcast(flat, src=0)              # pull
        if is_worker:
            worker_fwd_bwd()
        elif world > 1:
            grad.zero_()
        if world > 1:
            dist.reduce(grad, dst=0, op=dist.ReduceOp.SUM)   # push
        if is_ps:
            ps_apply()

    stream = torch.cuda.Stream(dev)
    K, W = args.steps, max(args.warmup, 3)
    graph = None
    unroll = next(u for u in (20, 16, 10, 8, 5, 4, 2, 1) if K % u == 0)
    with torch.cuda.stream(stream):
        for _ in range(W):
            step()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    graphed = False
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(unroll):
                    step()
            with torch.cuda.stream(stream):
                graph.replay()
            stream.synchronize()
            graphed = True
        except Exception as e:  # noqa: BLE001
            graph = None
            if rank == 0:
                print("nccl baseline: CUDA graph capture failed (%s); running eagerly" % type(e).__name__, file=sys.stderr)
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        if graphed:
            for _ in range(K // unroll):
                graph.replay()
        else:
            for _ in range(K):
                step()
        e1.record(stream)
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    return {
        "metric": "MNIST MLP samples/sec (whole box, device-timed, max over ranks), sync-replica PS",
        "value": num_workers * B * K / (ms / 1e3), "unit": "samples/sec", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic MNIST-shaped 28x28 (55000x784 fp32 in HBM), random-init weights",
        "impl": "nccl-baseline (torch.distributed NCCL broadcast/reduce + cuBLAS bf16 + torch ops%s)" % (
            ", CUDA graph" if graphed else ", eager"),
        "config": {"model": "MNIST MLP 784-%d-10, clipped batch-sum xent" % H, "global_batch": num_workers * B,
                   "parallelism": "ps1+worker%d" % num_workers if N > 1 else "single GPU", "optimizer": args.optimizer,
                   "cuda_graph_unroll": unroll if graphed else 0},
        "clocks": clocks, "global_step": int(gstep.item()), "final_loss": float(loss_buf.item()) if is_worker else None,
    }
