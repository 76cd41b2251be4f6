"""In-repo baseline: the same parameter-server workflow written with stock PyTorch pieces only
(``torch.distributed`` NCCL collectives, cuBLAS matmuls through ``torch.matmul``, torch ops for the
loss and the optimizer, CUDA graphs when capture succeeds).  It is what "calling the libraries" gives on
this box, i.e. the number the hand-written fabric engine has to beat.

Protocol per step (sync replicas, ``replicas_to_aggregate = num_workers``):
  pull  : ``dist.broadcast(flat_params, src=ps)``                (C1)
  work  : forward/backward of the 784-H-10 MLP on the worker     (K1-K4; autograd-free manual backward)
  push  : ``dist.reduce(flat_grads, dst=ps, op=SUM)``            (C2)
  apply : mean + SGD/Adam on the ps, ``global_step += 1``        (K5-K7)
The blocking collectives play the role of the token barrier.  With one GPU the ps and the worker are the
same process and the collectives disappear.  Topology follows the measured arm: ``ps_on_workers`` = every
rank trains and rank 0 additionally applies (N workers); otherwise rank 0 is a ps-only task (N - 1 workers).
Precision follows the measured arm too: ``tf32`` = fp32 tensors with cuBLAS TF32 matmuls
(``torch.backends.cuda.matmul.allow_tf32``), ``bf16`` = bf16 operands.

Measured with the same protocol as ``bench.py``: repetitions of EXACTLY K graph-replayed steps between CUDA events
(after a barrier and two untimed alignment steps), MAX over ranks, median repetition, >= ``min_ms`` of timed region;
``e2e`` = one step per call with the H2D copy of the batch from pinned memory and a D2H read of the loss.
"""
from __future__ import annotations

import math
import os
import statistics
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_nccl_baseline(args, rank: int, world: int, local_rank: int, sampler_cls=None, images=None, labels=None):
    if sampler_cls is None:
        from bench import ClockSampler as sampler_cls
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    dev = torch.device("cuda", local_rank)
    D, H, C, B = 784, args.hidden, 10, args.batch
    N = args.gpus
    pow_ = bool(getattr(args, "ps_on_workers", 1)) and not bool(getattr(args, "ps_only_task", 0))
    tf32 = getattr(args, "precision", "tf32") == "tf32"
    is_ps = rank == 0
    is_worker = (N == 1) or pow_ or rank > 0
    num_workers = N if (pow_ or N == 1) else N - 1
    widx = rank if (pow_ or N == 1) else max(rank - 1, 0)
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
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

    if is_worker:
        if images is None:
            images, labels = synthetic_mnist(args.num_train, seed=1)
        img_d = torch.as_tensor(images).to(dev)
        lab_d = torch.as_tensor(labels).to(dev)
    nb = args.num_train // B
    bidx = torch.zeros((), dtype=torch.int64, device=dev)       # device-side batch counter (graph-replayable)
    xs_stage = torch.zeros(B, D, device=dev)                    # e2e: this step's batch, copied from pinned host memory
    ys_stage = torch.zeros(B, C, device=dev)
    cast = (lambda t: t) if tf32 else (lambda t: t.bfloat16())

    def fwd_bwd(x, y):
        xc = cast(x)
        w1c, w2c = cast(pw1), cast(pw2)
        h = torch.relu((xc @ w1c).float() + pb1)
        hc = cast(h)
        logits = (hc @ w2c).float() + pb2
        p = torch.softmax(logits, -1)
        loss_buf.copy_(-(y * torch.log(torch.clamp(p, 1e-10, 1.0))).sum())
        dl = torch.where(p >= 1e-10, p * y.sum(-1, keepdim=True) - y, p * 0)
        dlc = cast(dl)
        gw2.copy_((hc.t() @ dlc).float())
        gb2.copy_(dl.sum(0))
        dh = cast((dlc @ w2c.t()).float() * (h > 0))
        gb1.copy_(dh.float().sum(0))
        gw1.copy_((xc.t() @ dh).float())

    def worker_from_dataset():
        i = (bidx * num_workers + widx) % nb
        idx = i * B + torch.arange(B, device=dev)
        fwd_bwd(img_d.index_select(0, idx), lab_d.index_select(0, idx))
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

    def step(staged=False):
        if world > 1:
            dist.broadcast(flat, src=0)              # pull
        if is_worker:
            if staged:
                fwd_bwd(xs_stage, ys_stage)
            else:
                worker_from_dataset()
        elif world > 1:
            grad.zero_()
        if world > 1:
            dist.reduce(grad, dst=0, op=dist.ReduceOp.SUM)   # push
        if is_ps:
            ps_apply()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(dev)
    K, W = args.steps, max(args.warmup, 3)
    unroll = next(u for u in (20, 16, 10, 8, 5, 4, 2, 1) if K % u == 0)
    with torch.cuda.stream(stream):
        for _ in range(W):
            step()
            step(staged=True)
    stream.synchronize()
    sync_all()
    graph = graph2 = graph_e = None
    graphed = False
    if not args.no_graph:
        try:
            graph, graph2, graph_e = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(unroll):
                    step()
            with torch.cuda.graph(graph2, stream=stream):
                for _ in range(2):
                    step()
            with torch.cuda.graph(graph_e, stream=stream):
                step(staged=True)
            with torch.cuda.stream(stream):
                graph.replay()
                graph2.replay()
                graph_e.replay()
            stream.synchronize()
            graphed = True
        except Exception as e:  # noqa: BLE001
            graph = graph2 = graph_e = None
            if rank == 0:
                print("nccl baseline: CUDA graph capture failed (%s); running eagerly" % type(e).__name__, file=sys.stderr)
            torch.cuda.synchronize()
    sync_all()

    def k_steps():
        if graphed:
            for _ in range(K // unroll):
                graph.replay()
        else:
            for _ in range(K):
                step()

    def align():
        if graphed:
            graph2.replay()
        else:
            step()
            step()

    def measure(body, min_ms, max_reps):
        times = []
        total = 0.0
        while len(times) < 3 or (total < min_ms and len(times) < max_reps):
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                align()
                e0.record(stream)
                body()
                e1.record(stream)
            stream.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t[0]))
            total += times[-1]
        return times

    sampler = sampler_cls(local_rank)
    sampler.start()
    time.sleep(0.25)
    t0 = time.time()
    times = measure(k_steps, getattr(args, "min_ms", 100.0), getattr(args, "max_reps", 400))
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms = statistics.median(times)

    # ---- e2e: one step per call, H2D of the batch from pinned host memory + D2H read of the loss, every step ----------------
    e2e = None
    if getattr(args, "e2e_steps", -1) != 0:
        n_use = min(args.num_train, 20000)
        hx = hy = None
        if is_worker:
            hx = torch.as_tensor(images[:n_use]).pin_memory()
            hy = torch.as_tensor(labels[:n_use]).pin_memory()
        nbe = n_use // B
        last = [None]
        ctr = [0]

        def e2e_k_steps():
            for _ in range(K):
                if is_worker:
                    b = (ctr[0] * num_workers + widx) % nbe
                    xs_stage.copy_(hx[b * B:(b + 1) * B], non_blocking=True)
                    ys_stage.copy_(hy[b * B:(b + 1) * B], non_blocking=True)
                if graphed:
                    graph_e.replay()
                else:
                    step(staged=True)
                if is_worker:
                    last[0] = float(loss_buf.item())     # D2H read of this step's loss
                ctr[0] += 1
        etimes = measure(e2e_k_steps, getattr(args, "min_ms", 100.0), getattr(args, "max_reps", 400))
        ems = statistics.median(etimes)
        e2e = {"value": num_workers * B * K / (ems / 1e3), "unit": "samples/sec", "ms_per_step": ems / K, "reps": len(etimes),
               "h2d_bytes_per_step": B * (D + C) * 4, "d2h_bytes_per_step": 4, "last_loss": last[0]}
    torch.backends.cuda.matmul.allow_tf32 = old_tf32
    return {
        "metric": "MNIST MLP samples/sec (whole box, device-timed, max over ranks), sync-replica PS",
        "value": num_workers * B * K / (ms / 1e3), "unit": "samples/sec", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "reps": len(times), "timed_ms_total": sum(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32" if tf32 else "bf16",
        "data": "synthetic MNIST-shaped 28x28 (55000x784 fp32 in HBM), random-init weights",
        "impl": "nccl-baseline (torch.distributed NCCL broadcast/reduce + cuBLAS %s + torch ops%s)" % (
            "TF32" if tf32 else "bf16", ", CUDA graph" if graphed else ", eager"),
        "config": {"model": "MNIST MLP 784-%d-10, clipped batch-sum xent" % H, "global_batch": num_workers * B,
                   "parallelism": ("ps1+worker%d%s" % (num_workers, " (rank 0 trains and applies)" if pow_ else "")) if N > 1 else "single GPU",
                   "optimizer": args.optimizer, "cuda_graph_unroll": unroll if graphed else 0},
        "clocks": clocks, "e2e": e2e, "global_step": int(gstep.item()), "final_loss": float(loss_buf.item()) if is_worker else None,
    }
