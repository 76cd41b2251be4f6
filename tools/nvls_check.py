"""Fabric collectives on symmetric buffers: correctness + push/pull bandwidth (BASELINE metric "PS push/pull GB/s
vs 900 GB/s/dir").

    python tools/nvls_check.py --gpus 8                      # ONE process drives all GPUs (in-graph topology)
    torchrun --nproc-per-node 8 tools/nvls_check.py          # one process per GPU (between-graph; fd passing)

Rank 0 plays the ps.  pull = ps -> every GPU's replica (NVLS: one multimem.st stream, the switch fans out;
unicast: one peer store per worker).  push = sum of every worker's gradient copy into the ps (NVLS: one
multimem.ld_reduce stream, the switch adds; unicast: one peer load per worker).  Reported GB/s are PAYLOAD bytes
delivered per second: pull = (N-1) * bytes / t, push = (N-1) * bytes / t (what a worker-by-worker transfer would
have to move), plus the ps link bytes actually crossing its NVLink port.  Writes gpurun_out/nvls_check_N.json.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.ops import cuda_lib  # noqa: E402
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--mbytes", type=float, default=44.7, help="payload (default: ResNet-18 fp32 gradient, 44.7 MB)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", type=int, default=1, help="1: try several (CTAs, 16-byte accesses in flight per thread) and keep the best")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lib = cuda_lib.load()
    if world > 1:
        import torch.distributed as dist
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        fabric = Fabric.from_torch_distributed()
        N, rank0_local = world, dist.get_rank() == 0
    else:
        N = args.gpus or torch.cuda.device_count()
        fabric = Fabric(N, {r: r for r in range(N)})
        rank0_local = True
    level = fabric.nvls_level()
    nfl = int(args.mbytes * 1e6 / 4) // 1024 * 1024
    nbytes = nfl * 4
    out = {"n_gpus": N, "vmm_level": level, "payload_bytes": nbytes, "topology": "multi-process" if world > 1 else "single-process"}
    grads = fabric.alloc_symmetric("chk_grads", nbytes)
    repl = fabric.alloc_symmetric("chk_replica", nbytes)
    out["multicast"] = grads.multicast
    # every rank's gradient copy = (rank + 1) (rank 0, the ps, contributes zeros like in the engine)
    for r in fabric.local_ranks:
        t = grads.local(r).tensor(torch.float32, 0, nfl)
        t.fill_(float(r) if r > 0 else 0.0)
        torch.cuda.synchronize(t.device)
    fabric.barrier()
    expect = float(sum(range(1, N)))

    def timed(fn, dev):
        with torch.cuda.device(dev):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.iters * 1e-3

    sampler = None
    if rank0_local:
        dev = fabric.local_ranks[0]
        sys.path.insert(0, ROOT)
        from bench import ClockSampler
        import time
        sampler = ClockSampler(dev)
        sampler.start()
        t_start = time.time()
        with torch.cuda.device(dev):
            dst = torch.zeros(nfl, dtype=torch.float32, device="cuda")
            src = torch.arange(nfl, dtype=torch.float32, device="cuda")
        peers_g = (ctypes.c_void_p * 16)(*[grads.peer(0, r).ptr for r in range(1, N)])
        peers_r = (ctypes.c_void_p * 16)(*[repl.peer(0, r).ptr for r in range(1, N)])
        st = lambda: torch.cuda.current_stream().cuda_stream
        res = {}
        for mode in (["nvls", "unicast"] if grads.multicast else ["unicast"]):
            mc_g = grads.mc(0) if mode == "nvls" else None
            mc_r = repl.mc(0) if mode == "nvls" else None
            shapes = [(296, 4), (592, 8), (1184, 8), (592, 16), (1184, 16), (2368, 16)] if args.sweep else [(0, 0)]
            # ---- push/reduce ----
            best, tried = None, {}
            for grid, unroll in shapes:
                dst.zero_()
                t = timed(lambda: lib.dtf_fabric_reduce_ex(mc_g, peers_g, N - 1, dst.data_ptr(), nfl, grid, unroll, st()), dev)
                ok = bool(torch.all(dst == expect).item())
                tried["%dx%d" % (grid, unroll)] = round((nbytes if mode == "nvls" else (N - 1) * nbytes) / t / 1e9, 1)
                if ok and (best is None or t < best[0]):
                    best = (t, grid, unroll)
            t = best[0]
            res["push_" + mode] = {"ok": best is not None, "seconds": t, "payload_GBps": (N - 1) * nbytes / t / 1e9,
                                   "ps_port_GBps": (nbytes if mode == "nvls" else (N - 1) * nbytes) / t / 1e9,
                                   "ctas_x_in_flight": [best[1], best[2]], "ps_port_GBps_by_shape": tried}
            # ---- pull/broadcast ----
            best, tried = None, {}
            for grid, unroll in shapes:
                t = timed(lambda: lib.dtf_fabric_bcast_ex(src.data_ptr(), mc_r, peers_r, N - 1, nbytes, grid, unroll, st()), dev)
                tried["%dx%d" % (grid, unroll)] = round((nbytes if mode == "nvls" else (N - 1) * nbytes) / t / 1e9, 1)
                if best is None or t < best[0]:
                    best = (t, grid, unroll)
            t = best[0]
            res["pull_" + mode] = {"seconds": t, "payload_GBps": (N - 1) * nbytes / t / 1e9,
                                   "ps_port_GBps": (nbytes if mode == "nvls" else (N - 1) * nbytes) / t / 1e9,
                                   "ctas_x_in_flight": [best[1], best[2]], "ps_port_GBps_by_shape": tried}
            for k in ("push_" + mode, "pull_" + mode):
                res[k]["ps_port_fraction_of_900"] = res[k]["ps_port_GBps"] / 900.0
            torch.cuda.synchronize(dev)
            fabric.barrier() if world == 1 else None
            if world == 1:
                oks = []
                for r in range(1, N):
                    tr = repl.local(r).tensor(torch.float32, 0, nfl)
                    oks.append(bool(torch.equal(tr.cpu(), src.cpu())))
                    tr.zero_()
                res["pull_" + mode]["ok"] = all(oks)
        out["results"] = res
        out["clocks"] = sampler.stop(t_start, time.time())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        # workers verify the LAST broadcast landed in their copy
        ok = True
        for r in fabric.local_ranks:
            if r > 0:
                tr = repl.local(r).tensor(torch.float32, 0, nfl)
                ok = bool(torch.equal(tr.cpu(), torch.arange(nfl, dtype=torch.float32)))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["pull_verified_on_workers"] = bool(flag.item())
        if grads.multicast:
            # ---- sharded ps (one shard per GPU, what `bench.py --model resnet18` runs at N > 1): EVERY rank reduces its 1/N
            # slice of the gradient with multimem.ld_reduce and publishes its 1/N slice of the parameters with multimem.st, all
            # at once -- the switch-side reduce-scatter + all-gather.  Per GPU: egress = the (N-1)/N of its gradient copy the
            # other shards pull, ingress = 1/N; the time is the max over ranks.
            me = dist.get_rank()
            dev = fabric.local_ranks[me]
            sl = nfl // N // 1024 * 1024
            off = me * sl
            with torch.cuda.device(dev):
                dst_s = torch.zeros(sl, dtype=torch.float32, device="cuda")
                src_s = torch.arange(off, off + sl, dtype=torch.float32, device="cuda")
            none = (ctypes.c_void_p * 16)()
            stc = lambda: torch.cuda.current_stream().cuda_stream
            sh = {}
            repl.local(me).tensor(torch.float32, 0, nfl).zero_()        # (the single-ps pull above left the expected values)
            for name, fn in (("push", lambda: lib.dtf_fabric_reduce_ex(grads.mc(me) + off * 4, none, 0, dst_s.data_ptr(), sl, 0, 0, stc())),
                             ("pull", lambda: lib.dtf_fabric_bcast_ex(src_s.data_ptr(), repl.mc(me) + off * 4, none, 0, sl * 4, 0, 0, stc()))):
                torch.cuda.synchronize(dev)
                dist.barrier()
                t = timed(fn, dev)
                tt = torch.tensor([t], device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
                sh[name] = {"seconds": t, "gradient_GBps": N * sl * 4 / t / 1e9,
                            "per_gpu_egress_GBps" if name == "push" else "per_gpu_ingress_GBps": (N - 1) * sl * 4 / t / 1e9,
                            "fraction_of_900": (N - 1) * sl * 4 / t / 1e9 / 900.0}
            okp = bool(torch.all(dst_s == expect).item())
            dist.barrier()
            tr = repl.local(me).tensor(torch.float32, 0, N * sl)
            okl = bool(torch.equal(tr.cpu(), torch.arange(N * sl, dtype=torch.float32)))
            flag = torch.tensor([1 if (okp and okl) else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            sh["ok"] = bool(flag.item())
            sh["slice_bytes"] = sl * 4
            out["sharded_nvls"] = sh
    if rank0_local:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "nvls_check_%d_%s.json" % (N, "mp" if world > 1 else "sp")), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    fabric.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
