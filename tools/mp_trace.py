"""Per-rank phase durations of the tf32 step inside a multi-GPU CUDA-graph replay (run under torchrun, one rank per GPU;
N workers on N GPUs, ps shard on rank 0's GPU).  %globaltimer is per GPU, so only DURATIONS on one GPU are compared:
worker kernel = [setup][token wait][post-token work ... fence + arrival]; ps_apply = [entry -> all arrivals seen][data pass]
[-> tokens released] (clock64 stamps of block 0 / the last block).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mp_trace.py  -> gpurun_out/mp_trace_N.json
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402
from tools.step_trace import ORDER, SLOTS  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nvls = {"0": False, "1": True}.get(os.environ.get("DTF_NVLS", "auto"), "auto")
    cfg = EngineConfig(num_ps=1, num_workers=world, optimizer={"kind": "sgd", "lr": 0.001}, nvls=nvls, ps_on_workers=True)
    eng = PSTrainEngine(MLPSpec(), cfg, Fabric.from_torch_distributed())
    eng.init_params()
    xs, ys = synthetic_mnist(20000, seed=1)
    eng.attach_dataset(rank, xs, ys)
    d = eng._w[rank]
    G = eng.step_ctas
    tr = eng.ranks[rank].bufs["steptrace_w%d" % rank]
    d["step_ds"].trace = tr.ptr
    tp = None
    if rank == 0:
        tp = torch.zeros(16, dtype=torch.int64, device="cuda")
        eng._p[0].phase_trace = tp.data_ptr()
    eng.enqueue_local_steps(10, "dataset")
    eng.synchronize()
    dist.barrier()
    eng.capture_graphs(20, "dataset")
    eng.replay_graphs(3)
    eng.synchronize()
    dist.barrier()
    eng.check_errors()
    t = tr.tensor(torch.int64, 0, 16 * 32).view(16, 32)[:G].cpu()
    t0 = int(t[:, 0].min())
    per = [[(int(v) - t0) / 1e3 if int(v) else None for v in row[:len(SLOTS)]] for row in t]
    med = [sorted(c[i] for c in per if c[i] is not None)[G // 2] for i in range(len(SLOTS))]
    rep = {"rank": rank, "worker_kernel_us": {SLOTS[i]: round(med[i], 2) for i in ORDER}}
    if tp is not None:
        pv = tp.cpu().tolist()
        ghz = 1.965
        rep["ps_apply_block0_us"] = {"arrivals_seen": round((pv[1] - pv[0]) / ghz / 1e3, 2), "slice_done": round((pv[2] - pv[0]) / ghz / 1e3, 2),
                                     "block0_end": round((pv[5] - pv[0]) / ghz / 1e3, 2)}
        ring = eng.ranks[0].bufs["trace0"].tensor(torch.int64, 0, eng.cfg.trace_cap * 4).view(-1, 4).cpu().tolist()
        rows = sorted((r[3], r[1], r[2]) for r in ring if r[0] == 1)[-8:]
        rep["ps_apply_entry_to_tokens_us"] = [round((b - a) / 1e3, 2) for _, a, b in rows]
        rep["ps_apply_period_us"] = [round((rows[i + 1][2] - rows[i][2]) / 1e3, 2) for i in range(len(rows) - 1)]
    allr = [None] * world
    dist.all_gather_object(allr, rep)
    if rank == 0:
        out = {"world": world, "nvls": bool(getattr(eng, "nvls", False)), "ranks": allr}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mp_trace_%d.json" % world), "w"), indent=1)
        for r in allr:
            print("rank", r["rank"])
            prev = 0.0
            for k, v in r["worker_kernel_us"].items():
                print("   %-34s %7.2f (+%.2f)" % (k, v, v - prev))
                prev = v
            for k in ("ps_apply_block0_us", "ps_apply_entry_to_tokens_us", "ps_apply_period_us"):
                if k in r:
                    print("  ", k, r[k])
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
