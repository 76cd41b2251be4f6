"""Multi-GPU correctness check of the fabric PS engine (run under torchrun, one rank per GPU):
sync (mean of N worker gradients per step, tokens) and async (every push applied, staleness counted)
against a CPU oracle that replays the same batches.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mp_check.py

``DTF_RTA=R`` (replicas_to_aggregate, R < workers): the backup-worker protocol instead -- no oracle (which replicas make an
aggregate is a race by design), the invariants: no device-side wait times out (the ``consumed`` handshake that keeps a straggler
from overwriting a push the ps may still read does not deadlock), exactly one global step per aggregate, R gradients folded per
aggregate, every other push dropped as stale or still pending, parameters finite and moved.
"""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402


PRECISION = os.environ.get("DTF_PRECISION", "tf32")


def grads(p, x, y):
    r = (lambda v: v.bfloat16().float()) if PRECISION == "bf16" else (lambda v: v)     # tf32 engines: the UNROUNDED fp32 model
    h = torch.relu(r(x) @ r(p["hid_w"]) + p["hid_b"])
    h16 = r(h)
    logits = h16 @ r(p["sm_w"]) + p["sm_b"]
    prob = torch.softmax(logits, -1)
    loss = -(y * torch.log(torch.clamp(prob, 1e-10, 1.0))).sum()
    dl = prob - y
    dhf = (dl @ r(p["sm_w"]).t()) * (h16 > 0)
    return {"sm_w": h16.t() @ dl, "sm_b": dl.sum(0), "hid_b": dhf.sum(0), "hid_w": r(x).t() @ r(dhf)}, float(loss)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    num_ps = int(os.environ.get("DTF_NUM_PS", "1"))
    nvls = os.environ.get("DTF_NVLS", "0")
    nvls = {"0": False, "1": True}.get(nvls, nvls)
    pow_ = os.environ.get("DTF_PS_ON_WORKERS", "0") == "1"      # every rank a worker, ps shards on the first ranks' GPUs
    W = world if pow_ else world - num_ps
    xs, ys = synthetic_mnist(100 * W * 8, seed=11)
    report = {}
    rta = int(os.environ.get("DTF_RTA", "0"))
    if rta:
        assert 0 < rta < W, "DTF_RTA must be below the number of workers (%d)" % W
        cfg = EngineConfig(num_ps=num_ps, num_workers=W, sync=True, replicas_to_aggregate=rta, optimizer={"kind": "sgd", "lr": 0.001},
                           seed=2, nvls=nvls, ps_on_workers=pow_, precision=PRECISION, timeout_ns=10_000_000_000)
        eng = PSTrainEngine(MLPSpec(), cfg, Fabric.from_torch_distributed())
        eng.init_params()
        p0 = eng.state_dict() if rank < num_ps else None
        for r in eng.ranks:
            if r in eng.worker_ranks:
                eng.attach_dataset(r, xs, ys)
        steps = 12
        eng.enqueue_local_steps(steps, "dataset")
        eng.synchronize()
        dist.barrier()
        err = None
        try:
            eng.check_errors()
        except RuntimeError as e:
            err = str(e)
        errs = [None] * world
        dist.all_gather_object(errs, err)
        if rank == 0:
            sd = eng.state_dict()
            applied, dropped = eng.read_ctl(0, "applied_total"), eng.read_ctl(0, "dropped_stale")
            moved = max(float((sd[k] - p0[k]).abs().max()) for k in ("hid_w", "hid_b", "sm_w", "sm_b"))
            finite = all(bool(torch.isfinite(sd[k]).all()) for k in ("hid_w", "hid_b", "sm_w", "sm_b"))
            report["backup_workers"] = {
                "replicas_to_aggregate": rta, "workers": W, "global_step": int(sd["global_step"]), "applied_total": int(applied),
                "dropped_stale": int(dropped), "errors": [e for e in errs if e], "moved": moved,
                "ok": (not any(errs)) and int(sd["global_step"]) == steps and int(applied) == steps * rta
                and int(dropped) <= steps * (W - rta) and finite and moved > 0}
            report.update(world=world, num_ps=num_ps, nvls=str(nvls), ps_on_workers=pow_, precision=PRECISION)
            print("MP_CHECK " + json.dumps(report))
        dist.barrier()
        eng.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    for mode in ("sync", "async"):
        cfg = EngineConfig(num_ps=num_ps, num_workers=W, sync=(mode == "sync"), optimizer={"kind": "sgd", "lr": 0.001},
                           seed=2, nvls=nvls, ps_on_workers=pow_, precision=PRECISION)
        eng = PSTrainEngine(MLPSpec(), cfg, Fabric.from_torch_distributed())
        eng.init_params()
        p0 = None
        if rank < num_ps:
            p0 = eng.state_dict()
        for r in eng.ranks:
            if r in eng.worker_ranks:
                eng.attach_dataset(r, xs, ys)
        steps = 6
        eng.enqueue_local_steps(steps, "dataset")
        eng.synchronize()
        dist.barrier()
        eng.check_errors()
        sd = eng.state_dict() if rank < num_ps else {}
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v for k, v in sd.items()})
        p0s = [None] * world
        dist.all_gather_object(p0s, p0)
        loss = eng.read_loss() if rank in eng.worker_ranks else None
        if rank == 0:
            final = {}
            for g in gathered[:num_ps]:
                final.update(g)
            init = {}
            for g in p0s[:num_ps]:
                init.update({k: v for k, v in g.items() if k in ("hid_w", "hid_b", "sm_w", "sm_b")})
            nb = xs.shape[0] // 100
            if mode == "sync":
                p = {k: v.clone() for k, v in init.items()}
                for t in range(steps):
                    acc = None
                    for w in range(W):
                        b = (t * W + w) % nb
                        g, _ = grads(p, torch.from_numpy(xs[b * 100:(b + 1) * 100]), torch.from_numpy(ys[b * 100:(b + 1) * 100]))
                        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
                    for k in p:
                        p[k] = p[k] - 0.001 * acc[k] / W
                errs = {k: float((final[k] - p[k]).abs().max() / (p[k].abs().max() + 1e-6)) for k in p}
                nerrs = {k: float((final[k].double() - p[k].double()).norm() / (p[k].double() - init[k].double()).norm()) for k in p}
                err = max(errs.values())
                report[mode] = {"global_step": int(final["global_step"]), "max_rel_err_vs_oracle": err, "per_var": errs,
                                "update_norm_rel_err": nerrs,
                                # per-variable max-abs error relative to the variable's largest entry.  The bias vectors start at
                                # zero and have moved ~1e-3 after 6 steps, so one ReLU gate that lands on the other side of zero
                                # (TF32 / bf16 rounding of a pre-activation) shows up as percent-level there; the matrices bound it
                                "ok": int(final["global_step"]) == steps and errs["hid_w"] < 5e-3 and errs["sm_w"] < 5e-3
                                and err < 5e-2 and max(nerrs.values()) < 3e-2}
            else:
                st = eng.staleness()
                moved = max(float((final[k] - init[k]).abs().max()) for k in init)
                report[mode] = {"global_step": int(final["global_step"]), "staleness": st, "moved": moved,
                                "ok": int(final["global_step"]) == steps * W and st["count"] == steps * W and moved > 0}
        dist.barrier()
        eng.close()
        dist.barrier()
    if rank == 0:
        report["world"] = world
        report["num_ps"] = num_ps
        report["nvls"] = str(nvls)
        report["ps_on_workers"] = pow_
        report["precision"] = PRECISION
        print("MP_CHECK " + json.dumps(report))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "mp_check_%d%s%s_%s.json" % (world, "_nvls" if nvls else "", "_pow" if pow_ else "",
                                                                               PRECISION)), "w") as f:
            json.dump(report, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
