"""Where one training step goes on ONE GPU (ps + worker colocated), default engine (precision tf32: one step kernel + one
ps kernel per step), measured IN SITU inside a CUDA-graph replay: %globaltimer stamps written by every CTA of the step
kernel (csrc/mlp_step.cu STAMP slots) and by ps_apply (entry / tokens released), all on one time axis.

    python tools/step_trace.py        -> gpurun_out/step_trace.json  (+ a table on stdout)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402

SLOTS = ["entry", "setup_done(barriers,TMEM)", "token_acquired+W1_TMA_issued", "F1_accum_ready", "partials_stored+flag",
         "all_partials_visible", "head_done+dh_flag", "B3_accum_ready", "dW1_stored", "fenced+arrived", "exit",
         "t0:x_landed", "t0:W1_landed", "t0:F1_MMAs_committed", "t0:x_refetch_issued", "h_finalised", "softmax_dlogits_done",
         "dh_rows_done", "dW2_db2_atomics_issued", "db1_issued", "t0:dh_flag_seen+TMA_issued", "t0:dh_and_x2_landed"]
ORDER = [0, 1, 2, 11, 12, 13, 14, 3, 4, 5, 15, 16, 17, 18, 19, 6, 20, 21, 7, 8, 9, 10]


def main():
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(20000, seed=1)
    eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.001},
                                           step_ctas=int(os.environ.get("DTF_STEP_CTAS", "0"))), Fabric(1, {0: 0}))
    eng.init_params()
    eng.attach_dataset(0, xs, ys)
    d = eng._w[0]
    G = eng.step_ctas
    tr = eng.ranks[0].bufs["steptrace_w0"]
    d["step_ds"].trace = tr.ptr
    eng.enqueue_local_steps(10, "dataset")
    eng.synchronize()
    eng.capture_graphs(20, "dataset")
    eng.replay_graphs(3)
    eng.synchronize()
    eng.check_errors()
    t = tr.tensor(torch.int64, 0, 16 * 32).view(16, 32)[:G].cpu()
    ring = eng.ranks[0].bufs["trace0"].tensor(torch.int64, 0, eng.cfg.trace_cap * 4).view(-1, 4).cpu()
    gs = int(eng.read_ctl(0, "global_step"))
    # the ps_apply launches around the LAST worker step: seq = gs - 1 (previous) and gs (the one after the last worker step)
    rows = {int(r[3]): (int(r[1]), int(r[2])) for r in ring.tolist() if r[0] == 1}
    t0 = int(t[:, 0].min())
    out = {"ctas": G, "slots": SLOTS, "order": ORDER,
           "per_cta_us": [[round((int(v) - t0) / 1e3, 2) if int(v) else None for v in row[:len(SLOTS)]] for row in t],
           "ps_apply_prev_us": [round((x - t0) / 1e3, 2) for x in rows.get(gs - 1, (0, 0))],
           "ps_apply_next_us": [round((x - t0) / 1e3, 2) for x in rows.get(gs, (0, 0))]}
    med = [sorted(c[i] for c in out["per_cta_us"] if c[i] is not None)[G // 2] for i in range(len(SLOTS))]
    out["median_us"] = dict(zip(SLOTS, med))
    print("step kernel (us since the first CTA's entry; median over %d CTAs; t0: = stamped by thread 0's serial path):" % G)
    prev = 0.0
    for i in ORDER:
        print("  %-34s %7.2f  (+%.2f)" % (SLOTS[i], med[i], med[i] - prev))
        prev = med[i]
    print("previous ps_apply [entry, tokens released]:", out["ps_apply_prev_us"])
    print("next     ps_apply [entry, tokens released]:", out["ps_apply_next_us"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "step_trace.json"), "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
