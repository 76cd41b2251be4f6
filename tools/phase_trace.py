"""Phase-level timing of the two step GEMMs from in-kernel clock64 stamps (gpurun_out/phase_trace.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402

NAMES = ["start", "setup_done", "token_acquired", "last_tma_issued", "first_stage_landed", "last_mma_issued",
         "accum_ready", "tile_stored", "signalled", "end"]
torch.cuda.set_device(0)
xs, ys = synthetic_mnist(5000, seed=1)
eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.001}), Fabric(1, {0: 0}))
eng.init_params()
eng.attach_dataset(0, xs, ys)
eng.enqueue_local_steps(10, "dataset")
eng.synchronize()
d = eng._w[0]
t1 = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
t3 = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
d["g1"].phase_trace = t1.data_ptr()
d["g3"].phase_trace = t3.data_ptr()
th = torch.zeros(16, dtype=torch.int64, device="cuda")
tp = torch.zeros(16, dtype=torch.int64, device="cuda")
d["head"].phase_trace = th.data_ptr()
eng._p[0].phase_trace = tp.data_ptr()
eng.enqueue_local_steps(5, "dataset")
eng.synchronize()
out = {}
for name, t, nc in (("F1", t1, 1), ("B3", t3, 7)):
    v = t.view(8, 16).cpu()
    rows = []
    for c in range(nc):
        base = int(v[c, 0])
        rows.append({NAMES[i]: int(v[c, i]) - base for i in range(10) if int(v[c, i])})
    out[name] = rows
    print(name, json.dumps(rows[0]))
    if nc > 1:
        print(name, "cta%d" % (nc - 1), json.dumps(rows[-1]))
hv = th.cpu().tolist()
HN = ["start", "inputs_in_smem", "logits", "softmax_dlogits", "dw2_dh_db1", "grads_stored", "fenced_signalled"]
out["head"] = {HN[i]: hv[i] - hv[0] for i in range(7)}
print("head", json.dumps(out["head"]))
pv = tp.cpu().tolist()
out["ps_apply"] = {"decision_known": pv[1] - pv[0], "block0_slice_done": pv[2] - pv[0], "block0_end": pv[5] - pv[0],
                   "last_block": pv[4], "last_block_done_clock_minus_block0_start": pv[3] - pv[0]}
print("ps_apply", json.dumps(out["ps_apply"]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "phase_trace.json"), "w"), indent=1)
eng.close()
