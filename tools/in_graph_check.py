"""In-graph replication on the fabric (BASELINE.json config 3): ONE process drives `--gpus` devices,
`--num-ps` ps shards (variables placed round-robin) + the rest workers, Adam, bf16 compute.
Checks the loss goes down, global_step == steps, and both shards applied every aggregate."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--num-ps", type=int, default=2)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
N = a.gpus
W = N - a.num_ps
xs, ys = synthetic_mnist(100 * W * 20, seed=4)
cfg = EngineConfig(num_ps=a.num_ps, num_workers=W, sync=True, optimizer={"kind": "adam", "lr": 0.005}, seed=5)
eng = PSTrainEngine(MLPSpec(), cfg, Fabric(N, {r: r for r in range(N)}))
eng.init_params()
for r in eng.worker_ranks:
    eng.attach_dataset(r, xs, ys)
eng.enqueue_local_steps(1, "dataset")
l0 = eng.read_loss()
eng.enqueue_local_steps(a.steps - 1, "dataset")
eng.synchronize()
eng.check_errors()
l1 = eng.read_loss()
sd = eng.state_dict()
out = {"gpus": N, "num_ps": a.num_ps, "workers": W, "first_loss": l0, "last_loss": l1,
       "global_step": int(sd["global_step"]), "shard_versions": [eng.read_ctl(s, "param_version") for s in range(a.num_ps)],
       "placement": {k: v.shard for k, v in eng.layout.items()},
       "ok": l1 < 0.5 * l0 and int(sd["global_step"]) == a.steps}
print("IN_GRAPH_CHECK " + json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "in_graph_check.json"), "w"), indent=1)
eng.close()
