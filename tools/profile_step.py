"""In-situ kernel timeline of the engine step loop (CUPTI via torch.profiler): per-kernel duration and the gap
to the previous kernel, for eager launches and for CUDA-graph replay.  Writes gpurun_out/step_timeline.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402


def summarize(prof, label):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name
           and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    rows = []
    prev_end = None
    for e in evs:
        s, t = e.time_range.start, e.time_range.end
        rows.append((e.name.split("(")[0][:40], t - s, (s - prev_end) if prev_end is not None else 0.0))
        prev_end = t
    agg = {}
    for name, dur, gap in rows[len(rows) // 4:]:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += dur
        a[2] += gap
    out = {k: {"n": v[0], "dur_us": round(v[1] / v[0], 2), "gap_before_us": round(v[2] / v[0], 2)} for k, v in agg.items()}
    print(label, json.dumps(out, indent=1))
    return out


def main():
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(20000, seed=1)
    eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.001}), Fabric(1, {0: 0}))
    eng.init_params()
    eng.attach_dataset(0, xs, ys)
    eng.enqueue_local_steps(10, "dataset")
    eng.synchronize()
    res = {}
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        eng.enqueue_local_steps(40, "dataset")
        eng.synchronize()
    res["eager"] = summarize(prof, "EAGER")
    eng.capture_graphs(20, "dataset")
    eng.replay_graphs(1)
    eng.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        eng.replay_graphs(3)
        eng.synchronize()
    res["graph"] = summarize(prof, "GRAPH")
    eng.check_errors()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_timeline.json"), "w") as f:
        json.dump(res, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
