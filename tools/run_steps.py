"""Plain step loop on one GPU (target for ncu).  usage: python tools/run_steps.py [steps] [optimizer]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = sys.argv[2] if len(sys.argv) > 2 else "sgd"
torch.cuda.set_device(0)
xs, ys = synthetic_mnist(5000, seed=1)
eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": kind, "lr": 0.001}), Fabric(1, {0: 0}))
eng.init_params()
eng.attach_dataset(0, xs, ys)
eng.enqueue_local_steps(steps, "dataset")
eng.synchronize()
eng.check_errors()
print("loss", eng.read_loss(), "global_step", eng.read_ctl(0, "global_step"))
eng.close()
