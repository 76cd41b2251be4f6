"""Driver for `ncu -k regex:mlp_step|ps_apply`: a few eager steps of the default (tf32, one-kernel-step) engine on one GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_tensorflow_b200.parallel.fabric import Fabric  # noqa: E402
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine  # noqa: E402
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist  # noqa: E402

torch.cuda.set_device(0)
xs, ys = synthetic_mnist(20000, seed=1)
eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.001}), Fabric(1, {0: 0}))
eng.init_params()
eng.attach_dataset(0, xs, ys)
eng.enqueue_local_steps(int(sys.argv[1]) if len(sys.argv) > 1 else 12, "dataset")
eng.synchronize()
eng.check_errors()
eng.close()
