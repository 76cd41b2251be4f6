"""``import tensorflow as tf`` -> distributed_tensorflow_b200 (see ``distributed_tensorflow_b200/compat/__init__.py``)."""
import distributed_tensorflow_b200 as _dtf
from distributed_tensorflow_b200 import *  # noqa: F401,F403

__version__ = "1.12.0-dtf_b200"

# The reference scripts hide the GPUs right after importing tensorflow (``os.environ["CUDA_VISIBLE_DEVICES"] = "-1"``,
# distributed_mnist.py:15).  A task the launcher bound to a B200 (DTF_GPU_INDEX) initialises CUDA HERE, at import time, so
# that later assignment has no effect on this process and the unmodified script trains on the fabric.
import os as _os
if _os.environ.get("DTF_GPU_INDEX", "") not in ("", "-1") and _os.environ.get("DTF_FABRIC", "auto") != "0":
    import torch as _torch
    if _torch.cuda.is_available():
        _torch.cuda.init()


def __getattr__(name):
    return getattr(_dtf, name)
