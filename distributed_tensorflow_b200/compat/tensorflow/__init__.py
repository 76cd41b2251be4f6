"""``import tensorflow as tf`` -> distributed_tensorflow_b200 (see ``distributed_tensorflow_b200/compat/__init__.py``)."""
import distributed_tensorflow_b200 as _dtf
from distributed_tensorflow_b200 import *  # noqa: F401,F403

__version__ = "1.12.0-dtf_b200"


def __getattr__(name):
    return getattr(_dtf, name)
