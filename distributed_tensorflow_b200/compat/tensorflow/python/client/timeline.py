"""``from tensorflow.python.client import timeline`` (reference ``example_in_graph.py:10,65``)."""
from distributed_tensorflow_b200.utils.timeline import *  # noqa: F401,F403
from distributed_tensorflow_b200.utils.timeline import Timeline  # noqa: F401
