"""Drop-in compatibility shim: lets the reference's UNMODIFIED TensorFlow-1.x scripts run on this framework.

``compat/tensorflow`` is a package named ``tensorflow`` whose attributes forward to
``distributed_tensorflow_b200`` (``tf.train.Server``, ``tf.app.flags``, ``tf.nn.xw_plus_b`` ...), plus the two
helper modules the reference imports by path (``tensorflow.examples.tutorials.mnist.input_data``,
``tensorflow.python.client.timeline``).  Put this directory first on ``sys.path``:

    python -m distributed_tensorflow_b200.compat.run /path/to/distributed_mnist.py --job_name=ps --task_index=0 ...

(reference scripts: ``distributed_mnist.py:10-12``, ``example_in_graph.py:7-10``).  It exists to demonstrate and
test API parity; new programs should ``import distributed_tensorflow_b200 as dtf``.
"""
import os

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))
