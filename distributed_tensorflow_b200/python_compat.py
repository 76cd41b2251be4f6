"""Compatibility namespaces for import paths the reference uses.

``from tensorflow.python.client import timeline`` (reference
``example_in_graph.py:10``) -> ``from distributed_tensorflow_b200 import timeline``;
``from tensorflow.examples.tutorials.mnist import input_data`` (reference
``distributed_mnist.py:12``) -> ``from distributed_tensorflow_b200 import input_data``.
"""
import types as _types

from .utils.timeline import Timeline
from .utils import mnist_data as input_data  # noqa: F401

timeline = _types.SimpleNamespace(Timeline=Timeline)
