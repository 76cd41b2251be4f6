"""``from tensorflow.examples.tutorials.mnist import input_data`` (reference ``distributed_mnist.py:12,81``)."""
from distributed_tensorflow_b200.utils.mnist_data import read_data_sets  # noqa: F401
