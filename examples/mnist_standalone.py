"""BASELINE.json config 1: the MNIST MLP in ONE process on the CPU (world_size = 1) -- the plumbing check that runs
anywhere, no cluster, no GPU.  Same model, loss and Adam as ``distributed_mnist.py`` (reference
``distributed_mnist.py:96-126``), driven by ``MonitoredTrainingSession`` with an in-process master (``master=""``),
a ``StopAtStepHook`` and checkpoints; prints steps/s and the validation cross-entropy.

    python examples/mnist_standalone.py --train_steps 500
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

import time

from _common import dtf
from distributed_tensorflow_b200 import input_data
from distributed_tensorflow_b200.models import build_mnist_mlp

F = dtf.app.flags
F.DEFINE_integer("train_steps", 500, "global steps to run")
F.DEFINE_integer("batch_size", 100, "examples per step")
F.DEFINE_integer("hidden_units", 100, "hidden width")
F.DEFINE_float("learning_rate", 0.01, "Adam step size")
F.DEFINE_integer("num_train", 5000, "synthetic training-split size")
F.DEFINE_string("train_dir", "", "checkpoint directory ('' = no checkpoints)")
F.DEFINE_string("device", "/cpu:0", "'/cpu:0' or '/gpu:0'")
FLAGS = F.FLAGS


def main():
    data = input_data.read_data_sets(None, one_hot=True, num_train=FLAGS.num_train)
    with dtf.device(FLAGS.device):
        net = build_mnist_mlp(hidden=FLAGS.hidden_units, fused=True)
        train_op = dtf.train.AdamOptimizer(FLAGS.learning_rate).minimize(net["loss"], global_step=net["global_step"])
    hooks = [dtf.train.StopAtStepHook(last_step=FLAGS.train_steps)]
    t0, steps, loss = time.time(), 0, float("nan")
    with dtf.train.MonitoredTrainingSession(master="", is_chief=True, checkpoint_dir=FLAGS.train_dir or None, hooks=hooks) as sess:
        while not sess.should_stop():
            xs, ys = data.train.next_batch(FLAGS.batch_size)
            _, loss = sess.run([train_op, net["loss"]], feed_dict={net["x"]: xs, net["y_"]: ys})
            steps += 1
        dt = time.time() - t0
    print("standalone: %d steps in %.2fs = %.0f steps/s (%.0f samples/s), last batch loss %.4f"
          % (steps, dt, steps / dt, steps * FLAGS.batch_size / dt, loss))


if __name__ == "__main__":
    main()
