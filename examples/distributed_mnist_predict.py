"""Inference from a training checkpoint, in one process (what the reference's ``distributed_mnist_predict.py`` does, S14).

The model is rebuilt with the SAME variable names the trainer used (``build_mnist_mlp``: ``hid_w, hid_b, sm_w,
sm_b``); ``Saver.restore`` is name-keyed and partial, so the optimizer slots, beta powers and ``global_step`` that
are also in the file are simply not read (the topology that wrote the checkpoint -- n ps tasks, m workers -- does
not matter).  Prints the number of predictions, the number of correct ones and the accuracy on the validation split.
With no checkpoint in ``--checkpoint_dir`` it says so and exits 1 (the reference would crash inside ``restore``).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

import numpy as np

from _common import dtf
from distributed_tensorflow_b200 import input_data
from distributed_tensorflow_b200.models import build_mnist_mlp

dtf.app.flags.DEFINE_string("data_dir", "/tmp/mnist-data", "MNIST IDX directory (synthetic split if absent)")
dtf.app.flags.DEFINE_string("checkpoint_dir", "/tmp/dtf_ckpt/mnist", "directory distributed_mnist.py checkpointed into")
dtf.app.flags.DEFINE_integer("hidden_units", 100, "hidden width the checkpoint was trained with")
FLAGS = dtf.app.flags.FLAGS


def main():
    val = input_data.read_data_sets(FLAGS.data_dir, one_hot=True, num_train=1000).validation
    print("len of validation images: ", len(val.images))
    net = build_mnist_mlp(hidden=FLAGS.hidden_units)
    predicted_class = dtf.arg_max(net["y"], dimension=1)
    state = dtf.train.get_checkpoint_state(FLAGS.checkpoint_dir)
    if not state:
        print("ckpt is none.")
        return 1
    with dtf.Session() as sess:
        dtf.train.Saver(var_list=list(net["vars"])).restore(sess, state.model_checkpoint_path)
        guess = sess.run(predicted_class, feed_dict={net["x"]: val.images})
    hits = int(np.sum(guess == np.argmax(val.labels, axis=1)))
    print("predict: ", len(guess))
    print(hits)
    print("accuracy: %.4f" % (hits / float(len(guess))))
    return 0


if __name__ == "__main__":
    sys.exit(main())
