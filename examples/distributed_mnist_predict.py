"""Single-process inference from a training checkpoint (reference ``distributed_mnist_predict.py``, S14).

Rebuilds the four MLP variables **by name** (``hid_w, hid_b, sm_w, sm_b``), restores just
those from the newest checkpoint in ``--checkpoint_dir`` (the file also holds ``global_step``,
Adam slots and beta powers: partial, name-keyed restore), and counts correct predictions on
the validation split.  Unlike the reference it stops with a clear message when there is no
checkpoint instead of crashing in ``restore``.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200 import input_data

dtf.app.flags.DEFINE_string("data_dir", "/tmp/mnist-data", "MNIST IDX directory (synthetic if absent)")
dtf.app.flags.DEFINE_string("checkpoint_dir", "/tmp/dtf_ckpt/mnist", "directory written by distributed_mnist.py")
dtf.app.flags.DEFINE_integer("hidden_units", 100, "hidden layer width used in training")
FLAGS = dtf.app.flags.FLAGS
IMAGE_PIXELS = 28


def main():
    mnist = input_data.read_data_sets(FLAGS.data_dir, one_hot=True, num_train=1000)
    print("len of validation images: ", len(mnist.validation.images))
    H = FLAGS.hidden_units
    hid_w = dtf.Variable(dtf.truncated_normal([IMAGE_PIXELS * IMAGE_PIXELS, H], stddev=1.0 / IMAGE_PIXELS), name='hid_w')
    hid_b = dtf.Variable(dtf.zeros([H]), name='hid_b')
    sm_w = dtf.Variable(dtf.truncated_normal([H, 10], stddev=1.0 / math.sqrt(H)), name='sm_w')
    sm_b = dtf.Variable(dtf.zeros([10]), name='sm_b')
    x = dtf.placeholder(dtf.float32, [None, IMAGE_PIXELS * IMAGE_PIXELS])
    hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    y = dtf.nn.softmax(dtf.nn.xw_plus_b(hid, sm_w, sm_b))
    pre = dtf.arg_max(y, dimension=1)

    with dtf.Session() as sess:
        restorer = dtf.train.Saver()
        check_point = dtf.train.get_checkpoint_state(FLAGS.checkpoint_dir)
        if not check_point:
            print("ckpt is none.")
            return 1
        restorer.restore(sess, check_point.model_checkpoint_path)
        pre_ = sess.run(pre, feed_dict={x: mnist.validation.images})
        print("predict: ", len(pre_))
        correct = int(np.sum(pre_ == np.argmax(mnist.validation.labels, axis=1)))
        print(correct)
        print("accuracy: %.4f" % (correct / float(len(pre_))))
    return 0


if __name__ == "__main__":
    sys.exit(main())
