"""The MNIST MLP fed from TFRecord files instead of ``feed_dict`` -- the input path of TF-1.x programs whose data set was converted
to records (the reference feeds ``mnist.train.next_batch`` through placeholders, ``/root/reference/distributed_mnist.py:149-152``).

    python examples/mnist_tfrecords.py --data_dir /tmp/mnist_records --train_steps 300

Step 1 (once): the training split is written as ``tf.train.Example`` records (``image_raw``: the fp32 pixels as bytes, ``label``:
int64) into ``--shards`` TFRecord files.  Step 2: ``TFRecordDataset -> shuffle -> repeat -> batch -> map(parse_batch) -> prefetch ->
get_next`` feeds the same model / clipped batch-sum cross-entropy / Adam as ``distributed_mnist.py`` under a
``MonitoredTrainingSession``; prints steps/s and the accuracy on held-out records.  Files are mapped and checksum-verified by the
native scanner (``csrc/runtime/bundle_io.cpp: dtf_tfrecord_scan``), batches are parsed by ``csrc/runtime/example_parser.cpp``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

import time

import numpy as np

from _common import dtf
from distributed_tensorflow_b200 import input_data

F = dtf.app.flags
F.DEFINE_string("data_dir", "/tmp/dtf_mnist_records", "where the TFRecord shards live (written on first use)")
F.DEFINE_integer("shards", 4, "TFRecord files of the training split")
F.DEFINE_integer("num_train", 4000, "synthetic training-split size")
F.DEFINE_integer("train_steps", 300, "global steps to run")
F.DEFINE_integer("batch_size", 100, "examples per step")
F.DEFINE_integer("hidden_units", 100, "hidden width")
F.DEFINE_float("learning_rate", 0.01, "Adam step size")
FLAGS = F.FLAGS


def _write(path, images, labels):
    with dtf.python_io.TFRecordWriter(path) as w:
        for img, lab in zip(images, labels):
            w.write(dtf.train.Example(features=dtf.train.Features(feature={
                "image_raw": dtf.train.Feature(bytes_list=dtf.train.BytesList(value=[np.ascontiguousarray(img, np.float32).tobytes()])),
                "label": dtf.train.Feature(int64_list=dtf.train.Int64List(value=[int(lab)]))})).SerializeToString())


def convert(data_dir: str):
    os.makedirs(data_dir, exist_ok=True)
    train = [os.path.join(data_dir, "train-%05d-of-%05d.tfrecord" % (k, FLAGS.shards)) for k in range(FLAGS.shards)]
    test = os.path.join(data_dir, "validation.tfrecord")
    marker = os.path.join(data_dir, "DONE-%d" % FLAGS.num_train)
    if not os.path.exists(marker):
        data = input_data.read_data_sets(None, one_hot=False, num_train=FLAGS.num_train)
        for k, path in enumerate(train):
            _write(path, data.train.images[k::FLAGS.shards], data.train.labels[k::FLAGS.shards])
        _write(test, data.validation.images[:1000], data.validation.labels[:1000])
        open(marker, "w").close()
    return train, test


SPEC = {"image_raw": dtf.FixedLenFeature([], dtf.string), "label": dtf.FixedLenFeature([], dtf.int64)}


def parse_batch(records):
    """A whole batch of serialized Examples at once (``batch`` BEFORE ``map``, the fast spelling in TensorFlow too): the native batch
    parser fills the label array and locates the image bytes, ``decode_raw`` views them as ``[batch, 784]`` floats."""
    d = dtf.parse_example(records, SPEC)
    return dtf.decode_raw(d["image_raw"], dtf.float32), np.eye(10, dtype=np.float32)[d["label"]]


def main():
    train_files, test_file = convert(FLAGS.data_dir)
    ds = dtf.data.TFRecordDataset(train_files).shuffle(2000, seed=0).repeat().batch(FLAGS.batch_size).map(parse_batch).prefetch(2)
    x, y_ = ds.make_one_shot_iterator().get_next()
    global_step = dtf.train.get_or_create_global_step()
    h = FLAGS.hidden_units
    hid_w = dtf.Variable(dtf.truncated_normal([784, h], stddev=1.0 / 28.0), name="hid_w")
    hid_b = dtf.Variable(dtf.zeros([h]), name="hid_b")
    sm_w = dtf.Variable(dtf.truncated_normal([h, 10], stddev=1.0 / np.sqrt(h)), name="sm_w")
    sm_b = dtf.Variable(dtf.zeros([10]), name="sm_b")

    def model(inp):
        return dtf.nn.xw_plus_b(dtf.nn.relu(dtf.nn.xw_plus_b(inp, hid_w, hid_b)), sm_w, sm_b)
    loss = -dtf.reduce_sum(y_ * dtf.log(dtf.clip_by_value(dtf.nn.softmax(model(x)), 1e-10, 1.0)))
    train_op = dtf.train.AdamOptimizer(FLAGS.learning_rate).minimize(loss, global_step=global_step)
    # held-out records, parsed eagerly (a python-side use of the same reader)
    vx, vy = parse_batch(list(dtf.python_io.tf_record_iterator(test_file)))
    px = dtf.placeholder(dtf.float32, [None, 784])
    pred = dtf.argmax(model(px), 1)
    t0, steps, last = time.time(), 0, float("nan")
    with dtf.train.MonitoredTrainingSession(master="", is_chief=True, hooks=[dtf.train.StopAtStepHook(last_step=FLAGS.train_steps)]) as sess:
        while not sess.should_stop():
            _, last = sess.run([train_op, loss])
            steps += 1
            if steps == FLAGS.train_steps:                       # (the session is still open: evaluate before the stop)
                acc = float((sess.run(pred, feed_dict={px: vx}) == vy.argmax(1)).mean())
        dt = time.time() - t0
    print("tfrecords: %d steps in %.2fs = %.0f steps/s, last batch loss %.4f, held-out accuracy %.3f (%d records in %d shards)"
          % (steps, dt, steps / dt, last, acc, FLAGS.num_train, FLAGS.shards))


if __name__ == "__main__":
    main()
