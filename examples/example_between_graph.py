"""Between-graph replication on the smallest possible model: fit ``y = 2x + 10`` with SGD on a ps cluster.

Counterpart of the reference's ``example_between_graph.py`` (S9/S10).  Every worker builds the same graph; the two
scalars ``weight`` and ``biase`` and the shared ``global_step`` live on the ps (``replica_device_setter``).  Async:
each worker's gradient is applied as it arrives, steps interleave.  ``--is_sync``: gradients of
``replicas_to_aggregate`` workers are averaged per update (``SyncReplicasOptimizer`` + its chief hook).  The chief
checkpoints into ``--ckpt_dir`` every ``--save_checkpoint_secs``; restart the job and it resumes from there.
``MonitoredTrainingSession.run`` transparently recovers when a ps is restarted (see ``test_recoverable_session``).

    python examples/launch_local.py examples/example_between_graph.py --num_ps 1 --num_workers 2 -- --is_sync=True
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py
from datetime import datetime

import numpy as np

from _common import bring_up, define_cluster_flags, dtf

F = dtf.app.flags
FLAGS = define_cluster_flags("127.0.0.1:2222", "127.0.0.1:2223,127.0.0.1:2224")
F.DEFINE_float("learning_rate", 0.03, "SGD step size")
F.DEFINE_integer("steps_to_validate", 1, "print the fitted line every N global steps")
F.DEFINE_bool("is_sync", False, "synchronous replicas")
F.DEFINE_integer("num_workers", 0, "total_num_replicas in sync mode (0 = all workers of the cluster)")
F.DEFINE_integer("replicas_to_aggregate", 0, "gradients averaged per update (0 = num_workers); fewer than num_workers "
                                             "= backup workers: the slowest / dead replicas are not waited for")
F.DEFINE_integer("num_steps", 2000, "global steps to run, counted from the step found at session creation")
F.DEFINE_string("ckpt_dir", "/tmp/dtf_ckpt/linear", "checkpoint directory every task can reach")
F.DEFINE_integer("save_checkpoint_secs", 60, "seconds between checkpoints")


def linear_model():
    """pred = x * weight + biase, mean squared error; variables created in this order -> ps placement round-robin."""
    step = dtf.Variable(0, name="global_step", trainable=False, dtype=dtf.int64)
    x, target = dtf.placeholder(dtf.float32), dtf.placeholder(dtf.float32)
    w = dtf.get_variable("weight", [1], dtf.float32, initializer=dtf.random_normal_initializer())
    b = dtf.get_variable("biase", [1], dtf.float32, initializer=dtf.random_normal_initializer())
    mse = dtf.reduce_mean(dtf.square(target - (dtf.multiply(x, w) + b)))
    return step, x, target, w, b, mse


def main():
    cluster, server, n_workers = bring_up(FLAGS)
    xs = np.random.rand(100).astype(np.float32)
    ys = 2.0 * xs + 10.0
    chief = FLAGS.task_index == 0
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:%d" % FLAGS.task_index)):
        step, x, target, w, b, mse = linear_model()
        sgd = dtf.train.GradientDescentOptimizer(FLAGS.learning_rate)
        hooks = [dtf.train.StopAtStepHook(num_steps=FLAGS.num_steps)]
        if FLAGS.is_sync:
            n = FLAGS.num_workers or n_workers
            sgd = dtf.train.SyncReplicasOptimizer(sgd, replicas_to_aggregate=FLAGS.replicas_to_aggregate or n,
                                                  total_num_replicas=n)
            hooks.append(sgd.make_session_run_hook(chief))
        update = sgd.minimize(mse, global_step=step)
    session_config = dtf.ConfigProto(gpu_options=dtf.GPUOptions(per_process_gpu_memory_fraction=0.1))
    with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=chief, checkpoint_dir=FLAGS.ckpt_dir, hooks=hooks,
                                            save_checkpoint_secs=FLAGS.save_checkpoint_secs, config=session_config) as sess:
        while not sess.should_stop():
            _, err, now_at = sess.run([update, mse, step], feed_dict={x: xs, target: ys})
            if now_at % FLAGS.steps_to_validate == 0 and not sess.should_stop():
                wv, bv = sess.run([w, b])                 # a second pull of the two scalars, just to show them
                print("time: %s, step: %d, weight: %f, biase: %f, loss: %f" % (datetime.now(), now_at, wv[0], bv[0], err))
    server.stop()


if __name__ == "__main__":
    main()
