"""The same MNIST ps/worker job as ``distributed_mnist.py``, driven the pre-MonitoredTrainingSession way: with
``tf.train.Supervisor``.

The reference script is derived from TensorFlow r1.3's ``mnist_replica.py`` (``distributed_mnist.py:57``), and that
generation of parameter-server programs looks like this: ``sv = Supervisor(is_chief, logdir, init_op, ...)``,
``sess = sv.prepare_or_wait_for_session(server.target)``; with ``--sync_replicas`` the chief runs the optimizer's
init-tokens op and starts its queue runner through ``sv.start_queue_runners``; every worker loops on ``sess.run`` until
the shared global step reaches ``--train_steps``.

    python examples/launch_local.py examples/mnist_supervisor.py --num_ps 1 --num_workers 2 -- --sync_replicas=True --train_steps=300
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

import time

from _common import bring_up, define_cluster_flags, dtf
from distributed_tensorflow_b200 import input_data
from distributed_tensorflow_b200.models import build_mnist_mlp

F = dtf.app.flags
FLAGS = define_cluster_flags("127.0.0.1:22231", "127.0.0.1:22232,127.0.0.1:22233")
F.DEFINE_integer("train_steps", 300, "stop when the shared global step reaches this")
F.DEFINE_integer("batch_size", 100, "examples per worker step")
F.DEFINE_integer("hidden_units", 100, "hidden width")
F.DEFINE_float("learning_rate", 0.01, "Adam step size")
F.DEFINE_bool("sync_replicas", False, "aggregate gradients with SyncReplicasOptimizer before applying them")
F.DEFINE_integer("replicas_to_aggregate", 0, "gradients per update (0 = number of workers)")
F.DEFINE_string("train_dir", "/tmp/dtf_ckpt/mnist_sv", "Supervisor logdir (checkpoints)")
F.DEFINE_integer("num_train", 5000, "synthetic training-split size")


def main():
    cluster, server, num_workers = bring_up(FLAGS)
    chief = FLAGS.task_index == 0
    data = input_data.read_data_sets(None, one_hot=True, num_train=FLAGS.num_train)
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:%d/cpu:0" % FLAGS.task_index)):
        net = build_mnist_mlp(hidden=FLAGS.hidden_units)
        opt = dtf.train.AdamOptimizer(FLAGS.learning_rate)
        sv_args = {}
        if FLAGS.sync_replicas:
            n = FLAGS.replicas_to_aggregate or num_workers
            opt = dtf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=n, total_num_replicas=num_workers, name="mnist_sync_replicas")
        train_op = opt.minimize(net["loss"], global_step=net["global_step"])
        if FLAGS.sync_replicas:
            sv_args = {"local_init_op": opt.chief_init_op if chief else opt.local_step_init_op,
                       "ready_for_local_init_op": opt.ready_for_local_init_op}
        init_op = dtf.global_variables_initializer()
    sv = dtf.train.Supervisor(is_chief=chief, logdir=FLAGS.train_dir, init_op=init_op, recovery_wait_secs=1,
                              global_step=net["global_step"], save_model_secs=30, **sv_args)
    print("Worker %d: %s" % (FLAGS.task_index, "Initializing session..." if chief else "Waiting for session to be initialized..."))
    sess = sv.prepare_or_wait_for_session(server.target)
    print("Worker %d: Session initialization complete." % FLAGS.task_index)
    if FLAGS.sync_replicas and chief:
        sess.run(opt.get_init_tokens_op())                         # lets the first step through
        sv.start_queue_runners(sess, [opt.get_chief_queue_runner()])
    t0, local_step, step = time.time(), 0, 0
    # sync mode: replicas can be a step apart, so only the chief decides when training ends -- the others keep feeding
    # gradients until its sv.stop() closes the token queue (OutOfRangeError below); stopping them on their own view of
    # global_step could leave the chief waiting for a gradient nobody sends
    follower = FLAGS.sync_replicas and not chief
    while not sv.should_stop() and (follower or step < FLAGS.train_steps):
        xs, ys = data.train.next_batch(FLAGS.batch_size)
        try:
            _, step = sess.run([train_op, net["global_step"]], feed_dict={net["x"]: xs, net["y_"]: ys})
        except dtf.errors.OutOfRangeError:                         # the chief finished and closed the token queue
            break
        local_step += 1
        if local_step % 50 == 0:
            print("%f: Worker %d: training step %d done (global step: %d)" % (time.time(), FLAGS.task_index, local_step, step))
    print("Training elapsed time: %f s" % (time.time() - t0))
    val = sess.run(net["loss"], feed_dict={net["x"]: data.validation.images, net["y_"]: data.validation.labels})
    print("After %d training step(s), validation cross entropy = %g" % (FLAGS.train_steps, val))
    sv.stop()
    server.stop()


if __name__ == "__main__":
    main()
