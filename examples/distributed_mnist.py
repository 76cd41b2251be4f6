"""MNIST MLP on a parameter-server cluster: between-graph replication, asynchronous or synchronous replicas.

What the reference's ``distributed_mnist.py`` does (S1-S8), written against this framework's helpers: the model
comes from ``dtf.models.build_mnist_mlp`` (784 -> hidden ReLU -> 10, clipped batch-SUM cross-entropy, variables
``hid_w, hid_b, sm_w, sm_b`` placed round-robin on the ps tasks by ``replica_device_setter``), Adam, optional
``SyncReplicasOptimizer``, ``MonitoredTrainingSession`` (chief initialises / restores and checkpoints, the others
wait), a stop hook on the shared global step, validation every ``--validate_every`` global steps.

    python examples/launch_local.py examples/distributed_mnist.py --num_ps 1 --num_workers 2 -- --issync=True --train_steps=2000

or one process per task by hand (``--job_name=ps|worker --task_index=N --ps_hosts=... --worker_hosts=...``).
``--engine=fabric`` keeps the same program but moves pull / push / aggregation / tokens onto the GPUs (NVLink).
Unlike the reference: ``--train_steps`` is honoured, the checkpoint directory is a flag, GPUs are used when visible.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py
import time

from datetime import datetime

from _common import bring_up, define_cluster_flags, dtf
from distributed_tensorflow_b200 import input_data
from distributed_tensorflow_b200.models import build_mnist_mlp

flags = dtf.app.flags
FLAGS = define_cluster_flags("127.0.0.1:22221", "127.0.0.1:22222,127.0.0.1:22223")
flags.DEFINE_string("data_dir", "/tmp/mnist-data", "MNIST IDX files; a synthetic MNIST-shaped split is used if absent")
flags.DEFINE_integer("hidden_units", 100, "width of the hidden layer")
flags.DEFINE_integer("train_steps", 10000, "global step at which every worker stops")
flags.DEFINE_integer("batch_size", 100, "examples per worker step")
flags.DEFINE_float("learning_rate", 0.01, "Adam step size")
flags.DEFINE_bool("issync", False, "aggregate the workers' gradients (SyncReplicasOptimizer) instead of applying each alone")
flags.DEFINE_string("train_dir", "/tmp/dtf_ckpt/mnist", "checkpoint directory on a filesystem every task can reach")
flags.DEFINE_integer("validate_every", 1000, "validate when global_step + 1 is a multiple of this")
flags.DEFINE_integer("num_train", 55000, "size of the synthetic training split")
flags.DEFINE_integer("log_every", 1, "print the step line every N local steps")
flags.DEFINE_bool("measure_staleness", False, "async mode: histogram of (global step at apply) - (global step at pull)")
flags.DEFINE_string("engine", "graph", "'graph' = control-plane tier over RPC, 'fabric' = NVLink peer-memory tier")


class AnnouncedStop(dtf.train.StopAtStepHook):
    """``StopAtStepHook`` that says where it will stop and when it did (cf. reference ``distributed_mnist.py:41-54``)."""

    def after_create_session(self, session, coord):
        super().after_create_session(session, coord)
        print("stop hook armed: last_step=%s" % self._last_step, flush=True)

    def after_run(self, run_context, run_values):
        if run_values.results >= self._last_step:
            print("global_step is %d when stop." % run_values.results, flush=True)
            run_context.request_stop()


def main():
    cluster, server, num_workers = bring_up(FLAGS)
    me, chief = FLAGS.task_index, FLAGS.task_index == 0
    fabric = dtf.fabric.FabricPSStrategy(server) if FLAGS.engine == "fabric" else None
    data = input_data.read_data_sets(FLAGS.data_dir, one_hot=True, num_train=FLAGS.num_train)
    print("len of train images: ", len(data.train.images))

    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:%d/cpu:0" % me)):
        net = build_mnist_mlp(hidden=FLAGS.hidden_units)
        gstep, loss = net["global_step"], net["loss"]
        opt = dtf.train.AdamOptimizer(FLAGS.learning_rate)
        hooks, stale = [AnnouncedStop(last_step=FLAGS.train_steps)], None
        if FLAGS.issync:
            print("is_sync:true")
            opt = dtf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=num_workers, total_num_replicas=num_workers)
            if fabric is None:
                hooks.append(opt.make_session_run_hook(chief))
        elif FLAGS.measure_staleness and fabric is None:
            stale = dtf.train.StalenessHook()
            hooks.append(stale)
        if fabric is None:
            train_op, loss_fetch = opt.minimize(loss, global_step=gstep), loss
        else:
            train_op, loss_fetch = fabric.minimize(opt, loss, gstep)       # same semantics, executed by the GPUs

    print("Worker %d: %s" % (me, "Initializing session..." if chief else "Waiting for session to be initialized..."))
    best, local_step, t0 = 10000.0, 0, time.time()
    print("Training begins @ %f" % t0)
    with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=chief, checkpoint_dir=FLAGS.train_dir,
                                            hooks=hooks) as sess:
        while not sess.should_stop():
            xs, ys = data.train.next_batch(FLAGS.batch_size)
            _, step, batch_loss = sess.run([train_op, gstep, loss_fetch], feed_dict={net["x"]: xs, net["y_"]: ys})
            local_step += 1
            if local_step % FLAGS.log_every == 0:
                print("time: %s | worker: %d | training step:%d | global step:%d | loss: %f"
                      % (datetime.now(), me, local_step, step, batch_loss))
            if (step + 1) % FLAGS.validate_every == 0 and not sess.should_stop():
                val = sess.run(loss, feed_dict={net["x"]: data.validation.images, net["y_"]: data.validation.labels})
                val /= len(data.validation.images)
                best = min(best, val)
                print("At global step: %d, validation cross entropy = %g (best %g)" % (step, val, best))
    t1 = time.time()
    print("Training ends @ %f" % t1)
    print("Worker %d | Training elapsed time: %f s | Train step: %d | best val loss: %f" % (me, t1 - t0, local_step, best))
    if stale is not None:
        print("Worker %d | staleness mean %.3f histogram %s" % (me, stale.mean(), stale.histogram()))
    server.stop()


if __name__ == "__main__":
    main()
