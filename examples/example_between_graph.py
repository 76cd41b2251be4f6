"""Between-graph PS training of ``y = w*x + b`` with SGD, async or sync.

Capability mirror of reference ``example_between_graph.py`` (S9/S10): cluster and server are
brought up at module level, parameters are placed by ``replica_device_setter``, the stop
condition is ``StopAtStepHook(num_steps=2000)``, checkpoints every 60 s, and every step does
a second ``run([weight, biase])`` to print the parameters approaching 2 and 10.
Fixes vs the reference: sync mode aggregates ``len(worker_hosts)`` replicas
(the decoupled ``--num_workers`` flag is kept only as an override), ``--steps_to_validate``
really throttles printing, the ``ConfigProto`` with the GPU memory fraction is passed on.
"""
import os
import sys
from datetime import datetime

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf

FLAGS = dtf.app.flags.FLAGS
dtf.app.flags.DEFINE_float('learning_rate', 0.03, 'Initial learning rate.')
dtf.app.flags.DEFINE_integer('steps_to_validate', 1, 'Print every N steps')
dtf.app.flags.DEFINE_string("ps_hosts", "127.0.0.1:2222", "Comma-separated list of hostname:port pairs")
dtf.app.flags.DEFINE_string("worker_hosts", "127.0.0.1:2223,127.0.0.1:2224", "Comma-separated list of hostname:port pairs")
dtf.app.flags.DEFINE_string("job_name", "worker", "One of 'ps', 'worker'")
dtf.app.flags.DEFINE_integer("task_index", 0, "Index of task within the job")
dtf.app.flags.DEFINE_bool("is_sync", False, "using synchronous training or not")
dtf.app.flags.DEFINE_integer("num_workers", 0, "replicas to aggregate in sync mode (0: number of worker hosts)")
dtf.app.flags.DEFINE_integer("num_steps", 2000, "steps to run after session creation")
dtf.app.flags.DEFINE_string("ckpt_dir", "/tmp/dtf_ckpt/linear", "checkpoint directory (shared filesystem)")
dtf.app.flags.DEFINE_integer("save_checkpoint_secs", 60, "checkpoint period")

learning_rate = FLAGS.learning_rate
ps_hosts = FLAGS.ps_hosts.split(",")
worker_hosts = FLAGS.worker_hosts.split(",")
cluster = dtf.train.ClusterSpec({"ps": ps_hosts, "worker": worker_hosts})
server = dtf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)

train_X = np.random.rand(100).astype(np.float32).reshape(-1)
train_Y = 2 * train_X + 10  # W=2, b=10

if FLAGS.job_name == "ps":
    server.join()
elif FLAGS.job_name == "worker":
    with dtf.device(dtf.train.replica_device_setter(worker_device="/job:worker/task:%d" % FLAGS.task_index,
                                                    cluster=cluster)):
        # shared by all workers: minimize() increments it, so it counts every worker's updates
        global_step = dtf.Variable(0, name='global_step', trainable=False, dtype=dtf.int64)
        X = dtf.placeholder(dtf.float32)
        y = dtf.placeholder(dtf.float32)
        weight = dtf.get_variable("weight", [1], dtf.float32, initializer=dtf.random_normal_initializer())
        biase = dtf.get_variable("biase", [1], dtf.float32, initializer=dtf.random_normal_initializer())
        pred = dtf.multiply(X, weight) + biase
        loss_value = dtf.reduce_mean(dtf.square(y - pred))
        optimizer = dtf.train.GradientDescentOptimizer(learning_rate)
        hooks = [dtf.train.StopAtStepHook(num_steps=FLAGS.num_steps)]
        if FLAGS.is_sync:
            n = FLAGS.num_workers or len(worker_hosts)
            optimizer = dtf.train.SyncReplicasOptimizer(optimizer, replicas_to_aggregate=n, total_num_replicas=n)
            hooks.append(optimizer.make_session_run_hook(FLAGS.task_index == 0))
        train_op = optimizer.minimize(loss_value, global_step=global_step)

        config = dtf.ConfigProto(gpu_options=dtf.GPUOptions(per_process_gpu_memory_fraction=0.1))
        with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=(FLAGS.task_index == 0),
                                                checkpoint_dir=FLAGS.ckpt_dir,
                                                save_checkpoint_secs=FLAGS.save_checkpoint_secs,
                                                hooks=hooks, config=config) as mon_sess:
            while not mon_sess.should_stop():
                # mon_sess.run recovers from AbortedError/UnavailableError when a ps is preempted
                _, loss, step = mon_sess.run([train_op, loss_value, global_step], feed_dict={X: train_X, y: train_Y})
                if step % FLAGS.steps_to_validate == 0 and not mon_sess.should_stop():
                    w, b = mon_sess.run([weight, biase])
                    print("time: %s, step: %d, weight: %f, biase: %f, loss: %f" % (
                        str(datetime.now()), step, w[0], b[0], loss))
    server.stop()
