"""Between-graph parameter-server training of the MNIST MLP, async or sync.

Capability mirror of reference ``distributed_mnist.py`` (S1-S8): same flags, same
model (784 -> hidden_units ReLU -> 10 softmax, batch-sum clipped cross-entropy),
Adam, ``replica_device_setter`` placement, optional ``SyncReplicasOptimizer``,
``MonitoredTrainingSession`` with a custom stop hook, per-step log line,
validation every 1000 global steps, final summary.

Run one process per task, e.g. on one box::

    python examples/distributed_mnist.py --job_name=ps     --task_index=0 --ps_hosts=127.0.0.1:22221 --worker_hosts=127.0.0.1:22222,127.0.0.1:22223 &
    python examples/distributed_mnist.py --job_name=worker --task_index=0 ... --issync=True &
    python examples/distributed_mnist.py --job_name=worker --task_index=1 ... --issync=True

Differences from the reference, on purpose: ``--train_steps`` is honoured
(the reference hard-codes 10000), the checkpoint dir is a flag, data is synthetic
MNIST-shaped when no IDX files are present, and GPUs are used when visible
(``DTF_GPU_INDEX`` binds a task to one B200) instead of being hidden.
"""
import math
import os
import sys
import time
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200 import input_data

flags = dtf.app.flags
IMAGE_PIXELS = 28
flags.DEFINE_string('data_dir', '/tmp/mnist-data', 'Directory for storing mnist data (IDX files); synthetic if absent')
flags.DEFINE_integer('hidden_units', 100, 'Number of units in the hidden layer of the NN')
flags.DEFINE_integer('train_steps', 10000, 'Global step at which training stops')
flags.DEFINE_integer('batch_size', 100, 'Training batch size')
flags.DEFINE_float('learning_rate', 0.01, 'Learning rate')
flags.DEFINE_string('ps_hosts', '127.0.0.1:22221', 'Comma-separated list of hostname:port pairs')
flags.DEFINE_string('worker_hosts', '127.0.0.1:22222,127.0.0.1:22223', 'Comma-separated list of hostname:port pairs')
flags.DEFINE_string('job_name', 'worker', 'job name: worker or ps')
flags.DEFINE_integer('task_index', 0, 'Index of task within the job')
flags.DEFINE_bool('issync', False, 'Use synchronous replicas (SyncReplicasOptimizer)')
flags.DEFINE_string('train_dir', '/tmp/dtf_ckpt/mnist', 'Checkpoint directory (shared filesystem)')
flags.DEFINE_integer('validate_every', 1000, 'Validate when (global_step+1) is a multiple of this')
flags.DEFINE_integer('num_train', 55000, 'Synthetic train-set size')
flags.DEFINE_integer('log_every', 1, 'Print the per-step line every N local steps')
flags.DEFINE_bool('measure_staleness', False, 'Async mode: record pull->apply staleness per step')
flags.DEFINE_string('engine', 'graph', "'graph': control-plane tier (RPC); 'fabric': parameters/gradients/tokens over NVLink peer memory")
FLAGS = flags.FLAGS


class MyStopAtStepHook(dtf.train.StopAtStepHook):
    """Stop hook that reports where it started and where it stops."""

    def after_create_session(self, session, coord):
        if self._last_step is None:
            global_step = session.run(self._global_step_tensor)
            self._last_step = global_step + self._num_steps
            print("now global_step is %d after create session, num_steps: %d, last_step:%d :"
                  % (global_step, self._num_steps, self._last_step))

    def after_run(self, run_context, run_values):
        global_step = run_values.results
        if global_step >= self._last_step:
            print("global_step is %d when stop." % global_step)
            run_context.request_stop()


def build_model(hidden_units):
    global_step = dtf.train.get_or_create_global_step()
    hid_w = dtf.Variable(dtf.truncated_normal([IMAGE_PIXELS * IMAGE_PIXELS, hidden_units],
                                              stddev=1.0 / IMAGE_PIXELS), name='hid_w')
    hid_b = dtf.Variable(dtf.zeros([hidden_units]), name='hid_b')
    sm_w = dtf.Variable(dtf.truncated_normal([hidden_units, 10], stddev=1.0 / math.sqrt(hidden_units)), name='sm_w')
    sm_b = dtf.Variable(dtf.zeros([10]), name='sm_b')
    x = dtf.placeholder(dtf.float32, [None, IMAGE_PIXELS * IMAGE_PIXELS])
    y_ = dtf.placeholder(dtf.float32, [None, 10])
    hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    y = dtf.nn.softmax(dtf.nn.xw_plus_b(hid, sm_w, sm_b))
    cross_entropy = -dtf.reduce_sum(y_ * dtf.log(dtf.clip_by_value(y, 1e-10, 1.0)))
    return global_step, x, y_, y, cross_entropy


def main():
    if FLAGS.job_name is None or FLAGS.job_name == '':
        raise ValueError('Must specify an explicit job_name !')
    print('job_name : %s' % FLAGS.job_name)
    if FLAGS.task_index is None or FLAGS.task_index == '':
        raise ValueError('Must specify an explicit task_index!')
    print('task_index : %d' % FLAGS.task_index)

    ps_spec = [h.strip() for h in FLAGS.ps_hosts.split(',')]
    worker_spec = [h.strip() for h in FLAGS.worker_hosts.split(',')]
    num_workers = len(worker_spec)
    cluster = dtf.train.ClusterSpec({'ps': ps_spec, 'worker': worker_spec})
    server = dtf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)
    if FLAGS.job_name == 'ps':
        server.join()        # the ps only owns variables / accumulators / queues (fabric: + the apply service); blocks
        return
    strategy = dtf.fabric.FabricPSStrategy(server) if FLAGS.engine == 'fabric' else None

    mnist = input_data.read_data_sets(FLAGS.data_dir, one_hot=True, num_train=FLAGS.num_train)
    print("len of train images: ", len(mnist.train.images))
    worker_device = '/job:worker/task:%d/cpu:0' % FLAGS.task_index
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device=worker_device)):
        global_step, x, y_, y, cross_entropy = build_model(FLAGS.hidden_units)
        opt = dtf.train.AdamOptimizer(FLAGS.learning_rate)
        hooks = [MyStopAtStepHook(last_step=FLAGS.train_steps)]
        staleness = None
        if FLAGS.issync:
            print("is_sync:true")
            opt = dtf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=num_workers,
                                                  total_num_replicas=num_workers)
            if strategy is None:
                hooks.append(opt.make_session_run_hook(FLAGS.task_index == 0))
        elif FLAGS.measure_staleness and strategy is None:
            staleness = dtf.train.StalenessHook()
            hooks.append(staleness)
        if strategy is not None:
            # same program, but pull / push / aggregate / tokens run on the GPUs over NVLink
            train_step, loss_fetch = strategy.minimize(opt, cross_entropy, global_step)
        else:
            train_step, loss_fetch = opt.minimize(cross_entropy, global_step=global_step), cross_entropy

        is_chief = (FLAGS.task_index == 0)
        if is_chief:
            print('Worker %d: Initializing session...' % FLAGS.task_index)
        else:
            print('Worker %d: Waiting for session to be initialized...' % FLAGS.task_index)

        local_step = 0
        best_val_loss = 10000.0
        time_begin = time.time()
        print('Training begins @ %f' % time_begin)
        with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=is_chief,
                                                checkpoint_dir=FLAGS.train_dir, hooks=hooks) as mon_sess:
            while not mon_sess.should_stop():
                batch_xs, batch_ys = mnist.train.next_batch(FLAGS.batch_size)
                _, step, loss = mon_sess.run([train_step, global_step, loss_fetch],
                                             feed_dict={x: batch_xs, y_: batch_ys})
                local_step += 1
                if local_step % FLAGS.log_every == 0:
                    print('time: %s | worker: %d | training step:%d | global step:%d | loss: %f' % (
                        str(datetime.now()), FLAGS.task_index, local_step, step, loss))
                if (step + 1) % FLAGS.validate_every == 0 and not mon_sess.should_stop():
                    val_feed = {x: mnist.validation.images, y_: mnist.validation.labels}
                    val_xent = mon_sess.run(cross_entropy, feed_dict=val_feed) / len(mnist.validation.images)
                    best_val_loss = min(best_val_loss, val_xent)
                    print('At global step: %d, validation cross entropy = %g (best %g)' % (step, val_xent, best_val_loss))
        time_end = time.time()
        print('Training ends @ %f' % time_end)
        print('Worker %d | Training elapsed time: %f s | Train step: %d | best val loss: %f' %
              (FLAGS.task_index, time_end - time_begin, local_step, best_val_loss))
        if staleness is not None:
            print('Worker %d | staleness mean %.3f histogram %s' % (FLAGS.task_index, staleness.mean(),
                                                                   staleness.histogram()))
    server.stop()


if __name__ == '__main__':
    main()
