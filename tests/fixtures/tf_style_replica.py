"""A TF-1.x style parameter-server program (``import tensorflow as tf``, ``tf.app.run``, ``tf.train.Supervisor``,
``SyncReplicasOptimizer`` with the chief queue runner), written for the test-suite in the idiom of the classic
``mnist_replica.py``-era scripts.  It is run through ``distributed_tensorflow_b200.compat`` by
``tests/test_reference_scripts_unmodified.py`` to check that programs of that generation work unchanged."""
import math
import tempfile
import time

import tensorflow as tf
from tensorflow.examples.tutorials.mnist import input_data

flags = tf.app.flags
flags.DEFINE_string("data_dir", "/tmp/mnist-data", "Directory for storing mnist data")
flags.DEFINE_integer("task_index", None, "Worker task index; task_index=0 is the chief")
flags.DEFINE_integer("replicas_to_aggregate", None, "Number of replicas to aggregate before parameter update")
flags.DEFINE_integer("hidden_units", 50, "Number of units in the hidden layer")
flags.DEFINE_integer("train_steps", 120, "Number of (global) training steps to perform")
flags.DEFINE_integer("batch_size", 100, "Training batch size")
flags.DEFINE_float("learning_rate", 0.01, "Learning rate")
flags.DEFINE_boolean("sync_replicas", False, "Use the sync_replicas (synchronized replicas) mode")
flags.DEFINE_string("ps_hosts", "localhost:2222", "Comma-separated list of hostname:port pairs")
flags.DEFINE_string("worker_hosts", "localhost:2223,localhost:2224", "Comma-separated list of hostname:port pairs")
flags.DEFINE_string("job_name", None, "job name: worker or ps")
FLAGS = flags.FLAGS
IMAGE_PIXELS = 28


def main(unused_argv):
    mnist = input_data.read_data_sets(FLAGS.data_dir, one_hot=True)
    if FLAGS.job_name is None or FLAGS.job_name == "":
        raise ValueError("Must specify an explicit `job_name`")
    if FLAGS.task_index is None or FLAGS.task_index == "":
        raise ValueError("Must specify an explicit `task_index`")
    ps_spec = FLAGS.ps_hosts.split(",")
    worker_spec = FLAGS.worker_hosts.split(",")
    num_workers = len(worker_spec)
    cluster = tf.train.ClusterSpec({"ps": ps_spec, "worker": worker_spec})
    server = tf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)
    if FLAGS.job_name == "ps":
        server.join()
    is_chief = (FLAGS.task_index == 0)
    worker_device = "/job:worker/task:%d/cpu:0" % FLAGS.task_index
    with tf.device(tf.train.replica_device_setter(worker_device=worker_device, ps_device="/job:ps/cpu:0", cluster=cluster)):
        global_step = tf.Variable(0, name="global_step", trainable=False)
        images = tf.placeholder(tf.float32, [None, IMAGE_PIXELS * IMAGE_PIXELS])
        targets = tf.placeholder(tf.float32, [None, 10])
        activations, fan_in = images, IMAGE_PIXELS * IMAGE_PIXELS
        for depth, width in enumerate([FLAGS.hidden_units, 10]):
            with tf.variable_scope("layer%d" % depth):
                weights = tf.get_variable("weights", [fan_in, width], initializer=tf.truncated_normal_initializer(stddev=1.0 / math.sqrt(fan_in)))
                biases = tf.get_variable("biases", [width], initializer=tf.zeros_initializer())
            activations = tf.nn.xw_plus_b(activations, weights, biases)
            if width != 10:
                activations = tf.nn.relu(activations)
            fan_in = width
        cross_entropy = tf.reduce_sum(tf.nn.softmax_cross_entropy_with_logits(labels=targets, logits=activations))
        x, y_ = images, targets
        opt = tf.train.AdamOptimizer(FLAGS.learning_rate)
        if FLAGS.sync_replicas:
            replicas_to_aggregate = num_workers if FLAGS.replicas_to_aggregate is None else FLAGS.replicas_to_aggregate
            opt = tf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=replicas_to_aggregate,
                                                 total_num_replicas=num_workers, name="mnist_sync_replicas")
        train_step = opt.minimize(cross_entropy, global_step=global_step)
        if FLAGS.sync_replicas:
            local_init_op = opt.local_step_init_op
            if is_chief:
                local_init_op = opt.chief_init_op
            ready_for_local_init_op = opt.ready_for_local_init_op
            chief_queue_runner = opt.get_chief_queue_runner()
            sync_init_op = opt.get_init_tokens_op()
        init_op = tf.global_variables_initializer()
        train_dir = tempfile.mkdtemp()
        if FLAGS.sync_replicas:
            sv = tf.train.Supervisor(is_chief=is_chief, logdir=train_dir, init_op=init_op, local_init_op=local_init_op,
                                     ready_for_local_init_op=ready_for_local_init_op, recovery_wait_secs=1, global_step=global_step)
        else:
            sv = tf.train.Supervisor(is_chief=is_chief, logdir=train_dir, init_op=init_op, recovery_wait_secs=1,
                                     global_step=global_step)
        sess_config = tf.ConfigProto(allow_soft_placement=True, log_device_placement=False,
                                     device_filters=["/job:ps", "/job:worker/task:%d" % FLAGS.task_index])
        if is_chief:
            print("Worker %d: Initializing session..." % FLAGS.task_index)
        else:
            print("Worker %d: Waiting for session to be initialized..." % FLAGS.task_index)
        sess = sv.prepare_or_wait_for_session(server.target, config=sess_config)
        print("Worker %d: Session initialization complete." % FLAGS.task_index)
        if FLAGS.sync_replicas and is_chief:
            sess.run(sync_init_op)
            sv.start_queue_runners(sess, [chief_queue_runner])
        time_begin = time.time()
        local_step = 0
        while True:
            batch_xs, batch_ys = mnist.train.next_batch(FLAGS.batch_size)
            try:
                _, step = sess.run([train_step, global_step], feed_dict={x: batch_xs, y_: batch_ys})
            except tf.errors.OutOfRangeError:
                break
            local_step += 1
            if step >= FLAGS.train_steps:
                break
        print("Training elapsed time: %f s" % (time.time() - time_begin))
        val_xent = sess.run(cross_entropy, feed_dict={x: mnist.validation.images, y_: mnist.validation.labels})
        print("After %d training step(s), validation cross entropy = %g" % (FLAGS.train_steps, val_xent))
        sv.stop()


if __name__ == "__main__":
    tf.app.run()
