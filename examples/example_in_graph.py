"""In-graph replication: ONE client graph spanning a ps and two workers, with a Timeline trace.

Capability mirror of reference ``example_in_graph.py`` (S11) and its twin
``example_distributed_server.py`` (S12): every task starts a Server; worker 0 is the only
client; variables are pinned to ``/job:ps/task:0``, the input is split on the ps, each half is
multiplied on a different worker, the results are concatenated on the ps; the last ``run`` is
traced (``FULL_TRACE``) and written as ``timeline_client.json``; the graph is dumped with
``summary.FileWriter``.  Expected output: ``[[9],[21],[33],[45]]``.
Fixes vs the reference: the master address comes from ``--worker_hosts`` instead of a
hard-coded ``grpc://localhost:2223``; ``/gpu:0`` placements fall back to CPU when the task
has no GPU (soft placement) instead of failing.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200 import timeline

dtf.app.flags.DEFINE_string("ps_hosts", "localhost:2222", "ps hosts")
dtf.app.flags.DEFINE_string("worker_hosts", "localhost:2223,localhost:2224", "worker hosts")
dtf.app.flags.DEFINE_string("job_name", "worker", "'ps' or 'worker'")
dtf.app.flags.DEFINE_integer("task_index", 0, "Index of task within the job")
dtf.app.flags.DEFINE_string("out_dir", ".", "where logs/ and timeline_client.json go")
FLAGS = dtf.app.flags.FLAGS


def main():
    ps_hosts = FLAGS.ps_hosts.split(",")
    worker_hosts = FLAGS.worker_hosts.split(",")
    # identical on every node
    cluster = dtf.train.ClusterSpec({"ps": ps_hosts, "worker": worker_hosts})
    server = dtf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)

    with dtf.device('/job:ps/task:0/cpu:0'):
        input_data = dtf.Variable([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.], [10., 11., 12.]], name="input_data")
        b = dtf.Variable([[1.], [1.], [2.]], name="w")
    inputs = dtf.split(input_data, 2)
    outputs = []

    run_options = dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE)
    run_metadata = dtf.RunMetadata()

    if FLAGS.job_name == 'ps' or FLAGS.task_index != 0:
        server.join()      # ps and non-client workers only serve
        return
    # in-graph replication: only worker 0 creates a client
    with dtf.Session("grpc://" + worker_hosts[0]) as sess:
        sess.run(dtf.global_variables_initializer())
        for i in range(len(worker_hosts)):
            with dtf.device("/job:worker/task:%d/gpu:0" % i):
                print("now is worker %d: " % i)
                print(sess.run(inputs[i % 2]))
                outputs.append(dtf.matmul(inputs[i % 2], b))
        with dtf.device('/job:ps/task:0/cpu:0'):
            output = dtf.concat(outputs[:2], axis=0)
            print(sess.run(output, options=run_options, run_metadata=run_metadata))
        dtf.summary.FileWriter(os.path.join(FLAGS.out_dir, "logs/"), sess.graph).close()
        tl = timeline.Timeline(step_stats=run_metadata.step_stats)
        with open(os.path.join(FLAGS.out_dir, 'timeline_client.json'), 'w') as f:
            f.write(tl.generate_chrome_trace_format())
    server.stop()


if __name__ == "__main__":
    main()
