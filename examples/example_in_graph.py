"""In-graph replication: ONE client builds ONE graph that spans the ps and every worker.

Counterpart of the reference's ``example_in_graph.py`` (S11).  A [4,3] table and a [3,1] column live on
``/job:ps/task:0``; the table is split row-wise, shard ``k`` is multiplied by the column on ``/job:worker/task:k``
and the partial products are concatenated back on the ps -- scatter, parallel compute, gather.  Expected output
``[[9],[21],[33],[45]]``.  Worker 0 is the client (it connects to its own server, which acts as master); the ps and
the other workers only serve.  The final run is traced (``RunOptions.FULL_TRACE`` -> ``timeline_client.json``, one
chrome-trace process per device) and the graph is dumped for TensorBoard under ``logs/``.

    python examples/launch_local.py examples/example_in_graph.py --num_ps 1 --num_workers 2 --wait first
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

from _common import bring_up, define_cluster_flags, dtf
from distributed_tensorflow_b200 import timeline

FLAGS = define_cluster_flags("localhost:2222", "localhost:2223,localhost:2224")
dtf.app.flags.DEFINE_string("out_dir", ".", "where logs/ and timeline_client.json are written")
PS0 = "/job:ps/task:0/cpu:0"


def scatter_compute_gather(num_workers):
    with dtf.device(PS0):
        table = dtf.Variable([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.], [10., 11., 12.]], name="input_data")
        column = dtf.Variable([[1.], [1.], [2.]], name="w")
    shards = dtf.split(table, num_workers)
    partial = []
    for k, shard in enumerate(shards):
        with dtf.device("/job:worker/task:%d/gpu:0" % k):        # falls back to the task's CPU when it has no GPU
            partial.append(dtf.matmul(shard, column))
    with dtf.device(PS0):
        return shards, dtf.concat(partial, axis=0)


def main():
    _, server, num_workers = bring_up(FLAGS, serve_only=lambda job, idx: job == "ps" or idx != 0)
    shards, gathered = scatter_compute_gather(num_workers)
    meta = dtf.RunMetadata()
    with dtf.Session(server.target) as sess:
        sess.run(dtf.global_variables_initializer())
        for k, shard in enumerate(shards):
            print("now is worker %d: " % k)
            print(sess.run(shard))
        print(sess.run(gathered, options=dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE), run_metadata=meta))
        dtf.summary.FileWriter(os.path.join(FLAGS.out_dir, "logs/"), sess.graph).close()
    with open(os.path.join(FLAGS.out_dir, "timeline_client.json"), "w") as f:
        f.write(timeline.Timeline(step_stats=meta.step_stats).generate_chrome_trace_format())
    server.stop()


if __name__ == "__main__":
    main()
