"""A pure client: no ClusterSpec, no Server -- just a master address (reference ``example_distributed_client.py``, S13).

Start the tasks first (``example_distributed_server.py`` once per ps / worker), then run this anywhere that can
reach worker 0.  The master tells the client the cluster layout; the program is the scatter / matmul-per-worker /
gather graph of ``example_in_graph.py`` with the matmuls pinned to each worker's CPU.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py

from _common import dtf
from distributed_tensorflow_b200 import timeline
from example_in_graph import PS0

dtf.app.flags.DEFINE_string("master", "grpc://localhost:2223", "target of worker 0's server (the master)")
dtf.app.flags.DEFINE_integer("workers", 2, "how many worker tasks the cluster has")
dtf.app.flags.DEFINE_string("out_dir", ".", "where logs/ and timeline_client.json are written")
FLAGS = dtf.app.flags.FLAGS


def main():
    with dtf.device(PS0):
        table = dtf.Variable([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.], [10., 11., 12.]], name="input_data")
        column = dtf.Variable([[1.], [1.], [2.]], name="w")
    pieces = dtf.split(table, FLAGS.workers)
    products = []
    meta = dtf.RunMetadata()
    with dtf.Session(FLAGS.master) as sess:
        sess.run(dtf.global_variables_initializer())
        for k, piece in enumerate(pieces):
            print(sess.run(piece))
            with dtf.device("/job:worker/task:%d/cpu:0" % k):
                products.append(dtf.matmul(piece, column))        # the graph may keep growing between runs
        with dtf.device(PS0):
            answer = dtf.concat(products, axis=0)
        print(sess.run(answer, options=dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE), run_metadata=meta))
        dtf.summary.FileWriter(os.path.join(FLAGS.out_dir, "logs/"), sess.graph).close()
    with open(os.path.join(FLAGS.out_dir, "timeline_client.json"), "w") as f:
        f.write(timeline.Timeline(step_stats=meta.step_stats).generate_chrome_trace_format())


if __name__ == "__main__":
    main()
