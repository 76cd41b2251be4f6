"""Pure client for an already-running cluster (reference ``example_distributed_client.py``, S13).

No ClusterSpec, no Server: the client connects to worker 0's master (``--master``), which
tells it the cluster layout; the graph places variables on the ps and one matmul per worker.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200 import timeline

dtf.app.flags.DEFINE_string("master", "grpc://localhost:2223", "master (worker 0) target")
dtf.app.flags.DEFINE_string("out_dir", ".", "where logs/ and timeline_client.json go")
FLAGS = dtf.app.flags.FLAGS


def main():
    with dtf.device('/job:ps/task:0/cpu:0'):
        input_data = dtf.Variable([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.], [10., 11., 12.]], name="input_data")
        b = dtf.Variable([[1.], [1.], [2.]], name="w")
    inputs = dtf.split(input_data, 2)
    outputs = []
    run_options = dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE)
    run_metadata = dtf.RunMetadata()
    # in-graph replication: this is the only client
    with dtf.Session(FLAGS.master) as sess:
        sess.run(dtf.global_variables_initializer())
        for i in range(2):   # 2 workers
            with dtf.device("/job:worker/task:%d/cpu:0" % i):
                print(sess.run(inputs[i]))
                outputs.append(dtf.matmul(inputs[i], b))
        with dtf.device('/job:ps/task:0/cpu:0'):
            output = dtf.concat(outputs, axis=0)
            print(sess.run(output, options=run_options, run_metadata=run_metadata))
        dtf.summary.FileWriter(os.path.join(FLAGS.out_dir, "logs/"), sess.graph).close()
        tl = timeline.Timeline(step_stats=run_metadata.step_stats)
        with open(os.path.join(FLAGS.out_dir, 'timeline_client.json'), 'w') as f:
            f.write(tl.generate_chrome_trace_format())


if __name__ == "__main__":
    main()
