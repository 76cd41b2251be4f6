"""Server side of the client/server in-graph example (reference ``example_distributed_server.py``, S12).

Starts this task's server and serves until killed.  Unlike the reference twin (which repeats
the whole client program), the server variant here ONLY serves: start one per task, then run
``example_distributed_client.py`` anywhere that can reach worker 0.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf

dtf.app.flags.DEFINE_string("ps_hosts", "localhost:2222", "ps hosts")
dtf.app.flags.DEFINE_string("worker_hosts", "localhost:2223,localhost:2224", "worker hosts")
dtf.app.flags.DEFINE_string("job_name", "worker", "'ps' or 'worker'")
dtf.app.flags.DEFINE_integer("task_index", 0, "Index of task within the job")
FLAGS = dtf.app.flags.FLAGS


def main():
    cluster = dtf.train.ClusterSpec({"ps": FLAGS.ps_hosts.split(","), "worker": FLAGS.worker_hosts.split(",")})
    server = dtf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)
    print("serving /job:%s/task:%d at %s" % (FLAGS.job_name, FLAGS.task_index, server.target), flush=True)
    server.join()


if __name__ == "__main__":
    main()
