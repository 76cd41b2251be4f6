"""Shared plumbing of the example programs: cluster flags, task bring-up, a per-step log hook.

Every example is one program started once per task (``--job_name ps|worker --task_index N``), like the reference's
scripts (``distributed_mnist.py:58-79``); the parts that are identical in all of them live here."""
import os
import sys
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf  # noqa: E402


def define_cluster_flags(ps_hosts, worker_hosts, job_name="worker"):
    f = dtf.app.flags
    f.DEFINE_string("ps_hosts", ps_hosts, "comma-separated host:port list of the ps tasks")
    f.DEFINE_string("worker_hosts", worker_hosts, "comma-separated host:port list of the worker tasks")
    f.DEFINE_string("job_name", job_name, "'ps' or 'worker'")
    f.DEFINE_integer("task_index", 0, "index of this task inside its job")
    return f.FLAGS


def bring_up(FLAGS, serve_only=None):
    """ClusterSpec + this task's Server.  Tasks for which ``serve_only(job, index)`` is true never return: they
    own variables / accumulators / queues and execute what clients send them (default: every ps task)."""
    if not FLAGS.job_name:
        raise ValueError("--job_name is required (ps or worker)")
    hosts = {"ps": [h.strip() for h in FLAGS.ps_hosts.split(",") if h.strip()],
             "worker": [h.strip() for h in FLAGS.worker_hosts.split(",") if h.strip()]}
    cluster = dtf.train.ClusterSpec(hosts)
    server = dtf.train.Server(cluster, job_name=FLAGS.job_name, task_index=FLAGS.task_index)
    print("job_name : %s\ntask_index : %d" % (FLAGS.job_name, FLAGS.task_index), flush=True)
    passive = serve_only(FLAGS.job_name, FLAGS.task_index) if serve_only else FLAGS.job_name == "ps"
    if passive:
        server.join()
        sys.exit(0)
    return cluster, server, len(hosts["worker"])


class StepLogger(dtf.train.SessionRunHook):
    """Prints ``fmt % values`` every ``every`` local steps; ``fetches`` are added to each run call."""

    def __init__(self, fetches, fmt, every=1, worker=0):
        self.fetches, self.fmt, self.every, self.worker, self.local_step = fetches, fmt, max(1, every), worker, 0

    def before_run(self, run_context):
        return dtf.train.SessionRunArgs(self.fetches)

    def after_run(self, run_context, run_values):
        self.local_step += 1
        if self.local_step % self.every == 0:
            print(self.fmt(datetime.now(), self.worker, self.local_step, run_values.results), flush=True)
