"""ClusterSpec: immutable job -> task-address map.

Capability parity: ``tf.train.ClusterSpec({'ps': [...], 'worker': [...]})`` as
used at reference ``distributed_mnist.py:74``, ``example_between_graph.py:31``,
``example_in_graph.py:28``.  Jobs are ``ps`` and ``worker`` in every reference
script but any job name is accepted.

B200 mapping: besides host:port addresses (control plane), a ClusterSpec can
assign every task a GPU ordinal on the local 8xB200 box
(:meth:`ClusterSpec.device_map`): ps tasks first, then workers, one GPU per
task, wrapping when tasks outnumber GPUs.  The fabric layer
(``parallel/fabric.py``) uses this to decide which ranks exchange peer-memory
handles.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Sequence, Tuple, Union

__all__ = ["ClusterSpec"]


def _clean(addr: str) -> str:
    # The reference's default worker_hosts has a space after the comma
    # (distributed_mnist.py:29); be forgiving about whitespace.
    return str(addr).strip()


class ClusterSpec:
    def __init__(self, cluster: Union["ClusterSpec", Mapping[str, Union[Sequence[str], Mapping[int, str]]]]):
        spec: Dict[str, Dict[int, str]] = {}
        if isinstance(cluster, ClusterSpec):
            spec = {j: dict(t) for j, t in cluster._spec.items()}
        elif isinstance(cluster, Mapping):
            for job, tasks in cluster.items():
                if isinstance(tasks, Mapping):
                    spec[str(job)] = {int(i): _clean(a) for i, a in tasks.items()}
                elif isinstance(tasks, str):
                    spec[str(job)] = {i: _clean(a) for i, a in enumerate(tasks.split(",")) if _clean(a)}
                else:
                    spec[str(job)] = {i: _clean(a) for i, a in enumerate(tasks)}
        else:
            raise TypeError("ClusterSpec needs a dict or another ClusterSpec, got %r" % type(cluster))
        for job, tasks in spec.items():
            for i, a in tasks.items():
                if not a:
                    raise ValueError("empty address for /job:%s/task:%d" % (job, i))
        self._spec = spec

    # -- queries ----------------------------------------------------------
    @property
    def jobs(self) -> List[str]:
        return sorted(self._spec)

    def num_tasks(self, job_name: str) -> int:
        return len(self._job(job_name))

    def task_indices(self, job_name: str) -> List[int]:
        return sorted(self._job(job_name))

    def task_address(self, job_name: str, task_index: int) -> str:
        job = self._job(job_name)
        try:
            return job[int(task_index)]
        except KeyError:
            raise ValueError("no task %d in job %r" % (task_index, job_name)) from None

    def job_tasks(self, job_name: str) -> List[str]:
        job = self._job(job_name)
        return [job[i] for i in sorted(job)]

    def as_dict(self) -> Dict[str, Union[List[str], Dict[int, str]]]:
        out: Dict[str, Union[List[str], Dict[int, str]]] = {}
        for job, tasks in self._spec.items():
            idx = sorted(tasks)
            if idx == list(range(len(idx))):
                out[job] = [tasks[i] for i in idx]
            else:
                out[job] = dict(tasks)
        return out

    def all_tasks(self) -> List[Tuple[str, int, str]]:
        """(job, task, address) triples; ``ps`` first then the other jobs sorted."""
        order = ([j for j in ("ps",) if j in self._spec] +
                 [j for j in sorted(self._spec) if j != "ps"])
        return [(j, i, self._spec[j][i]) for j in order for i in sorted(self._spec[j])]

    def find_task(self, address: str) -> Tuple[str, int]:
        """Reverse lookup: which /job/task listens on ``address``."""
        address = _clean(address)
        for job, i, a in self.all_tasks():
            if a == address or _same_endpoint(a, address):
                return job, i
        raise ValueError("address %r is not part of the cluster %r" % (address, self.as_dict()))

    def device_map(self, num_gpus: int) -> Dict[Tuple[str, int], int]:
        """Assign each task a GPU ordinal on one box: ps tasks, then workers."""
        out: Dict[Tuple[str, int], int] = {}
        if num_gpus <= 0:
            return {(j, i): -1 for j, i, _ in self.all_tasks()}
        for n, (j, i, _) in enumerate(self.all_tasks()):
            out[(j, i)] = n % num_gpus
        return out

    # -- plumbing ---------------------------------------------------------
    def _job(self, job_name: str) -> Dict[int, str]:
        try:
            return self._spec[job_name]
        except KeyError:
            raise ValueError("no such job %r in cluster (jobs: %s)" % (job_name, self.jobs)) from None

    def __bool__(self) -> bool:
        return bool(self._spec)

    def __eq__(self, other: object) -> bool:
        return isinstance(other, ClusterSpec) and self._spec == other._spec

    def __ne__(self, other: object) -> bool:
        return not self == other

    def __hash__(self) -> int:
        return hash(tuple((j, tuple(sorted(t.items()))) for j, t in sorted(self._spec.items())))

    def __repr__(self) -> str:
        return "ClusterSpec(%r)" % (self.as_dict(),)


def _same_endpoint(a: str, b: str) -> bool:
    def norm(x: str) -> Tuple[str, str]:
        host, _, port = x.rpartition(":")
        if host in ("localhost", "0.0.0.0", ""):
            host = "127.0.0.1"
        return host, port
    return norm(a) == norm(b)
