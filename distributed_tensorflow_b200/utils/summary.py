"""Event-file writer: graph dump + scalar summaries.

``summary.FileWriter("logs/", sess.graph)`` writes an event file holding the
graph description (reference ``example_in_graph.py:62``,
``example_distributed_client.py:41``).  Format: JSON-lines
``events.out.dtfevents.<time>.<host>`` -- one record per line
(``{"wall_time", "step", "graph_def" | "scalar": {tag, value}}``).
"""
from __future__ import annotations

import json
import os
import socket
import threading
import time
from typing import Any, Dict, List, Optional

from ..framework.graph import GraphKeys, get_default_graph

__all__ = ["FileWriter", "scalar", "merge_all", "read_events"]


class FileWriter:
    def __init__(self, logdir: str, graph=None, max_queue: int = 10, flush_secs: float = 120, filename_suffix: str = ""):
        os.makedirs(logdir, exist_ok=True)
        self._path = os.path.join(logdir, "events.out.dtfevents.%d.%s%s" % (time.time(), socket.gethostname(),
                                                                          filename_suffix))
        self._f = open(self._path, "a")
        self._lock = threading.Lock()
        self._write({"wall_time": time.time(), "file_version": "dtf-events-1"})
        if graph is not None:
            self.add_graph(graph)

    @property
    def path(self) -> str:
        return self._path

    def get_logdir(self) -> str:
        return os.path.dirname(self._path)

    def _write(self, rec: Dict[str, Any]) -> None:
        with self._lock:
            self._f.write(json.dumps(rec) + "\n")

    def add_graph(self, graph, global_step: Optional[int] = None) -> None:
        self._write({"wall_time": time.time(), "step": global_step, "graph_def": graph.as_graph_def()})
        self.flush()

    def add_scalar(self, tag: str, value: float, global_step: Optional[int] = None) -> None:
        self._write({"wall_time": time.time(), "step": global_step, "scalar": {"tag": tag, "value": float(value)}})

    def add_summary(self, summary: Dict[str, float], global_step: Optional[int] = None) -> None:
        for k, v in summary.items():
            self.add_scalar(k, v, global_step)

    def add_run_metadata(self, run_metadata, tag: str, global_step: Optional[int] = None) -> None:
        self._write({"wall_time": time.time(), "step": global_step, "run_metadata": {"tag": tag,
                     "step_stats": run_metadata.step_stats}})

    def flush(self) -> None:
        with self._lock:
            self._f.flush()

    def close(self) -> None:
        with self._lock:
            if not self._f.closed:
                self._f.flush()
                self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class _ScalarSummary:
    def __init__(self, tag: str, tensor):
        self.tag, self.tensor, self.name = tag, tensor, tag


def scalar(name: str, tensor, collections=None) -> _ScalarSummary:
    s = _ScalarSummary(name, tensor)
    get_default_graph().add_to_collection(GraphKeys.SUMMARIES, s)
    return s


def merge_all() -> Dict[str, Any]:
    return {s.tag: s.tensor for s in get_default_graph().get_collection(GraphKeys.SUMMARIES)}


def read_events(path: str) -> List[Dict[str, Any]]:
    with open(path) as f:
        return [json.loads(l) for l in f if l.strip()]
