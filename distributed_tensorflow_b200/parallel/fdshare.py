"""File-descriptor passing between the task processes of one machine (unix sockets, ``SCM_RIGHTS``).

CUDA VMM allocations and NVLS multicast objects are shared between processes as POSIX file descriptors
(``csrc/fabric_vmm.cu``); an fd number means nothing in another process, so the owner serves its descriptors
from a small unix-socket server and the peers fetch duplicates by name.  This is control plane only (runs once
per buffer at start-up) -- the TF analogue is the gRPC channel set-up behind ``tf.train.Server``
(reference ``distributed_mnist.py:75``).
"""
from __future__ import annotations

import os
import socket
import tempfile
import threading
import time
from typing import Dict, Optional

__all__ = ["FdServer", "fetch_fd"]


class FdServer:
    """Serves registered descriptors: a client sends a name (one datagram-sized line), gets ``b"ok"`` + the fd."""

    def __init__(self, tag: str = "dtf"):
        d = tempfile.mkdtemp(prefix="%s_fd_%d_" % (tag, os.getpid()))
        self.path = os.path.join(d, "s")
        self._fds: Dict[str, int] = {}
        self._cv = threading.Condition()
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._sock.bind(self.path)
        self._sock.listen(64)
        self._stop = False
        self._thread = threading.Thread(target=self._serve, name="dtf-fdserver", daemon=True)
        self._thread.start()

    def register(self, name: str, fd: int) -> None:
        with self._cv:
            self._fds[name] = fd
            self._cv.notify_all()

    def _serve(self) -> None:
        while not self._stop:
            try:
                conn, _ = self._sock.accept()
            except OSError:
                return
            threading.Thread(target=self._handle, args=(conn,), daemon=True).start()

    def _handle(self, conn: socket.socket) -> None:
        try:
            with conn:
                name = conn.recv(4096).decode()
                deadline = time.time() + 120.0
                with self._cv:
                    while name not in self._fds and time.time() < deadline and not self._stop:
                        self._cv.wait(timeout=0.5)
                    fd = self._fds.get(name)
                if fd is None:
                    conn.sendall(b"no")
                else:
                    socket.send_fds(conn, [b"ok"], [fd])
        except OSError:
            pass

    def close(self) -> None:
        self._stop = True
        try:
            self._sock.close()
        except OSError:
            pass
        try:
            os.unlink(self.path)
            os.rmdir(os.path.dirname(self.path))
        except OSError:
            pass


def fetch_fd(path: str, name: str, timeout: float = 120.0) -> int:
    """Duplicate of the descriptor another process registered under ``name`` (valid in THIS process)."""
    deadline = time.time() + timeout
    last: Optional[BaseException] = None
    while time.time() < deadline:
        try:
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.settimeout(max(1.0, deadline - time.time()))
                c.connect(path)
                c.sendall(name.encode())
                msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
                if msg == b"ok" and fds:
                    return fds[0]
                raise RuntimeError("fd %r is not served by %s" % (name, path))
        except (ConnectionRefusedError, FileNotFoundError) as e:      # server not up yet
            last = e
            time.sleep(0.05)
    raise TimeoutError("could not fetch fd %r from %s: %r" % (name, path, last))
