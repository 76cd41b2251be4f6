"""Native transport under the control-plane RPC (``csrc/runtime/transport.cpp``; SURVEY A2 / A3: TF's server and its
``RecvTensor`` path are C++).

A message is a frame of segments: the envelope (a restricted pickle, protocol 5) plus one segment per tensor, handed to the
kernel straight from the tensors' memory (``sendmsg`` gather) and received straight into freshly allocated, aligned arrays
(``recv`` scatter) -- no concatenation on the way out, no second copy on the way in, no GIL while blocked.  Connections are
``TCP_NODELAY`` (strict request / reply with small envelopes).  :class:`NativeConnection` also speaks the two calls
``multiprocessing.connection``'s HMAC challenge needs (``send_bytes`` / ``recv_bytes``), so the authentication handshake of
``parallel/rpc.py`` is unchanged.

``DTF_NATIVE_TRANSPORT=0`` (or a missing ``libdtf_runtime.so``) falls back to ``multiprocessing.connection``; every task of a
cluster must make the same choice (the first frame of the other protocol fails the handshake).
"""
from __future__ import annotations

import ctypes
import errno
import os
import pickle
import threading
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

__all__ = ["available", "NativeConnection", "NativeListener", "connect", "TransportTimeout"]

_OK, _EOF, _TIMEOUT, _BAD, _ERR = 0, -1, -2, -3, -4
_MAX_SEG = 4096
_LIB = None
_TRIED = False
_LOCK = threading.Lock()


class TransportTimeout(OSError):
    pass


def _lib():
    global _LIB, _TRIED
    with _LOCK:
        if _TRIED:
            return _LIB
        _TRIED = True
        if os.environ.get("DTF_NATIVE_TRANSPORT", "1") == "0":
            return None
        from ..utils import native_runtime
        lib = native_runtime.load()
        if lib is None or not hasattr(lib, "dtf_net_send"):
            return None
        c_int, c_double, c_void_p, c_char_p = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_char_p
        lib.dtf_net_listen.argtypes, lib.dtf_net_listen.restype = [c_char_p, c_int, c_int], c_int
        lib.dtf_net_accept.argtypes, lib.dtf_net_accept.restype = [c_int, c_double], c_int
        lib.dtf_net_connect.argtypes, lib.dtf_net_connect.restype = [c_char_p, c_int, c_double], c_int
        lib.dtf_net_send.argtypes, lib.dtf_net_send.restype = [c_int, c_void_p, c_void_p, c_int], c_int
        lib.dtf_net_recv_header.argtypes, lib.dtf_net_recv_header.restype = [c_int, c_void_p, c_int, c_double], c_int
        lib.dtf_net_recv_body.argtypes, lib.dtf_net_recv_body.restype = [c_int, c_void_p, c_void_p, c_int], c_int
        lib.dtf_net_peer_closed.argtypes, lib.dtf_net_peer_closed.restype = [c_int], c_int
        lib.dtf_net_local_port.argtypes, lib.dtf_net_local_port.restype = [c_int], c_int
        lib.dtf_net_shutdown.argtypes = [c_int]
        lib.dtf_net_close.argtypes = [c_int]
        _LIB = lib
        return _LIB


def available() -> bool:
    return _lib() is not None


def _raise(rc: int, what: str):
    if rc == _TIMEOUT:
        raise TransportTimeout("%s: timed out" % what)
    if rc == _EOF:
        raise EOFError("%s: connection closed by the peer" % what)
    if rc == _BAD:
        raise ConnectionError("%s: not a frame of this protocol (a task speaking the other transport, or a stray client)" % what)
    raise OSError("%s failed (code %d)" % (what, rc))


class NativeConnection:
    """One TCP connection.  One reader and one writer at a time (the RPC layer uses one connection per calling thread / one
    handler thread per peer); ``close()`` may come from ANY thread: it shuts the socket down at once -- a thread blocked in
    ``recv`` wakes up with EOF, later sends fail -- but the descriptor itself is released only when no native call is using it
    any more, so a handler thread that finishes its request after the server was stopped can never write into whatever
    socket the kernel gave the same descriptor number next."""

    def __init__(self, fd: int):
        self._fd = fd
        self._lib = _lib()
        self.timeout: Optional[float] = None          # seconds recv waits for the START of a frame (None: forever)
        self._lens = (ctypes.c_uint64 * _MAX_SEG)()
        self._closed = False
        self._busy = 0
        self._state = threading.Lock()

    def fileno(self) -> int:
        return self._fd

    @property
    def closed(self) -> bool:
        return self._closed

    def _enter(self) -> int:
        with self._state:
            if self._closed:
                raise OSError("connection is closed")
            self._busy += 1
            return self._fd

    def _exit(self) -> None:
        with self._state:
            self._busy -= 1
            if self._closed and self._busy == 0 and self._fd >= 0:
                self._lib.dtf_net_close(self._fd)
                self._fd = -1

    # -- frames ------------------------------------------------------------------------------------------------------------
    def send_segments(self, segments: List[Any]) -> None:
        """``segments``: bytes-like objects (contiguous); the kernel gathers them from where they are."""
        n = len(segments)
        if n == 0 or n > _MAX_SEG:
            raise ValueError("a frame has 1..%d segments, got %d" % (_MAX_SEG, n))
        ptrs = (ctypes.c_void_p * n)()
        lens = (ctypes.c_uint64 * n)()
        keep = []
        for i, s in enumerate(segments):
            a = np.frombuffer(s, dtype=np.uint8)          # zero-copy view of any contiguous buffer (read-only ones too)
            keep.append(a)
            lens[i] = a.size
            ptrs[i] = a.ctypes.data if a.size else None
        fd = self._enter()
        try:
            rc = self._lib.dtf_net_send(fd, ptrs, lens, n)
        finally:
            self._exit()
        del keep
        if rc != _OK:
            _raise(rc, "send")

    def recv_segments(self, timeout: Optional[float] = None) -> List[memoryview]:
        """Next frame; every segment in a writable, aligned buffer of its own."""
        t = self.timeout if timeout is None else timeout
        fd = self._enter()
        try:
            n = self._lib.dtf_net_recv_header(fd, self._lens, _MAX_SEG, -1.0 if t is None else float(t))
            if n <= 0:
                _raise(n if n < 0 else _BAD, "recv")
            bufs = [np.empty(int(self._lens[i]), dtype=np.uint8) for i in range(n)]
            ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data if b.size else None for b in bufs])
            rc = self._lib.dtf_net_recv_body(fd, ptrs, self._lens, n)
        finally:
            self._exit()
        if rc != _OK:
            _raise(rc, "recv")
        return [memoryview(b) for b in bufs]

    # -- objects: restricted pickle envelope + out-of-band tensor buffers -----------------------------------------------------
    def send_message(self, obj: Any) -> None:
        oob: List[pickle.PickleBuffer] = []
        env = pickle.dumps(obj, protocol=5, buffer_callback=oob.append)
        if len(oob) >= _MAX_SEG:                  # thousands of tensors in one message: keep them inside the envelope
            oob, env = [], pickle.dumps(obj, protocol=5)
        self.send_segments([env] + [b.raw() for b in oob])

    def recv_message(self, loads: Callable[..., Any], timeout: Optional[float] = None) -> Any:
        segs = self.recv_segments(timeout)
        return loads(segs[0], segs[1:])

    # -- the two calls multiprocessing.connection's challenge / response uses -------------------------------------------------
    def send_bytes(self, data) -> None:
        self.send_segments([bytes(data)])

    def recv_bytes(self, maxlength: Optional[int] = None) -> bytes:
        segs = self.recv_segments()
        if len(segs) != 1 or (maxlength is not None and segs[0].nbytes > maxlength):
            self.close()
            raise OSError("bad message length")
        return segs[0].tobytes()

    def peer_closed(self) -> bool:
        with self._state:
            if self._closed or self._fd < 0:
                return True
            return bool(self._lib.dtf_net_peer_closed(self._fd))

    def close(self) -> None:
        with self._state:
            if self._closed:
                return
            self._closed = True
            if self._fd >= 0:
                self._lib.dtf_net_shutdown(self._fd)      # wakes a thread blocked in recv on this connection
                if self._busy == 0:
                    self._lib.dtf_net_close(self._fd)
                    self._fd = -1

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass


class NativeListener:
    def __init__(self, host: str, port: int, backlog: int = 64):
        lib = _lib()
        fd = lib.dtf_net_listen(host.encode(), int(port), int(backlog))
        if fd < 0:
            raise OSError(-fd, "cannot listen on %s:%d: %s" % (host, port, os.strerror(-fd)))
        self._fd, self._lib = fd, lib
        self.port = lib.dtf_net_local_port(fd)
        self._closed = False

    def accept(self, timeout: float = 0.2) -> Optional[NativeConnection]:
        """The next connection, or None after ``timeout`` seconds (lets the accept loop notice a shutdown)."""
        fd = self._lib.dtf_net_accept(self._fd, float(timeout))
        if fd == _TIMEOUT:
            return None
        if fd < 0:
            if self._closed:
                return None
            raise OSError("accept failed (code %d)" % fd)
        return NativeConnection(fd)

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._lib.dtf_net_close(self._fd)


def connect(host: str, port: int, timeout: float = 5.0) -> NativeConnection:
    fd = _lib().dtf_net_connect(host.encode(), int(port), float(timeout))
    if fd < 0:
        e = -fd
        if e == errno.ECONNREFUSED:
            raise ConnectionRefusedError(e, "connection to %s:%d refused" % (host, port))
        raise OSError(e, "cannot connect to %s:%d: %s" % (host, port, os.strerror(e)))
    return NativeConnection(fd)
