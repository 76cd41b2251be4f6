"""Control-plane RPC between tasks (the role gRPC plays under TF).

One TCP listener per task (``multiprocessing.connection``: length-prefixed
pickles), one handler thread per connection so blocking calls (token dequeue,
``take_grad``) never stall other clients.  Tensors cross as CPU tensors.

This path carries *control* and small tensors.  On B200 the bulk traffic
(parameter pull, gradient push) does not use it: it goes through NVLink peer
memory from inside the kernels (``parallel/fabric.py`` + ``ops/``).
"""
from __future__ import annotations

import hashlib
import io
import os
import pickle
import select
import socket
import threading
import time
import traceback
import weakref
from multiprocessing import AuthenticationError
from multiprocessing.connection import Client, Connection, Listener, answer_challenge, deliver_challenge
from typing import Any, Callable, Dict, Optional, Tuple

import torch

from ..framework import errors
from . import transport as _transport

__all__ = ["RpcServer", "RpcClient", "parse_address", "to_wire", "from_wire", "PeerAwareCancel", "current_connection",
           "peer_closed"]

_LOOPBACK = ("127.0.0.1", "::1", "localhost")
_HANDSHAKE_TIMEOUT = 10.0


def cluster_authkey(host: str) -> bytes:
    """The shared secret both ends of a control-plane connection must prove (HMAC challenge, both directions).

    * ``DTF_CLUSTER_SECRET`` (the secret itself) or ``DTF_CLUSTER_SECRET_FILE`` (a file holding it): REQUIRED as soon as a
      task binds or dials a non-loopback address -- a multi-host cluster ships the same secret to every task, like an ssh
      key; without one such a bind is refused (anyone who can reach the port could otherwise drive the task).
    * loopback only (the reference's localhost layout, every test): a per-user random secret created on first use in
      ``~/.dtf_b200/cluster_secret`` (mode 0600), so other local users cannot connect either."""
    env = os.environ.get("DTF_CLUSTER_SECRET")
    if env:
        return hashlib.sha256(env.encode()).digest()
    path = os.environ.get("DTF_CLUSTER_SECRET_FILE")
    if path:
        with open(path, "rb") as f:
            return hashlib.sha256(f.read().strip()).digest()
    if host not in _LOOPBACK:
        raise PermissionError("refusing a control-plane endpoint on non-loopback address %r without a cluster secret: set "
                              "DTF_CLUSTER_SECRET or DTF_CLUSTER_SECRET_FILE to the same value on every task" % host)
    d = os.path.join(os.path.expanduser("~"), ".dtf_b200")
    f = os.path.join(d, "cluster_secret")
    try:
        with open(f, "rb") as fh:
            key = fh.read()
        if len(key) >= 32:
            return key[:64]
    except OSError:
        pass
    os.makedirs(d, mode=0o700, exist_ok=True)
    key = os.urandom(48)
    tmp = "%s.%d" % (f, os.getpid())
    fd = os.open(tmp, os.O_WRONLY | os.O_CREAT | os.O_EXCL, 0o600)
    with os.fdopen(fd, "wb") as fh:
        fh.write(key)
    try:
        os.link(tmp, f)                      # first writer wins; everybody then reads the same file
    except OSError:
        pass
    os.unlink(tmp)
    with open(f, "rb") as fh:
        return fh.read()[:64]


# ---- wire format: pickle, but only of a closed set of types ---------------------------------------------------------------
_SAFE_BUILTINS = {"set", "frozenset", "slice", "range", "complex", "bytearray", "tuple", "list", "dict", "int", "float", "str",
                  "bytes", "bool", "object", "NoneType", "Ellipsis"}
_SAFE_MODULE_PREFIXES = ("numpy", "torch", "collections", "distributed_tensorflow_b200")
_DENY_NAMES = {"eval", "exec", "compile", "open", "__import__", "getattr", "setattr", "delattr", "system", "popen", "Popen",
               "load", "loads", "load_state_dict", "run", "call", "check_output", "spawn", "fork", "execv", "execve"}


class _RestrictedUnpickler(pickle.Unpickler):
    """Requests and replies carry python scalars / containers, numpy arrays (tensors), torch dtypes and this package's own
    small value classes -- nothing else is allowed to be named by a pickle arriving over the network (the classic
    ``os.system`` / ``subprocess`` gadgets are not importable through it)."""

    def find_class(self, module: str, name: str):
        root = module.split(".")[0]
        ok = (module == "builtins" and name in _SAFE_BUILTINS) or \
             (root in _SAFE_MODULE_PREFIXES and name.split(".")[-1] not in _DENY_NAMES and not name.startswith("_import"))
        if not ok:
            raise pickle.UnpicklingError("control-plane message names %s.%s, which is not an allowed wire type" % (module, name))
        return super().find_class(module, name)


def _loads(data, buffers=None) -> Any:
    """``buffers``: the out-of-band tensor segments of a native-transport frame (pickle protocol 5)."""
    return _RestrictedUnpickler(io.BytesIO(data), buffers=buffers).load()


def _send_obj(conn, obj: Any) -> None:
    """One message: a native-transport frame (envelope + one segment per tensor, gathered from the tensors' memory) or a
    length-prefixed pickle on a ``multiprocessing.connection`` connection."""
    if isinstance(conn, _transport.NativeConnection):
        conn.send_message(obj)
    else:
        conn.send_bytes(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))


def _recv_obj(conn) -> Any:
    if isinstance(conn, _transport.NativeConnection):
        return conn.recv_message(_loads)
    return _loads(conn.recv_bytes())

_ERRORS = {c.__name__: c for c in (
    errors.OpError, errors.FailedPreconditionError, errors.AbortedError, errors.UnavailableError,
    errors.OutOfRangeError, errors.CancelledError, errors.DeadlineExceededError, errors.NotFoundError,
    errors.InvalidArgumentError, ValueError, KeyError, TypeError, RuntimeError, NotImplementedError)}


def parse_address(addr: str) -> Tuple[str, int]:
    addr = addr.strip()
    for prefix in ("grpc://", "dtf://", "tcp://"):
        if addr.startswith(prefix):
            addr = addr[len(prefix):]
    host, _, port = addr.rpartition(":")
    if not host or host in ("localhost", "0.0.0.0"):
        host = "127.0.0.1"
    return host, int(port)


_TENSOR_TAG = "__dtf_tensor__"
_VIEW_AS = {torch.bfloat16: torch.int16}            # dtypes numpy cannot represent travel as same-width integers


def to_wire(value: Any) -> Any:
    """Detach + move tensors to host memory, recursively, and encode them as ``(tag, dtype, ndarray)``: a numpy array
    pickles with one memcpy (protocol 5), a torch tensor goes through ``torch.save`` machinery (~10x slower for the
    small tensors of a parameter-server step -- measured 116 us vs 10 us for a 10-element vector)."""
    if isinstance(value, torch.Tensor):
        t = value.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        if not t.is_contiguous():
            t = t.contiguous()
        carrier = _VIEW_AS.get(t.dtype)
        if carrier is not None:
            return (_TENSOR_TAG, str(t.dtype), t.view(carrier).numpy())
        try:
            return (_TENSOR_TAG, None, t.numpy())
        except (TypeError, RuntimeError):            # exotic dtype: let torch pickle it
            return t
    if isinstance(value, dict):
        return {k: to_wire(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        if len(value) == 3 and isinstance(value[0], str) and value[0] == _TENSOR_TAG:
            return value                                 # already encoded
        return type(value)(to_wire(v) for v in value)
    return value


def from_wire(value: Any) -> Any:
    """Inverse of :func:`to_wire` on the receiving side."""
    if isinstance(value, tuple):
        if len(value) == 3 and isinstance(value[0], str) and value[0] == _TENSOR_TAG:
            t = torch.from_numpy(value[2])
            return t if value[1] is None else t.view(getattr(torch, value[1].split(".")[-1]))
        return tuple(from_wire(v) for v in value)
    if isinstance(value, list):
        return [from_wire(v) for v in value]
    if isinstance(value, dict):
        return {k: from_wire(v) for k, v in value.items()}
    return value


_current = threading.local()          # the connection whose request the calling thread is serving


def peer_closed(conn: Optional[Connection]) -> bool:
    """True once the client end of ``conn`` is gone (process killed, socket closed): the kernel reports a hang-up
    on the descriptor even while a handler thread is blocked inside a long-running request."""
    if conn is None:
        return False
    if isinstance(conn, _transport.NativeConnection):
        return conn.peer_closed()
    try:
        p = select.poll()
        p.register(conn.fileno(), select.POLLRDHUP | select.POLLHUP | select.POLLERR)
        return any(ev & (select.POLLRDHUP | select.POLLHUP | select.POLLERR | select.POLLNVAL) for _, ev in p.poll(0))
    except (OSError, ValueError):
        return True


class PeerAwareCancel:
    """Duck-typed ``threading.Event`` handed to blocking kernels (token dequeue, ``take_grad``) that run on behalf
    of a remote client: it fires when the session is cancelled OR the requesting client has died.  Without it a
    worker killed while it waits for a sync token would still consume the token when it arrives -- the reply goes
    nowhere, the token is lost and the surviving replicas starve (backup-worker tolerance)."""

    def __init__(self, event: threading.Event, conn: Optional[Connection]):
        self._event, self._conn = event, conn

    def is_set(self) -> bool:
        return self._event.is_set() or peer_closed(self._conn)

    def wait(self, timeout: Optional[float] = None) -> bool:
        if self._event.wait(timeout):
            return True
        return peer_closed(self._conn)


def current_connection() -> Optional[Connection]:
    """The RPC connection being served by this thread (None for in-process calls)."""
    return getattr(_current, "conn", None)


class RpcServer:
    def __init__(self, address: str, service: Any):
        self.host, self.port = parse_address(address)
        self._service = service
        self._authkey = cluster_authkey(self.host)
        self.native = _transport.available()
        # handshake: per connection, below
        self._listener = _transport.NativeListener(self.host, self.port, backlog=64) if self.native else \
            Listener((self.host, self.port), backlog=64)
        self._closed = threading.Event()
        self._conns = []
        self._serving = []               # connection threads (joined by close(): none of them may outlive the interpreter)
        self._thread = threading.Thread(target=self._accept_loop, name="dtf-rpc-accept-%d" % self.port, daemon=True)
        self._thread.start()

    def _accept_loop(self) -> None:
        while not self._closed.is_set():
            try:
                conn = self._listener.accept()
            except (OSError, EOFError, Exception):
                if self._closed.is_set():
                    return
                continue
            if conn is None:                 # native accept woke up on its timeout: look at the closed flag again
                continue
            self._conns = [c for c in self._conns if not getattr(c, "closed", False)]
            self._conns.append(conn)
            self._serving = [t for t in self._serving if t.is_alive()]
            try:
                t = threading.Thread(target=self._serve, args=(conn,), name="dtf-rpc-conn", daemon=True)
                t.start()
                self._serving.append(t)
            except RuntimeError:             # the interpreter is shutting down: no new threads
                conn.close()
                return

    def _handshake(self, conn: Connection) -> bool:
        """Mutual HMAC challenge in THIS connection's thread (the accept loop never waits on a client), bounded by a
        watchdog that closes a connection which does not complete it in time."""
        dog = threading.Timer(_HANDSHAKE_TIMEOUT, conn.close)
        dog.daemon = True
        dog.start()
        try:
            deliver_challenge(conn, self._authkey)
            answer_challenge(conn, self._authkey)
            return True
        except (AuthenticationError, EOFError, OSError, ValueError, TypeError, ConnectionError):
            return False
        finally:
            dog.cancel()

    def _serve(self, conn: Connection) -> None:
        try:
            if not self._handshake(conn):
                return
            while not self._closed.is_set():
                try:
                    method, args, kwargs = from_wire(_recv_obj(conn))
                except (EOFError, OSError, ConnectionError, TypeError, ValueError, pickle.UnpicklingError):
                    return               # peer gone, this server closed the connection, or a message outside the wire types
                _current.conn = conn
                try:
                    fn = getattr(self._service, "rpc_" + method)
                    result = ("ok", to_wire(fn(*args, **kwargs)))
                except BaseException as e:  # noqa: BLE001 - errors travel to the caller
                    result = ("err", type(e).__name__, str(e), traceback.format_exc())
                finally:
                    _current.conn = None
                try:
                    _send_obj(conn, result)
                except (OSError, ConnectionError, BrokenPipeError, EOFError):
                    return
        finally:
            try:
                conn.close()
            except OSError:
                pass

    def close(self) -> None:
        self._closed.set()
        # wake the accept() call first: the port is only released once no thread is blocked on the socket
        try:
            s = socket.create_connection((self.host, self.port), timeout=0.5)
            s.close()
        except OSError:
            pass
        self._thread.join(2.0)
        try:
            self._listener.close()
        except OSError:
            pass
        for c in self._conns:
            try:
                c.close()
            except OSError:
                pass
        # the connection threads were blocked in a receive: closing their connection wakes them; wait for them (bounded) so
        # that no thread of this server is inside a native call when the interpreter finalises
        deadline = time.time() + 2.0
        me = threading.current_thread()
        for t in self._serving:
            if t is not me:
                t.join(max(0.0, deadline - time.time()))
        self._serving = [t for t in self._serving if t.is_alive()]


_ALL_CLIENTS: "weakref.WeakSet" = weakref.WeakSet()


def close_all_clients() -> None:
    """Close every live RpcClient of this process (exit path: a session the script never closed must not make its own
    in-process servers wait for it)."""
    for c in list(_ALL_CLIENTS):
        try:
            c.close()
        except Exception:      # noqa: BLE001 - exit path
            pass


class RpcClient:
    """Client stub; one connection per calling thread (blocking calls do not serialise threads)."""

    def __init__(self, address: str, connect_timeout: float = 30.0):
        _ALL_CLIENTS.add(self)
        self.address = address
        self.host, self.port = parse_address(address)
        self._tls = threading.local()
        self._timeout = connect_timeout
        self._all = []
        self._lock = threading.Lock()

    def _conn(self) -> Connection:
        c = getattr(self._tls, "conn", None)
        if c is None:
            deadline = time.time() + self._timeout
            delay = 0.02
            while True:
                try:
                    c = _transport.connect(self.host, self.port, timeout=5.0) if _transport.available() else \
                        Client((self.host, self.port))
                    try:
                        key = cluster_authkey(self.host)
                        answer_challenge(c, key)
                        deliver_challenge(c, key)
                    except (AuthenticationError, EOFError, ConnectionError) as e:
                        c.close()
                        raise errors.UnavailableError("task at %s:%d rejected the cluster secret (%s): every task needs the same "
                                                      "DTF_CLUSTER_SECRET" % (self.host, self.port, e))
                    break
                except (ConnectionRefusedError, FileNotFoundError, OSError) as e:
                    if time.time() > deadline:
                        raise errors.UnavailableError("cannot reach task at %s:%d: %s" % (self.host, self.port, e))
                    time.sleep(delay)
                    delay = min(delay * 1.5, 0.5)
            self._tls.conn = c
            with self._lock:
                self._all.append(c)
        return c

    def call(self, method: str, *args, **kwargs) -> Any:
        try:
            c = self._conn()
            _send_obj(c, (method, to_wire(args), to_wire(kwargs)))
            reply = from_wire(_recv_obj(c))
        except (EOFError, ConnectionError, BrokenPipeError, OSError) as e:
            self._drop()
            raise errors.UnavailableError("task at %s:%d went away during %s: %s" % (self.host, self.port, method, e))
        if reply[0] == "ok":
            return reply[1]
        _, ename, msg, tb = reply
        exc = _ERRORS.get(ename, RuntimeError)
        raise exc("%s\n--- remote traceback (%s:%d) ---\n%s" % (msg, self.host, self.port, tb))

    def try_connect(self, timeout: float = 0.5) -> bool:
        try:
            s = socket.create_connection((self.host, self.port), timeout=timeout)
            s.close()
            return True
        except OSError:
            return False

    def _drop(self) -> None:
        c = getattr(self._tls, "conn", None)
        if c is not None:
            try:
                c.close()
            except OSError:
                pass
            self._tls.conn = None

    def close(self) -> None:
        with self._lock:
            for c in self._all:
                try:
                    c.close()
                except OSError:
                    pass
            self._all.clear()
        self._tls = threading.local()
