"""Native control-plane transport (csrc/runtime/transport.cpp + parallel/transport.py): frames of segments, out-of-band
tensors, deadlines, hang-up detection, close() waking a blocked reader."""
import os
import pickle
import threading
import time

import numpy as np
import pytest
import torch

from distributed_tensorflow_b200.parallel import rpc, transport

pytestmark = pytest.mark.skipif(not transport.available(), reason="native runtime not built")


@pytest.fixture
def pair():
    lst = transport.NativeListener("127.0.0.1", 0)
    got = []
    t = threading.Thread(target=lambda: got.append(lst.accept(timeout=5.0)))
    t.start()
    c = transport.connect("127.0.0.1", lst.port)
    t.join(5.0)
    s = got[0]
    assert s is not None
    yield c, s
    c.close()
    s.close()
    lst.close()


def test_segments_roundtrip_without_copies_into_one_buffer(pair):
    c, s = pair
    a = np.arange(1 << 20, dtype=np.float32)
    b = bytearray(b"xyz" * 1000)
    t = threading.Thread(target=lambda: c.send_segments([b"envelope", a, b, b""]))      # 4 MB: more than the socket buffers hold
    t.start()
    segs = s.recv_segments(timeout=5.0)
    t.join(5.0)
    assert [m.nbytes for m in segs] == [8, a.nbytes, 3000, 0]
    assert bytes(segs[0]) == b"envelope" and np.array_equal(np.frombuffer(segs[1], dtype=np.float32), a) and bytes(segs[2]) == bytes(b)
    assert np.frombuffer(segs[1], dtype=np.uint8).ctypes.data % 16 == 0          # aligned landing buffers
    with pytest.raises(ValueError):
        c.send_segments([])


def test_messages_carry_tensors_out_of_band(pair):
    c, s = pair
    x = torch.randn(300, 200)
    msg = ("call", rpc.to_wire({"x": x, "h": torch.ones(5, dtype=torch.bfloat16), "nc": torch.arange(12.0).view(3, 4).t()}), 7)
    c.send_message(msg)
    segs = s.recv_segments(timeout=5.0)
    assert len(segs) >= 3 and segs[0].nbytes < 1000 and any(m.nbytes == x.numel() * 4 for m in segs[1:])
    out = rpc.from_wire(rpc._loads(segs[0], segs[1:]))
    assert out[0] == "call" and out[2] == 7 and torch.equal(out[1]["x"], x) and out[1]["h"].dtype == torch.bfloat16
    assert torch.equal(out[1]["nc"], torch.arange(12.0).view(3, 4).t())
    # the restricted unpickler still guards the envelope
    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))
    c.send_message(Evil())
    with pytest.raises(pickle.UnpicklingError):
        s.recv_message(rpc._loads, timeout=5.0)


def test_deadline_applies_to_the_start_of_a_frame_only(pair):
    c, s = pair
    t0 = time.time()
    with pytest.raises(transport.TransportTimeout):
        s.recv_segments(timeout=0.15)
    assert 0.1 < time.time() - t0 < 2.0
    c.send_bytes(b"late")                                   # the stream is still in sync after a timeout
    assert s.recv_bytes() == b"late"
    with pytest.raises(OSError):
        c.send_bytes(b"x" * 300)
        s.recv_bytes(maxlength=256)


def test_hang_up_is_visible_and_close_wakes_a_blocked_reader(pair):
    c, s = pair
    assert not s.peer_closed()
    c.close()
    deadline = time.time() + 5.0
    while not s.peer_closed() and time.time() < deadline:
        time.sleep(0.01)
    assert s.peer_closed()
    with pytest.raises(EOFError):
        s.recv_segments(timeout=2.0)
    # a reader blocked without a deadline is released by close() from another thread
    lst = transport.NativeListener("127.0.0.1", 0)
    acc = []
    t = threading.Thread(target=lambda: acc.append(lst.accept(timeout=5.0)))
    t.start()
    c2 = transport.connect("127.0.0.1", lst.port)
    t.join(5.0)
    s2, res = acc[0], []

    def reader():
        try:
            s2.recv_segments()
            res.append("data")
        except (EOFError, OSError) as e:
            res.append(type(e).__name__)
    rt = threading.Thread(target=reader)
    rt.start()
    time.sleep(0.1)
    s2.close()
    rt.join(5.0)
    assert res and res[0] in ("EOFError", "OSError", "ConnectionError")
    c2.close()
    lst.close()
    with pytest.raises((ConnectionRefusedError, OSError)):
        transport.connect("127.0.0.1", lst.port, timeout=0.5)


def test_a_task_script_that_just_returns_leaves_no_server_thread_for_interpreter_finalisation(tmp_path):
    """The reference's worker-0 client returns from ``main()`` with its in-process Server running
    (``example_distributed_server.py:46-70``).  Its accept / connection threads sit in native socket calls; a daemon thread that
    comes back from one while CPython finalises is ended with ``pthread_exit`` -- seen once as ``terminate called without an
    active exception`` (exit -6) after the script's output was complete.  The package's ``atexit`` hook closes the listeners and
    connections and joins those threads before finalisation starts."""
    import subprocess
    import sys
    script = tmp_path / "task.py"
    script.write_text('''
import atexit, socket, sys, threading
def _dump():
    print("ALIVE_AT_EXIT", [t.name for t in threading.enumerate() if t.name.startswith("dtf-") and t.is_alive()])
atexit.register(_dump)           # registered first -> runs last, after the package's hook
import distributed_tensorflow_b200 as tf
from distributed_tensorflow_b200.parallel.rpc import RpcClient
def port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p
cluster = tf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % port()], "worker": ["127.0.0.1:%d" % port()]})
ps = tf.train.Server(cluster, job_name="ps", task_index=0)
wk = tf.train.Server(cluster, job_name="worker", task_index=0)
with tf.device("/job:ps/task:0"):
    w = tf.Variable(tf.ones([8, 8]), name="w")
with tf.device("/job:worker/task:0"):
    y = tf.matmul(w, w)
client = RpcClient(cluster.task_address("ps", 0))                          # kept alive: its connection stays open until exit
print(client.call("ping")["task"])                                         # a real connection -> a connection thread on the ps
with tf.Session(wk.target) as sess:
    sess.run(tf.global_variables_initializer()); print(float(sess.run(y)[0, 0]))
print("BEFORE", sorted(t.name for t in threading.enumerate() if t.name.startswith("dtf-rpc")))
''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "dtf-rpc-conn" in r.stdout.split("BEFORE")[1].splitlines()[0] and "8.0" in r.stdout
    assert "ALIVE_AT_EXIT []" in r.stdout, r.stdout


def test_an_exiting_task_lingers_while_a_peer_session_is_still_using_it(tmp_path):
    """In-graph replication makes tasks clients of each other (the reference's example_distributed_server.py: EVERY worker runs the
    client code).  A task script that returns while a peer's session still places ops on it used to fail that peer's next run with
    UnavailableError; now its exit waits (bounded by DTF_EXIT_LINGER_S) until the peer released the session or went away."""
    import socket
    import subprocess
    import sys
    import distributed_tensorflow_b200 as tf
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "serving_task.py"
    script.write_text('''
import sys, time
import distributed_tensorflow_b200 as tf
srv = tf.train.Server(tf.train.ClusterSpec({"worker": ["127.0.0.1:%d"]}), job_name="worker", task_index=0)
print("READY", flush=True)
sys.stdin.readline()                  # the peer ran its first step: this script is done and returns
print("RETURNING %%.3f" %% time.time(), flush=True)
''' % port)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="", DTF_EXIT_LINGER_S="20")
    p = subprocess.Popen([sys.executable, str(script)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env)
    try:
        assert p.stdout.readline().startswith("READY")
        tf.reset_default_graph()
        with tf.device("/job:worker/task:0"):
            v = tf.Variable(3.0, name="v")
            bump = tf.assign_add(v, 1.0)
        sess = tf.Session("grpc://127.0.0.1:%d" % port)
        sess.run(tf.global_variables_initializer())
        assert sess.run(bump) == 4.0
        p.stdin.write("go\n")
        p.stdin.flush()
        t_ret = float(p.stdout.readline().split()[1])
        time.sleep(1.0)                               # the task has returned from its script; our session is still open
        assert p.poll() is None                       # ... and it is still there
        assert sess.run(bump) == 5.0                  # the run that used to hit a dead task
        sess.close()                                  # released: nothing keeps the task any more
        p.wait(timeout=15)
        assert p.returncode == 0 and time.time() - t_ret < 10.0
    finally:
        if p.poll() is None:
            p.kill()
        tf.reset_default_graph()
