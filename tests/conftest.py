import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >=2 GPUs")


def free_ports(n):
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        ports.append(s.getsockname()[1])
        socks.append(s)
    for s in socks:
        s.close()
    return ports


@pytest.fixture
def ports():
    return free_ports


@pytest.fixture(autouse=True)
def fresh_graph():
    """Every test builds into its own default graph and flag registry."""
    import distributed_tensorflow_b200 as dtf
    from distributed_tensorflow_b200.utils import flags as F
    dtf.reset_default_graph()
    F.FLAGS.reset()
    yield
    dtf.reset_default_graph()
    F.FLAGS.reset()


@pytest.fixture
def cluster3(ports):
    """1 ps + 2 workers, all served from this process; yields (cluster, servers) and stops them."""
    import distributed_tensorflow_b200 as dtf
    p = ports(3)
    cluster = dtf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0]],
                                     "worker": ["127.0.0.1:%d" % p[1], "127.0.0.1:%d" % p[2]]})
    servers = [dtf.train.Server(cluster, "ps", 0), dtf.train.Server(cluster, "worker", 0),
               dtf.train.Server(cluster, "worker", 1)]
    yield cluster, servers
    for s in servers:
        s.stop()
