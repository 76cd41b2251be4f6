"""bench.py's clock sampler (NVML polled in-process during the timed region) against a mocked NVML: samples inside the
region, nearest samples when the region is shorter than the polling period, throttle-reason decoding, and the
no-NVML / no-nvidia-smi fallback never raising."""
import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture
def fake_nvml(monkeypatch):
    import pynvml
    state = {"reasons": 0, "mhz": 1965}
    monkeypatch.setattr(pynvml, "nvmlInit", lambda: None)
    monkeypatch.setattr(pynvml, "nvmlDeviceGetHandleByIndex", lambda i: ("h", i))
    monkeypatch.setattr(pynvml, "nvmlDeviceGetHandleByUUID", lambda u: ("h", u))
    monkeypatch.setattr(pynvml, "nvmlDeviceGetClockInfo", lambda h, c: state["mhz"])
    monkeypatch.setattr(pynvml, "nvmlDeviceGetMaxClockInfo", lambda h, c: 1965)
    monkeypatch.setattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", lambda h: state["reasons"], raising=False)
    monkeypatch.setattr(pynvml, "nvmlDeviceGetPowerUsage", lambda h: 612000)
    return state


def test_sampler_reports_median_clock_and_reasons(fake_nvml):
    import bench
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.02)
    t0 = time.time()
    fake_nvml["reasons"] = 0x4 | 0x40                      # sw_power_cap + hw_thermal_slowdown during the region
    time.sleep(0.06)
    t1 = time.time()
    out = s.stop(t0, t1)
    assert out["source"] == "nvml" and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["samples"] >= 5 and out["power_w_max"] == pytest.approx(612.0)
    assert set(out["reasons"]) == {"sw_power_cap", "hw_thermal_slowdown"}


def test_sampler_uses_nearest_samples_for_a_very_short_region(fake_nvml):
    import bench
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.03)
    t0 = time.time()
    out = s.stop(t0, t0 + 1e-6)
    assert out["sm_mhz"] == 1965.0 and out["reasons"] == [] and "nearest samples" in out.get("note", "")


def test_sampler_without_nvml_or_nvidia_smi_degrades_quietly(monkeypatch):
    import pynvml

    import bench

    def boom():
        raise RuntimeError("no NVML here")
    monkeypatch.setattr(pynvml, "nvmlInit", boom)
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop(time.time() - 0.1, time.time())
    assert out["sm_mhz"] is None and out["reasons"]


def test_reference_arm_reports_unavailable(capsys, monkeypatch):
    import json

    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2"])
    assert bench.main() == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line and line["n_gpus"] == 2


def test_greedy_ps_placement_balances_resnet18_bytes():
    """bench.py --model resnet18 places one ps shard per GPU with GreedyLoadBalancingStrategy semantics: creation order, least
    loaded task; 8 shards of ResNet-18 end up within 2.2x of the ideal share (round robin: the four 2.36 M-parameter stage-3
    filters would pile onto two tasks)."""
    import bench
    from distributed_tensorflow_b200.models import resnet18_param_shapes
    shapes = resnet18_param_shapes(10, "cifar")
    sh = bench.greedy_shards(shapes, 8)
    assert len(sh) == len(shapes) and sh[0] == 0 and set(sh) == set(range(8))

    def loads(assign):
        out = [0] * 8
        for (_, shp), t in zip(shapes, assign):
            n = 1
            for d in shp:
                n *= d
            out[t] += n
        return out
    total = sum(loads(sh))
    assert max(loads(sh)) <= 2.2 * total / 8
    rr = [i % 8 for i in range(len(shapes))]
    assert max(loads(sh)) <= max(loads(rr))
    assert bench.greedy_shards([("a", (4,)), ("b", (2,)), ("c", (1,)), ("d", (1,))], 2) == [0, 1, 1, 1]


def test_split_k_and_implicit_conv_eligibility_rules():
    from distributed_tensorflow_b200.ops import cuda_lib
    # conv weight gradient of ResNet stage 0: 5 output tiles, K = 65536 -> one wave of CTAs
    assert cuda_lib.auto_splits(576, 64, 65536, True, False) == 29
    # compute-sized and tiny problems are left alone
    assert cuda_lib.auto_splits(65536, 64, 576, True, False) == 1 and cuda_lib.auto_splits(100, 100, 784, True, False) == 1
    # few tiles, moderate K: at least 8 K blocks per split, atomic traffic bounded
    assert cuda_lib.auto_splits(1024, 512, 4608, True, False) == 9
    assert cuda_lib.auto_splits(4608, 512, 1024, True, False) == 2
    ok = cuda_lib.implicit_conv_ok
    assert ok(64, 32, 32, 64) and ok(64, 16, 16, 128) and ok(64, 8, 8, 256) and ok(64, 4, 4, 512) and ok(8, 4, 4, 512)
    assert not ok(64, 32, 32, 3) and not ok(64, 32, 32, 64, (2, 2)) and not ok(2, 4, 4, 512) and not ok(4, 24, 24, 64)
