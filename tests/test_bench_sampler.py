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
