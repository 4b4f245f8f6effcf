"""utils/timing.py on the CPU: host-clock fallbacks of DeviceTimer / StepClock, max_over_ranks over gloo, and the
nvidia-smi parser of ClockSampler (the bench's ``clocks`` record) against a fake nvidia-smi on PATH."""
import os
import stat
import time

import torch

from bert_pytorch_b200.utils import timing


def test_device_timer_and_step_clock_fall_back_to_the_host_clock():
    t = timing.DeviceTimer(torch.device("cpu"))
    t.start()
    time.sleep(0.02)
    assert 15.0 < t.stop() < 500.0
    c = timing.StepClock(torch.device("cpu"))
    c.mark(); time.sleep(0.01); c.mark(); time.sleep(0.02); c.mark()
    a, b = c.intervals()
    assert 5.0 < a < 200.0 and 15.0 < b < 400.0 and b > a
    c.reset()
    assert c.intervals() == []
    assert timing.max_over_ranks(3.5) == 3.5                      # no process group: identity
    timing.L2Flusher(torch.device("cpu")).flush()                  # no-op off CUDA
    with timing.nvtx_range("x"):
        pass


def test_clock_sampler_parses_nvidia_smi_lines(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\n"
                    "echo '0, 1650, 1965, 981.5, Not Active, Not Active, Not Active, Active'\n"
                    "echo '0, 1710, 1965, 995.0, Not Active, Not Active, Not Active, Active'\n"
                    "echo '0, [N/A], 1965, 10.0, Not Active, Not Active, Not Active, Not Active'\n"
                    "echo '0, 1680, 1965, 990.0, Not Active, Active, Not Active, Not Active'\n"
                    "sleep 30\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    s = timing.ClockSampler(0, period_ms=50)
    s.start()
    time.sleep(0.5)
    rec = s.stop()
    assert rec["samples"] == 3 and rec["sm_mhz"] == 1680.0 and rec["sm_max_mhz"] == 1965.0
    assert rec["power_w_max"] == 995.0 and rec["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]


def test_clock_sampler_without_nvidia_smi(monkeypatch, tmp_path):
    monkeypatch.setenv("PATH", str(tmp_path))
    s = timing.ClockSampler(0)
    s.start()
    rec = s.stop()
    assert rec["samples"] == 0 and rec["sm_mhz"] is None and rec["reasons"] == []
