"""CPU: host-side pieces of bench.py that the driver's contract depends on (clock sampling / throttle-reason parsing, peak
lookup, argument defaults).  The timed paths themselves need a GPU."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench_under_test"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_parses_nvidia_smi_rows():
    b = _bench()
    s = b.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.rows = ["1965, 1965, 650.1, Not Active, Not Active, Not Active, Active",
              "1800, 1965, 900.0, Not Active, Not Active, Not Active, Active",
              "1700, 1965, 910.0, Not Active, Not Active, Not Active, Not Active",
              "garbage line",
              "[N/A], 1965, 1, Not Active, Not Active, Not Active, Not Active"]
    out = s.stop()
    assert out["sm_mhz"] == 1800 and out["sm_max_mhz"] == 1965 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]
    s.rows = ["1965, 1965, 650.1, Active, Not Active, Active, Not Active"]
    assert s.stop()["reasons"] == ["hw_slowdown", "sw_thermal_slowdown"]


def test_clock_sampler_without_nvidia_smi_reports_it():
    b = _bench()
    s = b.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_peaks_come_from_the_driver_file_or_the_documented_fallback():
    b = _bench()
    peaks, source = b._peaks()
    assert peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops"] > 500
    assert source.startswith("measured") == os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))


def test_workload_is_baseline_cfg2():
    b = _bench()
    assert b.T_TOTAL == 768 and "cfg2" in b.WORKLOAD and b.METRIC == "Aria-25.3B bf16 prefill tokens/sec"
