"""Host-side logic of bench.py that needs no GPU: the clock sampler keeps only the nvidia-smi lines that arrived while
the GPU was under the workload's load, and summarises them (median SM clock, max clock, active throttle reasons)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


class _Proc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass


def _sampler(rows):
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.rows = rows
    return s


def test_clock_sampler_window_and_reasons():
    line = "0, {sm}, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, {cap}"
    rows = [(0.5, line.format(sm=210, cap="Not Active")),          # idle clock before the timed region: ignored
            (1.1, line.format(sm=1965, cap="Not Active")),
            (1.2, line.format(sm=1950, cap="Active")),
            (1.3, line.format(sm=1965, cap="Not Active")),
            (2.5, line.format(sm=300, cap="Not Active")),          # after the load: ignored
            (1.25, "garbage line")]
    s = _sampler(rows)
    assert s.count(1.0, 2.0) == 4                                  # three samples + the malformed line
    out = s.stop(1.0, 2.0)
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert _sampler(rows).stop()["samples"] == 5                   # no window: everything that parses


def test_clock_sampler_without_nvidia_smi():
    s = bench.ClockSampler(0)
    assert s.stop(0.0, 1.0)["reasons"] == ["nvidia-smi unavailable"]


def test_workload_registry_matches_the_documented_names():
    assert sorted(bench.WORKLOADS) == sorted(["ukfom", "usckf", "msckf", "fusion", "ekf", "msckf_ekf", "safefusion", "deadreckon"])
    for name, cls in bench.WORKLOADS.items():
        assert cls.name == name and cls.unit and cls.metric
