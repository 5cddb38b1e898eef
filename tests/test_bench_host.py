"""Host-side logic of bench.py that needs no GPU: the algorithmic-bytes formula of SURVEY 8(d), the workload table and the
clock sampler (driven by a fake NVML: fast queries, and queries slower than a timed step)."""
import importlib.util
import os
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_per_update_matches_survey(bench):
    # 28.99 B/update at (256, 64), 29.34 at (128, 50) (+ cardinality in/out for CPHD: 29.66), 28.78 at (128, 100)
    assert abs(bench.alg_bytes_per_update(256, 64) - 474912 / (256 * 64)) < 1e-12
    assert round(bench.alg_bytes_per_update(256, 64), 2) == 28.99
    assert round(bench.alg_bytes_per_update(128, 50), 2) == 29.34
    assert round(bench.alg_bytes_per_update(128, 50, 256), 2) == 29.66
    assert round(bench.alg_bytes_per_update(128, 100), 2) == 28.78


def test_default_workload_is_the_headline_configuration(bench):
    wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (wl["P"], wl["C"], wl["M"]) == (65536, 256, 64) and wl.get("filter_type", 0) == 0
    cphd = bench.WORKLOADS["synthetic_131072x128x50_cphd"]
    assert cphd["filter_type"] == 1 and cphd["P"] * 8 == 1048576          # BASELINE configs[3] over 8 GPUs


class _FakeNvml(object):
    NVML_CLOCK_SM = 1
    nvmlClocksThrottleReasonHwSlowdown = 8
    nvmlClocksThrottleReasonHwThermalSlowdown = 64
    nvmlClocksThrottleReasonSwThermalSlowdown = 32
    nvmlClocksThrottleReasonSwPowerCap = 4
    nvmlClocksThrottleReasonHwPowerBrakeSlowdown = 128

    def __init__(self, delay, reasons):
        self.delay, self.reasons = delay, reasons

    def nvmlDeviceGetMaxClockInfo(self, h, c):
        return 1965

    def nvmlDeviceGetClockInfo(self, h, c):
        time.sleep(self.delay)
        return 1900

    def nvmlDeviceGetCurrentClocksEventReasons(self, h):
        return self.reasons


def _run_sampler(bench, delay, reasons):
    s = bench.ClockSampler(0)
    s._nvml_handle = lambda: (_FakeNvml(delay, reasons), object())
    s.start()
    for _ in range(5):
        time.sleep(0.01)             # state restore between steps
        s.active = True
        time.sleep(0.013)            # one timed step
        s.active = False
    return s.stop()


def test_clock_sampler_reports_samples_inside_the_timed_steps(bench):
    c = _run_sampler(bench, 0.0005, 4)
    assert c["sm_mhz"] == 1900.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    assert c["samples"] == c["samples_inside_timed_steps"] >= 3


def test_clock_sampler_survives_queries_slower_than_a_step(bench):
    # a 30 ms query never ENDS inside a 13 ms step; samples are tagged at query start and the loop's samples are the fallback
    c = _run_sampler(bench, 0.03, 64)
    assert c["sm_mhz"] == 1900.0 and c["samples"] >= 1
    assert c["reasons"] == ["hw_thermal_slowdown"]


def test_text_config_runner_oracle_arm():
    """profiles/text_configs.py (BASELINE configs[0] / configs[1] in full) runs its CPU arm on the bundled text data"""
    import json
    import subprocess
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "profiles", "text_configs.py"), "--impl", "oracle",
                                   "--max-steps", "4", "--threads", "2"], text=True, timeout=300)
    lines = [json.loads(l) for l in out.strip().split("\n")]
    assert [l["particles"] for l in lines] == [256, 4096] and all(l["steps"] == 4 and l["ms_per_step_mean"] > 0 for l in lines)


def test_python_scene_config_is_byte_identical_to_the_library_one(bench):
    """bench.py's --impl reference arm builds its config without libphdslam.so (scene_config_py): same bytes"""
    from phdslam_b200 import scene as S
    for name, wl in bench.WORKLOADS.items():
        extra = {k: v for k, v in wl.items() if k not in bench.NON_CFG_KEYS}
        a = S.scene_config(wl["P"], wl["C"], wl["M"], max_components=wl["max_components"], **extra)
        b = S.scene_config_py(wl["P"], wl["C"], wl["M"], max_components=wl["max_components"], **extra)
        assert bytes(a) == bytes(b), name


def test_reference_arm_does_not_load_the_product_library():
    """`bench.py --impl reference` times the CPU oracle only: libphdslam.so must not be mapped into that process
    (the arm asserts it itself on /proc/self/maps; here the whole command is run on a small workload)"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "synthetic_1024x64x32_phd"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["single_thread_value"] > 0
    assert line["config"]["particles_per_gpu"] == 1024
