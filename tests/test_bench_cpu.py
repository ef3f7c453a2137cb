"""CPU checks of bench.py's host logic: both arms describe the workload with the same `config`
object, the clock sampler degrades to a record that says so when no GPU is there, and the
reference arm prints the contract line on a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config_is_shared_by_both_arms():
    for w in bench.WORKLOADS:
        c1, c8 = bench.config_of(w, 1), bench.config_of(w, 8)
        assert c1["workload"] == c8["workload"] and "model" not in c1
        assert c1["n"] == bench.WORKLOADS[w][1] and c1["d"] == bench.WORKLOADS[w][2]
        assert c8["rows_per_gpu"] * 8 >= c8["n"]


def test_clock_sampler_without_a_gpu_reports_that():
    s = bench.ClockSampler(0)
    s.start()
    rec = s.stop()
    assert set(rec) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if rec["sm_mhz"] is None:           # no NVML device, no nvidia-smi: the record says so
        assert rec.get("samples", 0) == 0 or rec["reasons"]


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "sift100k", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=580)
    assert p.returncode == 0, p.stderr[-400:]
    line = json.loads([x for x in p.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "exact kNN queries/sec"
    assert line["config"] == bench.config_of("sift100k", 1)
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
