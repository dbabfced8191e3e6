"""The reference arm of bench.py runs on the host cores (no GPU): check its JSON line against the
contract keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "config1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "region-days/sec" and line["unit"] == "region-days/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0
