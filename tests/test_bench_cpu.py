"""bench.py's CPU arm (`--impl reference`): one JSON line with the contract's keys on stdout, nothing else — for the
headline workload (configs[3]) and for the daily one (configs[2]).  Tiny element samples: seconds on any host."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config, window", [(3, 48), (2, 20)])
def test_reference_arm_prints_one_contract_line(config, window):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(config), "--steps", "1",
           "--warmup", "0", "--cpu-elements", "64", "--window", str(window)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "zone-timesteps/s"
    assert line["metric"].startswith("zone-timesteps/sec") and line["value"] > 0.0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert f"configs[{config}]" in line["config"]["workload"] and line["config"]["window_steps"] == window
    assert ("daily" if config == 2 else "hourly") in line["cpu_baseline"]["sample"]
