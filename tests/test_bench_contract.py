"""bench.py contract checks that run without a GPU: the reference arm (the oracle port on the host cores) prints ONE JSON line
with the keys the driver reads; the algorithmic-work formulas match SURVEY.md §8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--frames", "50"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("motion-latent frames/s") and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_work_matches_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    from oracle.synth import FmtDims
    w = bench.algorithmic_work(FmtDims(), 1, 3, 9)
    # 203.4 MB of step weights + 18.43 MB of AdaLN table rows per clip and step; 155 196 416 MACs per token row
    assert abs(w["step_bytes"] - (203.4e6 + 18.43e6)) / 221.8e6 < 2e-3
    assert w["rows"] == 180
    macs_row = (w["window_flops"] / 9 / 180 - 4 * 60 * 1024) / 2
    assert macs_row == 155_196_416
