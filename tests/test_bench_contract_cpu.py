"""CPU: the reference arm of bench.py (`--impl reference`: the reference's CPU implementation of the step, oracle port)
prints ONE JSON line carrying the contract's keys - the arm the driver runs first on every box.  Config 1 (batch 128) so the
whole run is a few seconds; nothing here needs a GPU or the CUDA library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "train samples/sec (fwd+bwd)" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["n_gpus"] == 1
    assert line["value"] > 0 and abs(line["value"] - line["cpu_baseline"]["value"]) < 1e-9 * line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("cfg1") and "model" not in line["config"]
    # the batch it times is the batch it prints
    assert line["config"]["batch_per_gpu"] == 128 and "128" in line["cpu_baseline"]["sample"]
