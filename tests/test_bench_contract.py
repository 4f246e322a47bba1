"""bench.py's output contract, as far as it can be exercised without a GPU: the reference arm (the oracle port of the
reference algorithm on the host cores) prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("denoise steps/sec") and line["value"] > 0 and line["n_gpus"] == 1
    assert line["steps"] == 1 and line["warmup"] == 0 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "L_v=4400" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_non_reporting_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_flop_model_of_the_bench_matches_the_oracle():
    sys.path.insert(0, ROOT)
    import bench
    import mova_oracle as O

    assert abs(bench.flops_forward(bench.FULL_360P) - O.flops_forward(bench.FULL_360P)) <= 1e-6 * O.flops_forward(bench.FULL_360P)
    assert abs(2 * bench.flops_forward(bench.FULL_360P) / 1e12 - 5532.4) < 6.0  # 5532.4 TFLOP per CFG step
