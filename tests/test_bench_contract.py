"""bench.py's JSON contract. The reference arm (`--impl reference` = the CPU restatement, the one place besides the
cpu_baseline leg where bench.py executes oracle/) runs here without a GPU; the committed B200 line
(profiles/r1_bench_S3_n1.json) is checked for the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _baseline_metric():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)["metric"]


def test_reference_arm_prints_one_json_line_on_the_cpu():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.pop("WORLD_SIZE", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-cells", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                      # exactly one line on stdout
    line = json.loads(lines[0])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["value"] > 0 and line["unit"] == "Melements/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] >= 1 and "Kuhn tets" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and "workload" in line["config"]
    assert "CG momentum+tracer assembly" in line["metric"] and "CG momentum+tracer assembly" in _baseline_metric()


def test_non_zero_ranks_of_the_reference_arm_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-cells", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_committed_b200_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r1_bench_S3_n1.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and line["dtype"] == "f64" and line["n_gpus"] == 1 and line["warmup"] >= 3
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
