"""bench.py's output contract, checked where no GPU exists: the reference arm (`--impl reference`, the CPU oracle on a
bounded sample) must print ONE JSON line with the keys the driver reads, and under a multi-rank launch only rank 0
works.  The GPU arm's line is produced on the GPU box (profiles/r01_q_bench_512.json is a committed sample, checked
here for the same keys plus `roofline`, `cpu_baseline`, `e2e`, `clocks` and `gpu_launches`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-size", "16"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and "workload" in d["config"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_committed_gpu_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_q_bench_512.json")) as fh:
        d = json.loads(fh.read())
    assert BASE_KEYS <= set(d) and d["n_gpus"] == 1 and d["config"]["size"] == 512
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["traffic"] > 0
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == []
