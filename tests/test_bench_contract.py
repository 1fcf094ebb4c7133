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


def test_committed_round_2_lines_carry_parity_and_secondary():
    """Round 2: every line has a `parity` block (error norms against the exact solution + bit checksum); multi-GPU lines say
    whether the checksum equals the one-box replica's; the N = 1 line carries the secondary measurements."""
    def load(name):
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            return json.loads(fh.read().strip().splitlines()[-1])

    one = load("r02_d_bench_512.json")
    assert BASE_KEYS <= set(one) and one["n_gpus"] == 1
    assert {"ns", "ns_wcns6ld", "shock", "fe"} <= set(one["secondary"])
    assert one["parity"]["matches_n1"] is True and one["parity"]["L1_error"] < 1e-12
    for name, n in (("r02_b_bench_512_2gpu.json", 2), ("r02_m_bench_512_4gpu.json", 4), ("r02_f_bench_512_8gpu.json", 8)):
        d = load(name)
        assert d["n_gpus"] == n and d["scaling"] == "strong"
        p = d["parity"]
        assert p["matches_n1"] is True and p["checksum"] == p["n1_checksum"] == one["parity"]["checksum"]
        assert p["L1_error"] == one["parity"]["L1_error"]
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["clocks"]["reasons"] == []
    eight = load("r02_f_bench_512_8gpu.json")
    assert eight["secondary"]["ns"]["checksum"] == one["secondary"]["ns"]["checksum"]      # Navier-Stokes: same bits on 8 GPUs
