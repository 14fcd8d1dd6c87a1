"""bench.py's reference arm runs on CPU (it times the oracle port), so its JSON contract can be checked without a GPU;
the same keys are produced by the GPU arm (checked on the box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, ADDAPT_BENCH_REF_SECONDS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "fold-evals/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("sgRNA fold-evals/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["config"]["workload"].startswith("configs[2]")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_term_count_matches_the_survey_figures():
    sys.path.insert(0, ROOT)
    import numpy as np
    from bench import count_terms, random_seqs
    w = np.array([count_terms(r) for r in random_seqs(6, 7, 100)])
    # random 100-mers have W_in ~ 0.76 M and W_out ~ 0.39 M terms
    assert 0.70e6 < w[:, 0].mean() < 0.82e6 and 0.35e6 < w[:, 1].mean() < 0.44e6
