"""The C++ host facade (addapt_b200/host): the reference's API -- Device / Aptamer / Context, RnaFold / ScoreFunction /
ScoreTerm / MacrostateProbTerm, MonteCarlo / moves / thermostats / reporters, the config mini-languages -- bound to the
C ABI.  Its test runner restates the reference's own Catch cases (tests/test_model.cc, test_sampling.cc,
test_scoring.cc); the CPU cases use fake fold engines, the GPU cases fold on the B200 engine through B200RnaFold."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "addapt_b200", "host")


def build():
    subprocess.check_call(["make", "-C", ROOT, "-s", "addapt_b200/libaddapt_b200.so"])
    subprocess.check_call(["make", "-C", HOST, "-s"])


def run(*flags):
    p = subprocess.run([os.path.join(HOST, "run_host_tests"), *flags], capture_output=True, text=True, timeout=600)
    print(p.stdout[-4000:], p.stderr[-2000:])
    return p


def test_host_facade_cpu_cases():
    build()
    p = run()
    assert p.returncode == 0, p.stdout[-2000:]
    assert " 0 failures" in p.stdout and "18 cases" in p.stdout


def test_host_library_exports_the_reference_api():
    build()
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", os.path.join(HOST, "libaddapt_host.so")], text=True)
    for name in ("addapt::Device::macrostate", "addapt::Device::mutate", "addapt::ScoreFunction::evaluate", "addapt::MacrostateProbTerm::evaluate",
                 "addapt::MonteCarlo::apply", "addapt::mutate_recursively", "addapt::can_be_freely_mutated", "addapt::UnbiasedMutationMove::apply",
                 "addapt::AnnealingThermostat::adjust", "addapt::AutoScalingThermostat::adjust", "addapt::TsvTrajectoryReporter::update",
                 "addapt::thermostat_from_str", "addapt::score_term_from_str", "addapt::device_from_yaml", "addapt::scorefxn_from_yaml",
                 "addapt::B200RnaFold::macrostate_prob", "addapt::B200RnaFold::base_pair_prob", "addapt::B200ScoreFunction::evaluate_batch",
                 "addapt::BatchedMonteCarlo::apply", "addapt::normalize_index", "addapt::PairProbTerm::evaluate",
                 "addapt::PairedDensityTerm::evaluate", "addapt::BatchedMonteCarlo::trajectory_files", "addapt::bp_term_from_str"):
        assert name in syms, name


@pytest.mark.gpu
def test_host_facade_gpu_cases():
    """reference tests/test_scoring.cc:83-259 against B200RnaFold, plus the batched score / Monte-Carlo bindings"""
    if not os.path.exists(os.path.join(HOST, "run_host_tests")):
        build()
    p = run("--gpu")
    assert p.returncode == 0, p.stdout[-3000:]
    assert " 0 failures" in p.stdout and "8 cases" in p.stdout


RHF6_YAML = """\
sequence: guuuuagagcuagaaauagcaaguuaaaauaaggcuaguccCuUUUCGCCgauaccagccgaaaggcccuuggcagcGACggcaccgagucggugcuuuuuu
macrostates:
  active: (............................)xx..xxxxx..xxxxxx(...............................).(.............)......
objective:
  apo: not active
  holo: active
aptamer:
  sequence: GAUACCAGCCGAAAGGCCCUUGGCAGC
  fold: (...((.(((....)))....))...)
  affinity: 0.32
thermostat: 5 to 0 in 300 steps
"""


def test_cli_usage_without_a_gpu():
    build()
    exe = os.path.join(HOST, "addapt")
    p = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "addapt <config>... [options]" in p.stdout and "--output-interval" in p.stdout
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 1
    p = subprocess.run([exe, "/nonexistent.yml"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stderr.startswith("Error: ")


@pytest.mark.gpu
def test_cli_design_run(tmp_path):
    """the reference's driver (apps/addapt.cc:18-93): `addapt <config> -n -T -r -o -i` on the GPU engine; the trajectory file is
    the one the reference's tools read (tools/utils.py:6-31)"""
    import pandas as pd
    if not os.path.exists(os.path.join(HOST, "addapt")):
        build()
    cfg = tmp_path / "rhf6.yml"
    cfg.write_text(RHF6_YAML)
    out = tmp_path / "traj.tsv"
    p = subprocess.run([os.path.join(HOST, "addapt"), str(cfg), "-n", "40", "-r", "7", "-o", str(out), "-i", "2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    first = open(out).readline().split("\t")
    assert first[0] == "#" and first[1] == "initial_seq" and first[2].strip().startswith("guuuuagagcuagaaauagc")
    traj = pd.read_table(out, comment="#")
    assert list(traj["step"]) == list(range(0, 40, 2)) and set(traj["num_steps"]) == {40}
    assert list(traj.columns[:8]) == ["step", "num_steps", "current_score", "proposed_score", "term_weight[apo: not active]",
                                      "term_value[apo: not active]", "term_weight[holo: active]", "term_value[holo: active]"]
    assert traj["temperature"].iloc[1] == pytest.approx(5.0 - 2 * 5.0 / 300)
    # same seed, same trajectory; four trajectories at once: trajectory 0 is the single run again
    out4 = tmp_path / "many.tsv"
    p = subprocess.run([os.path.join(HOST, "addapt"), str(cfg), "-n", "40", "-r", "7", "-o", str(out4), "-i", "2", "--trajectories", "4"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    t0 = pd.read_table(tmp_path / "many.0.tsv", comment="#")
    assert t0.equals(traj)
    assert (tmp_path / "many.3.tsv").exists()
