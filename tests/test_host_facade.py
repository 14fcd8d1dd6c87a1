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
    assert " 0 failures" in p.stdout and "16 cases" in p.stdout


def test_host_library_exports_the_reference_api():
    build()
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", os.path.join(HOST, "libaddapt_host.so")], text=True)
    for name in ("addapt::Device::macrostate", "addapt::Device::mutate", "addapt::ScoreFunction::evaluate", "addapt::MacrostateProbTerm::evaluate",
                 "addapt::MonteCarlo::apply", "addapt::mutate_recursively", "addapt::can_be_freely_mutated", "addapt::UnbiasedMutationMove::apply",
                 "addapt::AnnealingThermostat::adjust", "addapt::AutoScalingThermostat::adjust", "addapt::TsvTrajectoryReporter::update",
                 "addapt::thermostat_from_str", "addapt::score_term_from_str", "addapt::device_from_yaml", "addapt::scorefxn_from_yaml",
                 "addapt::B200RnaFold::macrostate_prob", "addapt::B200RnaFold::base_pair_prob", "addapt::B200ScoreFunction::evaluate_batch",
                 "addapt::BatchedMonteCarlo::apply", "addapt::normalize_index"):
        assert name in syms, name


@pytest.mark.gpu
def test_host_facade_gpu_cases():
    """reference tests/test_scoring.cc:83-259 against B200RnaFold, plus the batched score / Monte-Carlo bindings"""
    if not os.path.exists(os.path.join(HOST, "run_host_tests")):
        build()
    p = run("--gpu")
    assert p.returncode == 0, p.stdout[-3000:]
    assert " 0 failures" in p.stdout and "5 cases" in p.stdout
