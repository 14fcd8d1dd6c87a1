#!/usr/bin/env python3
"""Writes tests/golden/oracle_trajectory.tsv: a 30-step annealing trajectory of a small design problem, produced by the ORACLE's
Monte-Carlo restatement (design_oracle.c) and laid out by addapt_b200/trajectory.py in TsvTrajectoryReporter's format
(reference src/sampling.cc:430-489).  The test compares a fresh oracle run, the emulation and (on the GPU) the engine with it, so
drift of the file format, of the random-stream convention or of the scores is caught.   python tests/golden/make_trajectory_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from addapt_b200 import trajectory  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.test_moves_and_trace import small_problem  # noqa: E402

pr = small_problem()
o = Oracle().mc_run(pr.as_oracle_dict(), 2016, 7, 30, trace_interval=1)
path = os.path.join(ROOT, "tests", "golden", "oracle_trajectory.tsv")
trajectory.write_tsv(path, pr, o["trace_values"], o["trace_seqs"], 30)
print("wrote", path)
