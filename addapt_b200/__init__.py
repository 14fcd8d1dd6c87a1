"""addapt_b200 -- B200-native fold / score / Monte-Carlo engine for kalekundert/addapt's hot path.

The product is the C-ABI library ``libaddapt_b200.so`` (hand-written sm_100a CUDA, see
``include/addapt_b200.h``) and the C++ host facade in ``addapt_b200/host`` that mirrors the reference's
``RnaFold`` / ``ScoreFunction`` / ``MonteCarlo`` API.  This Python package is only the thin ctypes binding
the tests and ``bench.py`` drive the library through.  There is no CPU fallback: importing works anywhere,
but creating an :class:`Engine` without the built library or without a CUDA device raises.
"""
from .capi import Engine, EngineError, Problem, THEOPHYLLINE, library_path, load_library  # noqa: F401
