"""The C-ABI library: builds in-tree, loads, exports every symbol include/addapt_b200.h declares, and refuses to
run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "addapt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import addapt_b200
    lib = addapt_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libaddapt_b200.so does not export " + s


def test_library_is_built_for_sm100a_only():
    import subprocess
    import addapt_b200
    out = subprocess.run(["cuobjdump", "--list-elf", addapt_b200.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import addapt_b200
    lib = addapt_b200.load_library()
    if lib.adp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(addapt_b200.EngineError) as ei:
        addapt_b200.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    # the oracle is test infrastructure: nothing under addapt_b200/ or include/ may reference it
    for base in ("addapt_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dp, f), errors="ignore").read()
                for line in text.splitlines():
                    if re.search(r"#include.*oracle|import\s+oracle|from\s+oracle|liboracle|orc_[a-z]+\(", line):
                        raise AssertionError("%s references the oracle: %s" % (os.path.join(dp, f), line.strip()))


def test_header_is_plain_c_and_cxx11():
    """the boundary is a C ABI: include/addapt_b200.h must compile as C99 (cgo / ctypes / JNI side) and as C++11 (the reference's
    language level, configure.ac), with no torch or CUDA types in it"""
    import subprocess
    hdr = os.path.join(ROOT, "include", "addapt_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    includes = [ln for ln in open(hdr) if ln.lstrip().startswith("#include")]
    assert all("<std" in ln for ln in includes), includes          # <stdint.h> / <stddef.h> only: no torch, no CUDA headers
