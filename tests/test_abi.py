"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ps_b200.h declares; struct mirrors agree; without a GPU the library refuses (no CPU path)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from partitionsolvers_b200 import capi

HEADER = os.path.join(ROOT, "include", "ps_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", src)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.ABI_SYMBOLS) == declared
    assert lib.ps_abi_version() == 3


def test_header_is_plain_c_and_struct_mirrors_match(tmp_path):
    # compile the header as C and print the struct sizes / offsets the ctypes mirrors must match
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ps_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ps_problem), sizeof(ps_result),'
                    'sizeof(ps_timings), offsetof(ps_problem, perm), offsetof(ps_timings, level_cells),'
                    'offsetof(ps_timings, kernel_launches));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o",
                    str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(capi.PsProblem)
    assert sizes[1] == C.sizeof(capi.PsResult)
    assert sizes[2] == C.sizeof(capi.PsTimings)
    assert sizes[3] == capi.PsProblem.perm.offset
    assert sizes[4] == capi.PsTimings.level_cells.offset
    assert sizes[5] == capi.PsTimings.kernel_launches.offset


def test_status_strings():
    lib = capi.load()
    assert lib.ps_status_string(0) == b"ok"
    assert b"distributional" in lib.ps_status_string(capi.PS_ERR_DIST)
    assert b"no CPU path" in lib.ps_status_string(capi.PS_ERR_NODEVICE)


def test_no_cpu_fallback_without_device():
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PsError) as e:
        capi.Context()
    assert e.value.status == capi.PS_ERR_NODEVICE


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under partitionsolvers_b200/ or include/ may reference it."""
    bad = []
    for base in ("partitionsolvers_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_lib" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".i")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle_py|libps_oracle|libps_ref|dp_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_reference_sources_compile_against_the_facade(tmp_path):
    """Source compatibility of the drop-in boundary: the reference's OWN callers of the hot path
    (python_dpsolver.cpp, python_ltsssolver.cpp, the two demos, solver_timer.cpp) must compile unchanged when the
    six solver headers are replaced by include/partitionsolvers/compat/*.  The include directory is assembled in a
    temporary directory from symlinks into /root/reference (nothing is copied into the repository)."""
    ref_inc = "/root/reference/src/cpp/include"
    if not os.path.isdir(ref_inc):
        pytest.skip("/root/reference not present on this machine")
    inc = tmp_path / "include"
    inc.mkdir()
    replaced = {"DP.hpp", "DP_impl.hpp", "score.hpp", "score_impl.hpp", "LTSS.hpp", "LTSS_impl.hpp"}
    for f in os.listdir(ref_inc):
        if f not in replaced:
            os.symlink(os.path.join(ref_inc, f), inc / f)
    compat = os.path.join(ROOT, "include", "partitionsolvers", "compat")
    for f in os.listdir(compat):
        os.symlink(os.path.join(compat, f), inc / f)
    for src in ("python_dpsolver.cpp", "python_ltsssolver.cpp", "DP_solver_demo.cpp", "LTSS_solver_demo.cpp",
                "solver_timer.cpp"):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", str(inc), "-I",
                            os.path.join(ROOT, "include"), os.path.join("/root/reference/src/cpp", src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, f"{src}:\n{r.stderr[:2000]}"
