"""The R-side shim cannot be built here (no R).  Type-check it against a minimal stub of the R C
API so that at least every call into include/ldsr_b200.h matches the declared signatures, and
check that it registers the reference's four .Call names with the reference's arities
(src/RcppExports.cpp:133-136)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "ldsr_b200", "r", "ldsr_b200_shim.c")


def test_shim_type_checks_against_the_abi_header():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "r_stub"), SHIM],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_registers_reference_entry_points():
    src = open(SHIM).read()
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(_ldsr_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', src))
    for name, arity in (("_ldsr_Kalman_smoother", 5), ("_ldsr_Mstep", 4), ("_ldsr_LDS_EM", 6), ("_ldsr_propagate", 5)):
        assert table.get(name) == arity
    assert "R_init_ldsr" in src and "R_useDynamicSymbols(dll, FALSE)" in src
