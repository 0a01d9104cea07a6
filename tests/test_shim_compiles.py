"""The `.Call` shim the R package would ship (src/ppcseq_b200_shim.c) cannot be built here
-- no R headers in this image -- so it is type-checked against minimal stand-ins for R's public C API
(tests/r_stubs/) and the real include/ppcseq_b200.h: every ppcseq_* call in the shim must match the C ABI, and the
argument counts in its R_CallMethodDef table must match the entry points' parameter lists."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "src", "ppcseq_b200_shim.c")


def test_shim_type_checks_against_the_c_abi():
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
           "-Werror=int-conversion", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "r_stubs"),
           "-I" + os.path.join(ROOT, "include"), SHIM]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_call_table_matches_the_entry_points():
    src = open(SHIM).read()
    table = dict((m.group(1), int(m.group(2)))
                 for m in re.finditer(r'\{"(ppcseqb200_\w+)",\s*\(DL_FUNC\)\s*&\s*\1,\s*(\d+)\}', src))
    assert len(table) >= 8
    for name, nargs in table.items():
        m = re.search(r"\bSEXP\s+" + name + r"\s*\(([^)]*)\)", src)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == nargs, (name, nargs, len(params))
    # every .Call entry point defined in the shim is registered
    defined = set(re.findall(r"^SEXP\s+(ppcseqb200_\w+)\s*\(", src, flags=re.M))
    assert defined == set(table), defined ^ set(table)
