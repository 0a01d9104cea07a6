"""CPU: the C-ABI library loads and exports exactly the symbols include/ppcseq_b200.h declares."""
import os
import re

import ppcseq_b200
from ppcseq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ppcseq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(ppcseq_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    L = ppcseq_b200.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_the_header():
    assert _declared() == set(_lib.SIGNATURES), "ppcseq_b200/_lib.py out of sync with include/ppcseq_b200.h"


def test_abi_version_and_error_string():
    L = ppcseq_b200.lib()
    assert L.ppcseq_abi_version() >= 1
    assert isinstance(L.ppcseq_last_error(), bytes)
