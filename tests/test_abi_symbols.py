"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol
include/ldsr_b200.h declares, and refuses loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

from ldsr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ldsr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldsr_[a-z0-9_]+)\s*\(", src)) - {"ldsr_poll_fn"})


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(_lib.EXPORTS) == names
    assert L.ldsr_abi_version() == 1


def test_struct_layouts_match_header():
    # field order of the ctypes mirrors == declaration order in the header
    src = open(os.path.join(ROOT, "include", "ldsr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(struct_name):
        end = src.index("} %s;" % struct_name)
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                out.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        return out

    assert fields("ldsr_batch") == [f[0] for f in _lib.Batch._fields_]
    assert fields("ldsr_em_result") == [f[0] for f in _lib.EmResult._fields_]
    assert fields("ldsr_options") == [f[0] for f in _lib.Options._fields_]


def test_no_cpu_fallback_without_device():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    y = np.zeros(10)
    th = np.array([[0.5, 0.5, 1, 1, 0, 1.0]])
    with pytest.raises(_lib.LdsrError) as e:
        _lib.em_batch([dict(y=y, u=None, v=None, p=0, q=0)], [0], None, [0], th, 10, 1e-5)
    assert e.value.code == _lib.ERR_CUDA
    with pytest.raises(_lib.LdsrError):
        _lib.Plan([dict(y=y, u=None, v=None, p=0, q=0)], [0], None, [0], th)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ldsr_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".R")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ldsr_oracle" not in txt, f
