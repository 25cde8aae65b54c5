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


def test_argument_validation_runs_on_the_host():
    """Argument errors are reported (code LDSR_ERR_ARG / UNSUPPORTED) before any device is touched,
    so these hold with or without a GPU; with valid arguments and no device the new entry points
    refuse with LDSR_ERR_CUDA like the rest (no CPU fallback)."""
    y = np.zeros(12)
    with pytest.raises(_lib.LdsrError) as e:  # state dimension beyond LDSR_MAX_STATE_DIM
        _lib.smoother_d(5, y, None, None, np.zeros(2 * 25 + 5 + 1 + 5 + 25))
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.LdsrError) as e:  # theta shorter than the layout needs
        _lib.smoother_d(2, y, None, None, np.zeros(5))
    assert e.value.code == _lib.ERR_ARG
    yy = y.copy()
    yy[3] = np.inf
    with pytest.raises(_lib.LdsrError) as e:
        _lib.smoother_d(1, yy, None, None, np.array([0.5, 1.0, 1.0, 1.0, 0.0, 1.0]))
    assert e.value.code == _lib.ERR_ARG
    sim = np.ones((2, 12))
    with pytest.raises(_lib.LdsrError) as e:  # R indices are 1-based: 0 is out of range
        _lib.cv_metrics(sim, y + 1.0, [np.array([0, 1]), np.array([2, 3])])
    assert e.value.code == _lib.ERR_ARG
    with pytest.raises(_lib.LdsrError) as e:
        _lib.cv_metrics(sim, y + 1.0, [np.array([1, 13]), np.array([2, 3])])
    assert e.value.code == _lib.ERR_ARG
    if _lib.device_count() == 0:
        with pytest.raises(_lib.LdsrError) as e:
            _lib.smoother_d(1, y, None, None, np.array([0.5, 1.0, 1.0, 1.0, 0.0, 1.0]))
        assert e.value.code == _lib.ERR_CUDA
        with pytest.raises(_lib.LdsrError) as e:
            _lib.cv_metrics(sim, y + 1.0, [np.array([1, 2]), np.array([2, 3])])
        assert e.value.code == _lib.ERR_CUDA
