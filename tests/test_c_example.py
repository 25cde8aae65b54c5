"""The drop-in boundary is a C ABI: a plain C program that includes only include/ldsr_b200.h and
links libldsr_b200.so must build with gcc (no torch, no Python in the link), refuse loudly without a
GPU, and run the reference's restart selection (R/LDS_reconstruction.R:42-62) with one."""
import os
import subprocess

import pytest

from ldsr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "em_restarts")
    libdir = os.path.dirname(_lib.SO_PATH)
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "em_restarts.c"), "-L", libdir, "-lldsr_b200",
                        "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_builds_and_refuses_without_a_device(tmp_path):
    exe = _build(tmp_path)
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libldsr_b200" in needed and "torch" not in needed and "python" not in needed
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stderr)
    assert "no CPU path" in r.stderr or "CUDA" in r.stderr


@pytest.mark.gpu
def test_c_program_runs_the_restart_selection(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 9 and lines[-1].startswith("selected restart")
