"""The exact pass divides every row element by the row norm BEFORE the dot (sklearn's normalize() order, lib.py:51).  On the
GPU that quotient is computed from ONE reciprocal per row plus two FMA corrections (csrc/exact.cuh: div_by_norm) instead of a
full fp64 division per element.  This test restates the sequence in C (gcc, hardware FMA) and holds it to IEEE division."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reciprocal_plus_two_fma_corrections_equals_ieee_division(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "exact_division")
    r = subprocess.run([gcc, "-O2", "-mfma", "-o", exe, os.path.join(HERE, "c", "exact_division.c"), "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no hardware FMA on this host: " + r.stderr[-200:])
    r = subprocess.run([exe, "10000000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "5-op mismatches 0" in r.stdout, r.stdout
