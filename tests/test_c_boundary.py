"""The drop-in boundary driven from plain C (tests/c_boundary/sod_abi.c): init -> block_create -> block_set_geometry ->
block_set_bc -> commit -> upload_flow -> compute_dt -> step x N -> download_conserved -> finalize, the call sequence of
INTEGRATION.md section 2, with the product and the oracle loaded by dlopen and no Python between them.  The closest
stand-in for the D shim (extern(C) over the same symbols) that this image's toolchains allow."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_boundary", "sod_abi.c")
ORACLE = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
PRODUCT = os.path.join(ROOT, "gdtk_b200", "csrc", "libeb200.so")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "sod_abi")
    subprocess.run(["gcc", "-O1", "-std=c99", "-D_GNU_SOURCE", "-o", exe, SRC, "-ldl", "-lm"], check=True, capture_output=True, text=True)
    return exe


def test_c_driver_runs_the_oracle(tmp_path):
    """CPU only: the C program builds against include/eb200.h and drives the oracle through the whole call sequence."""
    exe = _build(tmp_path)
    r = subprocess.run([exe, "-", ORACLE, "20"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "oracle only: ok" in r.stdout


@pytest.mark.gpu
def test_c_driver_product_matches_oracle(tmp_path):
    """GPU: the same C driver runs libeb200.so; FMA-free build bit-identical to the oracle, throughput build < 1e-10."""
    exe = _build(tmp_path)
    r = subprocess.run([exe, PRODUCT, ORACLE, "40"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-identical: yes" in r.stdout and "C boundary: ok" in r.stdout
