"""Committed golden vectors (tests/golden/flow_*.npz, made by tests/golden/make_flow_fixtures.py):
the oracle must keep reproducing them bit for bit (CPU), and the FMA-free CUDA build must match
them bit for bit, the throughput build within 1e-10 (GPU)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_flow_fixtures as mk  # noqa: E402


def load(name):
    return dict(np.load(os.path.join(ROOT, "tests", "golden", f"flow_{name}.npz")))


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_oracle_reproduces_golden(oracle, name):
    got, ref = mk.run(oracle, name), load(name)
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_cuda_matches_golden(product, name):
    from gdtk_b200 import cases
    from util import run_case, max_rel_diff
    ref = load(name)
    fac, kw, nsteps = mk.CASES[name]
    for strict in (True, False):
        sim, U, _ = run_case(getattr(cases, fac), product, nsteps, strict=strict, **kw)
        assert sim.kernel_launches() > 0
        R = {bid: [ref[f"U_b{bid}_q{q}"] for q in range(len(arrs))] for bid, arrs in U.items()}
        if strict and not any(tag in name for tag in mk.NOT_BITWISE):       # (exp / pow / log of CUDA and glibc differ in the last place)
            for bid, arrs in U.items():
                for q, a in enumerate(arrs):
                    assert np.array_equal(a, R[bid][q]), (name, bid, q)
        else:
            # each variable against its scale over the whole field (momentum components share one), as in the parity tests
            assert max_rel_diff(U, R) < 1.0e-10, name
        dt = np.array(sim.dt_history)
        assert np.max(np.abs(dt - ref["dt_history"]) / ref["dt_history"]) < 1.0e-9
        sim.close()
