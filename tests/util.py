"""Helpers shared by the parity tests: run the same job through two implementations of the
C ABI (product CUDA library / CPU oracle) and compare conserved quantities, primitives and
time-step histories."""
import numpy as np

from gdtk_b200 import Simulation


def run_case(factory, lib, nsteps, strict=None, fixed_dt=None, **kw):
    """Build the case with `factory(**kw)`, run nsteps, return (sim, {blk_id: U list}, {blk_id: prim list})."""
    cfg, gm, blocks = factory(**kw)
    if strict is not None:
        cfg.strict_fp = strict
    sim = Simulation(cfg, gm, blocks, lib=lib)
    if fixed_dt is not None:
        cfg.fixed_time_step = True
        sim.dt_global = fixed_dt
    sim.run(max_step=nsteps, max_time=1.0e30)
    U = {b.id: [sim.interior(b.id, a).copy() for a in sim.download_conserved(b.id)] for b in sim.local_blocks}
    P = {b.id: [sim.interior(b.id, a).copy() for a in sim.download_flow(b.id)] for b in sim.local_blocks}
    return sim, U, P


def max_rel_diff(A, B):
    """Largest |a-b| / scale over corresponding arrays.  scale = max|b| of the variable over ALL
    blocks; the components of a vector (momentum in U lists, velocity in primitive lists) share
    one scale, so that a cross-flow component that is pure round-off, or a block the wave has not
    reached yet, is measured against the flow and not against its own noise."""
    bids = list(B)
    n = len(B[bids[0]])
    if n in (4, 5, 9, 10):                      # conserved: mass, momentum x dims, energy, [species]
        dims = 3 if n in (5, 10) else 2
        vec = list(range(1, 1 + dims))
    else:                                       # primitives: rho u p T a velx vely velz ...
        vec = [5, 6, 7]
    gmax = [max(float(np.max(np.abs(B[bid][q]))) for bid in bids) for q in range(n)]
    vscale = max(gmax[q] for q in vec)
    worst = 0.0
    for bid in bids:
        for q, (a, b) in enumerate(zip(A[bid], B[bid])):
            scale = vscale if q in vec else gmax[q]
            if scale == 0.0:
                scale = 1.0
            worst = max(worst, float(np.max(np.abs(a - b))) / scale)
    return worst


def identical(A, B):
    return all(np.array_equal(a, b) for bid in B for a, b in zip(A[bid], B[bid]))


def cellwise_rel_diff(A, B):
    """Largest |a-b| / |b| cell by cell over the positive-definite conserved quantities (mass, total energy and
    species densities above 1e-12 of the mixture); the momentum components have no cell-wise scale of their own
    (they pass through zero) and are measured by max_rel_diff."""
    worst = 0.0
    for bid in B:
        n = len(B[bid])
        dims = 3 if n in (5, 10) else 2
        rho = np.abs(B[bid][0])
        for q in [0, 1 + dims] + list(range(2 + dims, n)):
            a, b = A[bid][q], B[bid][q]
            keep = np.abs(b) > 1.0e-12 * rho if q > 1 + dims else np.ones_like(b, dtype=bool)
            if keep.any():
                worst = max(worst, float(np.max(np.abs(a - b)[keep] / np.abs(b)[keep])))
    return worst
