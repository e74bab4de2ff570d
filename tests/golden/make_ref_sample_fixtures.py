"""Writes tests/golden/ref_sample_cone20.npz from the reference's own sample output
(src/eilmer/sample-data/cone20.{grid,flow}.b000{0,1}.t0000.gz of gdtk-uq/gdtk): the vertex
coordinates of the two cone20 grids as the reference's preparation stage wrote them (block 1 is the
area-orthogonality grid) and every column of its initial flow files (cell centres, volumes, gas
state).  The arrays pin gdtk_b200/geometry.py and serve as the reference grid of a cone20 run.

    python tests/golden/make_ref_sample_fixtures.py [/root/reference]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    from gdtk_b200 import io
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    d = os.path.join(ref, "src", "eilmer", "sample-data")
    out = {}
    for b in (0, 1):
        g = io.read_grid(os.path.join(d, f"cone20.grid.b{b:04d}.t0000.gz"))
        out[f"b{b}_X"], out[f"b{b}_Y"] = g["X"][0], g["Y"][0]
        f = io.read_flow(os.path.join(d, f"cone20.flow.b{b:04d}.t0000.gz"))
        out[f"b{b}_sim_time"] = np.array(f["sim_time"])
        out[f"b{b}_names"] = np.array(f["names"])
        for name, a in f["data"].items():
            out[f"b{b}_{name}"] = a
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_sample_cone20.npz"), **out)
    print("wrote ref_sample_cone20.npz with", len(out), "arrays")
