"""Writes a job directory for the reference's chicken solver (src/chicken, the reference-side CUDA code for this kind
of update: AUSMDV + van Albada + TVD-RK3 on structured blocks): the 3D ideal-air box of BASELINE.json configs[2]/[3]
at any size, 2 x 2 x 2 blocks.  Test/bench infrastructure like the rest of oracle/: the comparator bench.py times
beside the product (`also.chicken`), never part of it.

Formats (read from the reference's writers, nothing is copied): config.json = oracle/chicken/template_config.json
(written by chkn-prep for oracle/chicken/box.py, see make_template.sh) with the block sizes replaced;
grid/grid-iiii-jjjj-kkkk.bin = doubles [dims,0,0],[niv,njv,nkv], then (x,y,z) per vertex, i fastest
(gdtk/geom/sgrid.py:275-289); flow/t0000/flow-....bin = one array per iovar_names entry, doubles
(chkn_prep.py:749-777; the position arrays are flattened k fastest there, the flow arrays i fastest).

usage: python oracle/chicken/make_job.py <job_dir> <cells per side> [max_step]"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_job(job_dir, n, max_step=20, print_count=5, seed=1234):
    with open(os.path.join(HERE, "template_config.json")) as f:
        J = json.load(f)
    nb = J["nib"]
    assert (J["njb"], J["nkb"]) == (nb, nb) and n % nb == 0
    m = n // nb
    J["nics"] = J["njcs"] = J["nkcs"] = [m] * nb
    J["max_step"], J["print_count"] = int(max_step), int(print_count)
    J["max_time"] = 1.0
    J["dt_plot"] = [1.0e9]                      # no flow dumps inside the timed steps
    os.makedirs(os.path.join(job_dir, "grid"), exist_ok=True)
    os.makedirs(os.path.join(job_dir, "flow", "t0000"), exist_ok=True)
    with open(os.path.join(job_dir, "config.json"), "w") as f:
        json.dump(J, f, indent=1)
    with open(os.path.join(job_dir, "times.data"), "w") as f:
        f.write("# tindx t\n0 0.0\n")
    gas, fs = J["gas_model"], J["flow_states"][0]
    h = 1.0 / n
    rng = np.random.default_rng(seed)
    for blk in J["fluid_blocks"]:
        ib, jb, kb = blk["i"], blk["j"], blk["k"]
        tag = "%04d-%04d-%04d" % (ib, jb, kb)
        xv = (np.arange(m + 1) + ib * m) * h
        yv = (np.arange(m + 1) + jb * m) * h
        zv = (np.arange(m + 1) + kb * m) * h
        Z, Y, X = np.meshgrid(zv, yv, xv, indexing="ij")          # (k, j, i): i fastest when flattened
        data = np.zeros(((m + 1) ** 3 + 2, 3))
        data[0, :] = [3.0, 0.0, 0.0]
        data[1, :] = [m + 1, m + 1, m + 1]
        data[2:, 0], data[2:, 1], data[2:, 2] = X.ravel(), Y.ravel(), Z.ravel()
        data.tofile(os.path.join(job_dir, "grid", f"grid-{tag}.bin"))
        xc, yc, zc = 0.5 * (xv[1:] + xv[:-1]), 0.5 * (yv[1:] + yv[:-1]), 0.5 * (zv[1:] + zv[:-1])
        Zc, Yc, Xc = np.meshgrid(zc, yc, xc, indexing="ij")       # (k, j, i)
        # the same initial state as gdtk_b200.cases.box3d: inflow + smooth perturbation + seeded noise on rho and p
        fac = 1.0 + 1.0e-3 * np.sin(2 * np.pi * Xc) * np.sin(2 * np.pi * Yc) * np.sin(2 * np.pi * Zc) + 1.0e-6 * (rng.random(Xc.shape) - 0.5)
        g = fs["gas"]
        arrays = {"posx": np.transpose(Xc, (2, 1, 0)), "posy": np.transpose(Yc, (2, 1, 0)), "posz": np.transpose(Zc, (2, 1, 0)),
                  "vol": np.full((m, m, m), h ** 3), "p": g["p"] * fac, "T": np.full(Xc.shape, float(g["T"])),
                  "rho": g["rho"] * fac, "e": np.full(Xc.shape, float(g["e"])), "YB": np.zeros(Xc.shape),
                  "a": np.full(Xc.shape, float(g["a"])), "velx": np.full(Xc.shape, float(fs["vel"][0])),
                  "vely": np.full(Xc.shape, float(fs["vel"][1])), "velz": np.full(Xc.shape, float(fs["vel"][2]))}
        with open(os.path.join(job_dir, "flow", "t0000", f"flow-{tag}.bin"), "wb") as f:
            for name in J["iovar_names"]:
                f.write(np.ascontiguousarray(arrays[name], dtype=np.float64).tobytes())
    return n ** 3


if __name__ == "__main__":
    cells = make_job(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 20)
    print(f"{cells} cells in {sys.argv[1]}")
