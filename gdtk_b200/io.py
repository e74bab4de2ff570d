"""Eilmer's native on-disk formats for structured blocks: what sits either side of
the accelerated path on disk.

* grid files  ``grid/tNNNN/<job>.grid.bBBBB.tNNNN.gz``: "structured_grid 1.1" formatted text,
  ``StructuredGrid.read_from_gzip_file`` / ``write_to_gzip_file`` (src/geom/grid/sgrid.d:854-922,999-1040)
* flow files  ``flow/tNNNN/<job>.flow.bBBBB.tNNNN.gz``: "structured_grid_flow 1.0" formatted text,
  ``read_legacy_solution`` / ``write_legacy_solution`` (src/eilmer/fluidblockio_old.d:167-275,380-445),
  one line per cell in the layout of ``cell_data_as_string`` (:1487-1575), variable list of
  ``build_flow_variable_list`` (:107-152)
* the files of src/eilmer/sample-data (an earlier layout of the same two formats without keywords in
  the header) are read as well; they are the reference's own sample output and serve as golden data
  (tests/golden/ref_sample_cone20.npz made by tests/golden/make_ref_sample_fixtures.py; tests/test_io_formats.py).

Host-side plumbing only (numpy + gzip); nothing here touches the GPU.
"""
import gzip
import io as _io
import os

import numpy as np

from .geometry import NG


def _open_text(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return _io.TextIOWrapper(gzip.open(path, "rb"), encoding="ascii")
    return open(path, "r", encoding="ascii")


def _keyed(line, key):
    head, _, val = line.partition(":")
    if head.strip() != key:
        raise ValueError(f"expected '{key}:' but found {line!r}")
    return val.strip()


# ------------------------------------------------------------------------------------------- grids

def read_grid(path, scale=1.0):
    """Returns {'dimensions', 'label', 'tags', 'X', 'Y', 'Z'}; the coordinate arrays have shape
    (nkv, njv, niv), vertex order k, j, i with i fastest like the file."""
    with _open_text(path) as f:
        first = f.readline()
        label, tags, dims = "", [], None
        if first.startswith("structured_grid"):
            version = first.split()[1]
            if version not in ("1.0", "1.1"):
                raise ValueError(f"read_grid: invalid format version found: {version}")
            label = _keyed(f.readline(), "label")
            dims = int(_keyed(f.readline(), "dimensions"))
            niv = int(_keyed(f.readline(), "niv"))
            njv = int(_keyed(f.readline(), "njv"))
            nkv = int(_keyed(f.readline(), "nkv"))
        else:                                   # sample-data layout: "ni nj nk  # comment"
            version = "0"
            niv, njv, nkv = (int(t) for t in first.split("#")[0].split())
        n = niv * njv * nkv
        xyz = np.loadtxt(f, dtype=np.float64, max_rows=n).reshape(n, 3) * scale
        if version == "1.1":
            ntags = int(_keyed(f.readline(), "ntags"))
            for i in range(ntags):
                tags.append(_keyed(f.readline(), f"tag[{i}]"))
    if dims is None:
        dims = 3 if nkv > 1 else (2 if njv > 1 else 1)
    shp = (nkv, njv, niv)
    return {"dimensions": dims, "label": label, "tags": tags,
            "X": xyz[:, 0].reshape(shp).copy(), "Y": xyz[:, 1].reshape(shp).copy(), "Z": xyz[:, 2].reshape(shp).copy()}


def write_grid(path, X, Y, Z=None, label="", tags=None, dimensions=None):
    """Format version 1.1, "%.18e" like the reference.  2D grids may be given as (njv, niv) arrays."""
    X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
    if X.ndim == 2:
        X, Y = X[None], Y[None]
    Z = np.zeros_like(X) if Z is None else np.asarray(Z, dtype=np.float64).reshape(X.shape)
    nkv, njv, niv = X.shape
    dims = dimensions or (3 if nkv > 1 else 2)
    tags = list(tags) if tags is not None else [""] * (6 if dims == 3 else 4)
    with gzip.open(path, "wt", encoding="ascii") as f:
        f.write(f"structured_grid 1.1\nlabel: {label}\ndimensions: {dims}\nniv: {niv}\nnjv: {njv}\nnkv: {nkv}\n")
        rows = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        np.savetxt(f, rows, fmt="%.18e")
        f.write(f"ntags: {len(tags)}\n")
        for i, t in enumerate(tags):
            f.write(f"tag[{i}]: {t}\n")


def grid_arrays(g):
    """The tuple geometry_2d / geometry_3d / FluidBlock expect from a read_grid() result."""
    if g["dimensions"] == 3:
        return g["X"], g["Y"], g["Z"]
    return g["X"][0], g["Y"][0]


# ------------------------------------------------------------------------------------------- flow

def flow_variable_list(gmodel):
    """build_flow_variable_list() for the configurations of this path (no MHD, no turbulence model,
    no thermal modes, no radiation)."""
    names = ["pos.x", "pos.y", "pos.z", "volume", "rho", "vel.x", "vel.y", "vel.z", "p", "a", "mu", "k", "mu_t", "k_t", "S"]
    species = list(getattr(gmodel, "species_names", None) or ["air"])
    for i, s in enumerate(species):
        names.append(f"massf[{i}]-" + "-".join(str(s).split()))
    if len(species) > 1:
        names.append("dt_chem")
    names += ["u", "T"]
    return names


def read_flow(path):
    """Returns {'sim_time', 'label', 'dimensions', 'names', 'data'}; data[name] has shape (nkc, njc, nic)."""
    with _open_text(path) as f:
        first = f.readline()
        label, dims = "", None
        if first.startswith("structured_grid_flow"):
            version = first.split()[1]
            if version != "1.0":
                raise ValueError(f"read_flow: file format version found: {version}")
            label = _keyed(f.readline(), "label")
            sim_time = float(_keyed(f.readline(), "sim_time"))
            nvar = int(_keyed(f.readline(), "variables"))
            names = [t.strip('"') for t in f.readline().split()]
            if len(names) != nvar:
                raise ValueError("read_flow: variable count does not match the list of names")
            dims = int(_keyed(f.readline(), "dimensions"))
            nic = int(_keyed(f.readline(), "nicell"))
            njc = int(_keyed(f.readline(), "njcell"))
            nkc = int(_keyed(f.readline(), "nkcell"))
        else:                                   # sample-data layout: time, names, "nic njc nkc"
            sim_time = float(first)
            names = [t.strip('"') for t in f.readline().split()]
            nic, njc, nkc = (int(t) for t in f.readline().split())
        n = nic * njc * nkc
        rows = np.loadtxt(f, dtype=np.float64, max_rows=n).reshape(n, len(names))
    if dims is None:
        dims = 3 if nkc > 1 else 2
    data = {name: rows[:, q].reshape(nkc, njc, nic).copy() for q, name in enumerate(names)}
    return {"sim_time": sim_time, "label": label, "dimensions": dims, "names": names, "data": data}


def write_flow(path, sim, blk_id, sim_time, label=""):
    """Write block blk_id of a Simulation the way write_legacy_solution does.  mu, k, mu_t, k_t are
    written as zero (the inviscid path carries no transport coefficients)."""
    blk = next(b for b in sim.local_blocks if b.id == blk_id)
    g = blk.geom
    P = [sim.interior(blk_id, a) for a in sim.download_flow(blk_id)]
    sl = (slice(g.kg, g.kg + g.nkc), slice(NG, NG + g.njc), slice(NG, NG + g.nic))
    nsp = len(getattr(sim.gmodel, "species_names", None) or ["air"])
    zero = np.zeros_like(P[0])
    cols = [g.pos[0][sl], g.pos[1][sl], g.pos[2][sl], g.vol[sl], P[0], P[5], P[6], P[7], P[2], P[4], zero, zero, zero, zero, zero]
    if nsp > 1:
        cols += [P[8 + i] for i in range(nsp)] + [np.full_like(zero, -1.0)]
    else:
        cols.append(np.ones_like(zero))
    cols += [P[1], P[3]]
    names = flow_variable_list(sim.gmodel)
    assert len(names) == len(cols)
    with gzip.open(path, "wt", encoding="ascii") as f:
        f.write("structured_grid_flow 1.0\n")
        f.write(f"label: {label}\nsim_time: {sim_time:.18e}\nvariables: {len(names)}\n")
        f.write("".join(f' "{n}"' for n in names) + "\n")
        f.write(f"dimensions: {sim.config.dimensions}\nnicell: {g.nic}\nnjcell: {g.njc}\nnkcell: {g.nkc}\n")
        rows = np.stack([np.asarray(c, dtype=np.float64).ravel() for c in cols], axis=1)
        for r in rows:
            f.write(" " + " ".join(f"{v:.18e}" for v in r) + "\n")


class FlowFromFile:
    """Initial state of a FluidBlock taken from a flow file (what read_solution does at start-up):
    rho, u, p, T, a and the velocity are copied; the library re-encodes and decodes them like
    simcore.d:325-334."""

    def __init__(self, flow, nsp=1):
        self.flow, self.nsp = flow, nsp

    def padded_arrays(self, geom):
        d = self.flow["data"]
        shp = (geom.NK, geom.NJ, geom.NI)
        sl = (slice(geom.kg, geom.kg + geom.nkc), slice(NG, NG + geom.njc), slice(NG, NG + geom.nic))
        energy = d["u"] if "u" in d else d["e[0]"]
        temp = d["T"] if "T" in d else d["T[0]"]
        src = [d["rho"], energy, d["p"], temp, d["a"], d["vel.x"], d["vel.y"], d["vel.z"]]
        if self.nsp > 1:
            mf = [d[n] for n in self.flow["names"] if n.startswith("massf[")]
            src += mf + [m * d["rho"] for m in mf]
        out = []
        for a in src:
            full = np.zeros(shp)
            # ghost cells get a valid state too (the first ghost-cell fill overwrites them)
            full[...] = a.ravel()[0]
            full[sl] = a
            out.append(full)
        return out


def write_history_file(path, sim, key):
    """hist/<job>-blk-B-cell-C.dat.0 of history.d:28-95: a header naming the columns, then one line per sample:
    time and the cell's flow-file line.  `key` is a history point of sim (Simulation.set_history_point)."""
    ib, i, j, k = key
    blk = next(b for b in sim.local_blocks if b.id == ib)
    g = blk.geom
    c = (g.kg + k, NG + j, NG + i)
    names = flow_variable_list(sim.gmodel)
    nsp = len(getattr(sim.gmodel, "species_names", None) or ["air"])
    with open(path, "w", encoding="ascii") as f:
        f.write("# 1:t " + "".join(f"{n + 2}:{v} " for n, v in enumerate(names)) + "\n")
        for row in sim.history[key]:
            t, P = row[0], row[1:]
            vals = [g.pos[0][c], g.pos[1][c], g.pos[2][c], g.vol[c], P[0], P[5], P[6], P[7], P[2], P[4], 0.0, 0.0, 0.0, 0.0, 0.0]
            vals += (list(P[8:8 + nsp]) + [-1.0]) if nsp > 1 else [1.0]
            vals += [P[1], P[3]]
            f.write(f"{t:.18e} " + " ".join(f"{v:.18e}" for v in vals) + "\n")


def job_file(job_dir, job, kind, blk_id, tindx):
    """<job_dir>/<kind>/tNNNN/<job>.<kind>.bBBBB.tNNNN.gz (kind = 'grid' or 'flow'), simcore_io.d naming."""
    return os.path.join(job_dir, kind, f"t{tindx:04d}", f"{job}.{kind}.b{blk_id:04d}.t{tindx:04d}.gz")


def write_solution_files(job_dir, job, sim, tindx):
    """write_solution_files (simcore_io.d:78-132) for the blocks of this process: flow/tNNNN/<job>.flow.bBBBB.tNNNN.gz
    for every local block and one line "tindx time dt" appended to config/<job>.times.  With the grid files and the
    .config the reference's preparation wrote, ``e4shared --post`` reads the result."""
    d = os.path.join(job_dir, "flow", f"t{tindx:04d}")
    os.makedirs(d, exist_ok=True)
    for b in sim.local_blocks:
        write_flow(job_file(job_dir, job, "flow", b.id, tindx), sim, b.id, sim.time, label=getattr(b, "label", "") or "")
    if getattr(sim, "rank", 0) == 0:
        os.makedirs(os.path.join(job_dir, "config"), exist_ok=True)
        with open(os.path.join(job_dir, "config", f"{job}.times"), "a", encoding="ascii") as f:
            f.write(f"{tindx:04d} {sim.time:.18e} {sim.dt_global:.18e}\n")


def read_times(job_dir, job):
    """config/<job>.times -> {tindx: (time, dt)}; lines starting with # are comments."""
    out = {}
    with open(os.path.join(job_dir, "config", f"{job}.times"), encoding="ascii") as f:
        for line in f:
            if line.strip() and not line.startswith("#"):
                t = line.split()
                out[int(t[0])] = (float(t[1]), float(t[2]))
    return out
