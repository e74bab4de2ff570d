"""Small structured-grid generators for job scripts, tests and the benchmark.

Grid generation is prep-stage host work in Eilmer (src/geom, src/eilmer/fbarray.lua) and
outside the accelerated path; these helpers cover what the configurations of
BASELINE.json need: boxes, straight-sided quadrilateral patches (Coons/TFI patch with
straight edges = bilinear map) and splitting a grid into an array of blocks
(``FBArray:new{grid=, nib=, njb=, nkb=}``).
"""
import numpy as np

from . import _abi
from .geometry import BlockGeometry, NG


def box_grid_2d(x0, x1, y0, y1, nic, njc):
    x = x0 + (x1 - x0) * (np.arange(nic + 1) / nic)
    y = y0 + (y1 - y0) * (np.arange(njc + 1) / njc)
    X, Y = np.meshgrid(x, y, indexing="xy")          # shape (njv, niv)
    return X.copy(), Y.copy()


def box_grid_3d(p0, p1, nic, njc, nkc):
    x = p0[0] + (p1[0] - p0[0]) * (np.arange(nic + 1) / nic)
    y = p0[1] + (p1[1] - p0[1]) * (np.arange(njc + 1) / njc)
    z = p0[2] + (p1[2] - p0[2]) * (np.arange(nkc + 1) / nkc)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")    # shape (nkv, njv, niv)
    return X.copy(), Y.copy(), Z.copy()


def quad_patch_grid(p00, p10, p11, p01, nic, njc):
    """Straight-edged quadrilateral p00 (i=0,j=0), p10 (i=max,j=0), p11, p01; uniform
    parameter distribution (makePatch with Line edges, default Coons patch)."""
    r = (np.arange(nic + 1) / nic)[None, :]
    s = (np.arange(njc + 1) / njc)[:, None]
    out = []
    for m in range(2):
        out.append((1 - r) * (1 - s) * p00[m] + r * (1 - s) * p10[m] + r * s * p11[m] + (1 - r) * s * p01[m])
    return out[0], out[1]


def roberts_function(end0, end1, beta):
    """RobertsFunction of src/geom/misc/univariatefunctions.d:110-181 as a callable on [0, 1]."""
    cluster = (end0 or end1) and beta > 1.0
    alpha = 0.5 if (end0 and end1) else 0.0
    reverse = bool(end0 and not end1)

    def f(t):
        t = np.asarray(t, dtype=float)
        if reverse:
            t = 1.0 - t
        if cluster:
            lam = ((beta + 1.0) / (beta - 1.0)) ** ((t - alpha) / (1.0 - alpha))
            tbar = ((beta + 2.0 * alpha) * lam - beta + 2.0 * alpha) / ((2.0 * alpha + 1.0) * (1.0 + lam))
        else:
            tbar = t
        return 1.0 - tbar if reverse else tbar
    return f


def hex_volume_grid(corners, niv, njv, nkv, cf_r=None, cf_s=None, cf_t=None):
    """StructuredGrid:new{pvolume=TFIVolume:new{vertices=p0..p7}, niv=, njv=, nkv=, cfList=} for a
    hexahedron with straight edges (then the TFI volume is the trilinear map of its corners) and one
    cluster function per parameter direction (all four edges of a direction alike, so the blended
    parameter of sgrid.d:690-732 is the edge value).  corners: p0..p7 in Eilmer's order (p0-p3 bottom
    face counter-clockwise, p4-p7 above them).  Returns X, Y, Z of shape (nkv, njv, niv)."""
    ident = lambda t: np.asarray(t, dtype=float)
    r = (cf_r or ident)(np.arange(niv) / (niv - 1))[None, None, :]
    s = (cf_s or ident)(np.arange(njv) / (njv - 1))[None, :, None]
    t = (cf_t or ident)(np.arange(nkv) / (nkv - 1))[:, None, None]
    P = np.asarray(corners, dtype=float)
    w = [(1 - r) * (1 - s) * (1 - t), r * (1 - s) * (1 - t), r * s * (1 - t), (1 - r) * s * (1 - t),
         (1 - r) * (1 - s) * t, r * (1 - s) * t, r * s * t, (1 - r) * s * t]
    out = []
    for m in range(3):
        a = 0.0
        for n in range(8):
            a = a + w[n] * P[n][m]
        out.append(np.ascontiguousarray(a))
    return out[0], out[1], out[2]


def uniform_box_geometry(dims, nic, njc, nkc, dx, dy, dz=1.0):
    """BlockGeometry of a uniform Cartesian block without touching every vertex.

    The metrics of one cell are computed with the general formulas (geometry_2d / geometry_3d)
    on a 2-cell-wide reference block and replicated.  For spacings and origins that are
    dyadic fractions all coordinate differences are exact, so every cell of the real block
    gets bit-identical metrics from the general formulas (tests/test_geometry.py checks this);
    this is the fast set-up route for the large benchmark grids.
    """
    from .geometry import geometry_2d, geometry_3d
    g = BlockGeometry(dims, nic, njc, nkc if dims == 3 else 1)
    if dims == 3:
        ref = geometry_3d(*box_grid_3d((0.0, 0.0, 0.0), (2 * dx, 2 * dy, 2 * dz), 2, 2, 2))
        c = (NG, NG, NG)
        g.vol[...] = ref.vol[c]
    else:
        ref = geometry_2d(*box_grid_2d(0.0, 2 * dx, 0.0, 2 * dy, 2, 2))
        c = (0, NG, NG)
        g.vol[...] = ref.vol[c]
        g.areaxy[...] = ref.areaxy[c]
    for d in range(dims):
        g.len[d][...] = ref.len[d][c]
        for m in range(10):
            g.face[d][m][...] = ref.face[d][m][c]
    return g


def split_grid(grid, nib, njb, nkb=1):
    """Cut a vertex grid into nib x njb x nkb sub-grids (FBArray); returns a list of
    (ib, jb, kb, subgrid) in the reference's block order (i fastest? no: fbarray.lua
    numbers blocks with k fastest, then j, then i)."""
    P = [np.asarray(a) for a in grid]
    dims = len(P)
    if dims == 2:
        njv, niv = P[0].shape
        nkc = 1
    else:
        nkv, njv, niv = P[0].shape
        nkc = nkv - 1
    nic, njc = niv - 1, njv - 1

    def cuts(n, nb):
        base, extra = divmod(n, nb)
        edges = [0]
        for b in range(nb):
            edges.append(edges[-1] + base + (1 if b < extra else 0))
        return edges
    ci, cj, ck = cuts(nic, nib), cuts(njc, njb), cuts(nkc, nkb) if dims == 3 else [0, 1]
    out = []
    for ib in range(nib):
        for jb in range(njb):
            for kb in range(nkb if dims == 3 else 1):
                if dims == 2:
                    sub = tuple(a[cj[jb]:cj[jb + 1] + 1, ci[ib]:ci[ib + 1] + 1].copy() for a in P)
                else:
                    sub = tuple(a[ck[kb]:ck[kb + 1] + 1, cj[jb]:cj[jb + 1] + 1, ci[ib]:ci[ib + 1] + 1].copy() for a in P)
                out.append((ib, jb, kb, sub))
    return out


def connect_block_array(blocks_by_index, dims):
    """Give every interior face of a regular block array its ExchangeBC_FullFace.
    blocks_by_index: dict (ib, jb, kb) -> FluidBlock (with .id set)."""
    from .sim import ExchangeBC_FullFace
    for (ib, jb, kb), blk in blocks_by_index.items():
        for d, (di, dj, dk) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
            if d >= dims:
                continue
            nb = blocks_by_index.get((ib + di, jb + dj, kb + dk))
            if nb is not None:
                blk.bcList[_abi.FACE_NAMES[2 * d + 1]] = ExchangeBC_FullFace(nb.id, 2 * d, 0)
                nb.bcList[_abi.FACE_NAMES[2 * d]] = ExchangeBC_FullFace(blk.id, 2 * d + 1, 0)
