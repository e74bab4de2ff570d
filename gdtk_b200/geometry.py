"""Set-up-time geometry of a structured block, in the padded block layout.

This is the host-side work Eilmer does once in
``SFluidBlock.compute_primary_cell_geometric_data`` (reference
src/eilmer/sfluidblock.d:871-1088) before time stepping; the formulas follow

* 2D cells   ``xyplane_quad_cell_properties``  src/geom/elements/properties.d:176-238,
             ``FVCell.update_2D_geometric_data`` src/eilmer/fvcell.d:396-438
* 2D faces   ``FVInterface.update_2D_geometric_data`` src/eilmer/fvinterface.d:301-332
* 3D cells   ``hex_cell_properties`` src/geom/elements/properties.d:589-672
* 3D faces   ``quad_properties`` src/geom/elements/properties.d:96-139 with the vertex
             cycles of src/eilmer/sfluidblock.d:735-850
* ghost-cell lengths mirrored from the interior (ghost n <- interior n),
  src/eilmer/sfluidblock.d:897-1086.

Arrays are numpy float64 with shape (NK, NJ, NI) = padded layout of include/eb200.h.
"""
import numpy as np

NG = 2


class BlockGeometry:
    """vol, areaxy, len[3], face[d] (10, NK, NJ, NI), cell centroids pos (3, NK, NJ, NI)."""

    def __init__(self, dims, nic, njc, nkc):
        self.dims = dims
        self.nic, self.njc, self.nkc = nic, njc, nkc
        self.NI, self.NJ = nic + 2 * NG, njc + 2 * NG
        self.NK = nkc + 2 * NG if dims == 3 else 1
        self.kg = NG if dims == 3 else 0
        shp = (self.NK, self.NJ, self.NI)
        self.vol = np.zeros(shp)
        self.areaxy = np.zeros(shp)
        self.len = [np.zeros(shp) for _ in range(3)]
        self.face = [np.zeros((10,) + shp) for _ in range(dims)]
        self.pos = np.zeros((3,) + shp)

    def interior(self, a):
        return a[..., self.kg:self.kg + self.nkc, NG:NG + self.njc, NG:NG + self.nic]

    def mirror_ghost_lengths(self, face_id):
        """Ghost cell n takes the lengths of interior cell n (sfluidblock.d:902-933)."""
        d, hi = face_id // 2, face_id & 1
        n = (self.nic, self.njc, self.nkc)[d]
        off = NG if d < 2 else self.kg
        ax = 2 - d  # numpy axis of index direction d
        for arr in self.len:
            for layer in range(NG):
                src = off + (n - 1 - layer if hi else layer)
                dst = off + (n + layer if hi else -1 - layer)
                sl_src = [slice(None)] * 3
                sl_dst = [slice(None)] * 3
                sl_src[ax] = src
                sl_dst[ax] = dst
                arr[tuple(sl_dst)] = arr[tuple(sl_src)]


def _cross(ax, ay, az, bx, by, bz):
    return ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx


def geometry_2d(x, y, axisymmetric=False):
    """x, y: vertex coordinates, shape (njv, niv)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    njv, niv = x.shape
    nic, njc = niv - 1, njv - 1
    g = BlockGeometry(2, nic, njc, 1)
    # --- cells: p0=(i,j) p1=(i+1,j) p2=(i+1,j+1) p3=(i,j+1); A=p1 B=p2 C=p3 D=p0
    xD, yD = x[:-1, :-1], y[:-1, :-1]
    xA, yA = x[:-1, 1:], y[:-1, 1:]
    xB, yB = x[1:, 1:], y[1:, 1:]
    xC, yC = x[1:, :-1], y[1:, :-1]
    area = 0.5 * ((xB + xA) * (yB - yA) + (xC + xB) * (yC - yB) +
                  (xD + xC) * (yD - yC) + (xA + xD) * (yA - yD))
    cx = 1.0 / (area * 6.0) * ((yB - yA) * (xA * xA + xA * xB + xB * xB) +
                               (yC - yB) * (xB * xB + xB * xC + xC * xC) +
                               (yD - yC) * (xC * xC + xC * xD + xD * xD) +
                               (yA - yD) * (xD * xD + xD * xA + xA * xA))
    cy = -1.0 / (area * 6.0) * ((xB - xA) * (yA * yA + yA * yB + yB * yB) +
                                (xC - xB) * (yB * yB + yB * yC + yC * yC) +
                                (xD - xC) * (yC * yC + yC * yD + yD * yD) +
                                (xA - xD) * (yD * yD + yD * yA + yA * yA))
    xN, yN = 0.5 * (xC + xB), 0.5 * (yC + yB)
    xS, yS = 0.5 * (xD + xA), 0.5 * (yD + yA)
    xE, yE = 0.5 * (xA + xB), 0.5 * (yA + yB)
    xW, yW = 0.5 * (xD + xC), 0.5 * (yD + yC)
    dx, dy = xN - xS, yN - yS
    jLen = np.sqrt(dx * dx + dy * dy)
    dx, dy = xE - xW, yE - yW
    iLen = np.sqrt(dx * dx + dy * dy)
    vol = area * cy if axisymmetric else area
    if np.any(vol < 0.0):
        raise ValueError("Negative cell volume")
    inner = (0, slice(NG, NG + njc), slice(NG, NG + nic))
    g.vol[inner] = vol
    g.areaxy[inner] = area
    g.len[0][inner] = iLen
    g.len[1][inner] = jLen
    g.pos[0][inner] = cx
    g.pos[1][inner] = cy

    def face2d(xA, yA, xB, yB):
        LAB = np.sqrt((xB - xA) * (xB - xA) + (yB - yA) * (yB - yA))
        nx = (yB - yA) / LAB
        ny = -(xB - xA) / LAB
        nz = np.zeros_like(nx)
        t2x, t2y, t2z = np.zeros_like(nx), np.zeros_like(nx), np.ones_like(nx)
        t1x, t1y, t1z = _cross(nx, ny, nz, t2x, t2y, t2z)
        Ybar = 0.5 * (yA + yB)
        area = LAB * Ybar if axisymmetric else LAB
        return np.stack([nx, ny, nz, t1x, t1y, t1z, t2x, t2y, t2z, area])

    # i-faces: vtx0=(i,j), vtx1=(i,j+1); face i=0..nic -> padded right cell i+2
    g.face[0][:, 0, NG:NG + njc, NG:NG + nic + 1] = face2d(x[:-1, :], y[:-1, :], x[1:, :], y[1:, :])
    # j-faces: vtx0=(i+1,j), vtx1=(i,j)
    g.face[1][:, 0, NG:NG + njc + 1, NG:NG + nic] = face2d(x[:, 1:], y[:, 1:], x[:, :-1], y[:, :-1])
    for f in range(4):
        g.mirror_ghost_lengths(f)
    return g


def _tet_volume(p0, p1, p2, p3):
    d01 = [p1[m] - p0[m] for m in range(3)]
    d02 = [p2[m] - p0[m] for m in range(3)]
    c = _cross(*d01, *d02)
    d03 = [p3[m] - p0[m] for m in range(3)]
    return (d03[0] * c[0] + d03[1] * c[1] + d03[2] * c[2]) / 6.0


def _pyramid_volume(p0, p1, p2, p3, p4):
    pmB = [0.25 * (p0[m] + p1[m] + p2[m] + p3[m]) for m in range(3)]
    vol = 0.0
    vol = vol + _tet_volume(p0, p1, pmB, p4)
    vol = vol + _tet_volume(p1, p2, pmB, p4)
    vol = vol + _tet_volume(p2, p3, pmB, p4)
    vol = vol + _tet_volume(p3, p0, pmB, p4)
    return vol


def _quad_face(p0, p1, p2, p3):
    """quad_properties: n, t1, t2, area from the vertex cycle p0..p3."""
    p01 = [p1[m] - p0[m] + p2[m] - p3[m] for m in range(3)]
    p03 = [p3[m] - p0[m] + p2[m] - p1[m] for m in range(3)]
    va = _cross(*p01, *p03)
    va = [0.25 * v for v in va]
    area = np.sqrt(va[0] * va[0] + va[1] * va[1] + va[2] * va[2])
    n = [v / area for v in va]
    mag = np.sqrt(p01[0] * p01[0] + p01[1] * p01[1] + p01[2] * p01[2])
    t1 = [v / mag for v in p01]                      # Vector3.normalize
    t2 = _cross(*n, *t1)
    mag2 = np.sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2])
    t2 = [v / mag2 for v in t2]
    return np.stack(n + t1 + list(t2) + [area])


def geometry_3d(x, y, z):
    """x, y, z: vertex coordinates, shape (nkv, njv, niv)."""
    P = [np.asarray(a, dtype=np.float64) for a in (x, y, z)]
    nkv, njv, niv = P[0].shape
    nic, njc, nkc = niv - 1, njv - 1, nkv - 1
    g = BlockGeometry(3, nic, njc, nkc)

    def V(di, dj, dk):
        """Vertex (i+di, j+dj, k+dk) of every cell (i, j, k)."""
        return [a[dk:dk + nkc, dj:dj + njc, di:di + nic] for a in P]

    p0, p1, p2, p3 = V(0, 0, 0), V(1, 0, 0), V(1, 1, 0), V(0, 1, 0)
    p4, p5, p6, p7 = V(0, 0, 1), V(1, 0, 1), V(1, 1, 1), V(0, 1, 1)
    cen = [0.125 * (p0[m] + p1[m] + p2[m] + p3[m] + p4[m] + p5[m] + p6[m] + p7[m]) for m in range(3)]
    pmN = [0.25 * (p3[m] + p2[m] + p6[m] + p7[m]) for m in range(3)]
    pmE = [0.25 * (p1[m] + p2[m] + p6[m] + p5[m]) for m in range(3)]
    pmS = [0.25 * (p0[m] + p1[m] + p5[m] + p4[m]) for m in range(3)]
    pmW = [0.25 * (p0[m] + p3[m] + p7[m] + p4[m]) for m in range(3)]
    pmT = [0.25 * (p4[m] + p5[m] + p6[m] + p7[m]) for m in range(3)]
    pmB = [0.25 * (p0[m] + p1[m] + p2[m] + p3[m]) for m in range(3)]

    def dist(a, b):
        dx, dy, dz = a[0] - b[0], a[1] - b[1], a[2] - b[2]
        return np.sqrt(dx * dx + dy * dy + dz * dz)

    iLen, jLen, kLen = dist(pmE, pmW), dist(pmN, pmS), dist(pmT, pmB)
    vol = 0.0
    vol = vol + _pyramid_volume(p6, p7, p3, p2, cen)
    vol = vol + _pyramid_volume(p5, p6, p2, p1, cen)
    vol = vol + _pyramid_volume(p4, p5, p1, p0, cen)
    vol = vol + _pyramid_volume(p7, p4, p0, p3, cen)
    vol = vol + _pyramid_volume(p7, p6, p5, p4, cen)
    vol = vol + _pyramid_volume(p0, p1, p2, p3, cen)
    if np.any(vol <= 0.0):
        raise ValueError("Invalid (non-positive) cell volume")
    inner = (slice(NG, NG + nkc), slice(NG, NG + njc), slice(NG, NG + nic))
    g.vol[inner] = vol
    g.len[0][inner], g.len[1][inner], g.len[2][inner] = iLen, jLen, kLen
    for m in range(3):
        g.pos[m][inner] = cen[m]

    def at(di, dj, dk, ni, nj, nk):
        return [a[dk:dk + nk, dj:dj + nj, di:di + ni] for a in P]

    # i-faces (i=0..nic): (i,j,k) (i,j+1,k) (i,j+1,k+1) (i,j,k+1)
    g.face[0][:, NG:NG + nkc, NG:NG + njc, NG:NG + nic + 1] = _quad_face(
        at(0, 0, 0, niv, njc, nkc), at(0, 1, 0, niv, njc, nkc), at(0, 1, 1, niv, njc, nkc), at(0, 0, 1, niv, njc, nkc))
    # j-faces (j=0..njc): (i,j,k) (i,j,k+1) (i+1,j,k+1) (i+1,j,k)
    g.face[1][:, NG:NG + nkc, NG:NG + njc + 1, NG:NG + nic] = _quad_face(
        at(0, 0, 0, nic, njv, nkc), at(0, 0, 1, nic, njv, nkc), at(1, 0, 1, nic, njv, nkc), at(1, 0, 0, nic, njv, nkc))
    # k-faces (k=0..nkc): (i,j,k) (i+1,j,k) (i+1,j+1,k) (i,j+1,k)
    g.face[2][:, NG:NG + nkc + 1, NG:NG + njc, NG:NG + nic] = _quad_face(
        at(0, 0, 0, nic, njc, nkv), at(1, 0, 0, nic, njc, nkv), at(1, 1, 0, nic, njc, nkv), at(0, 1, 0, nic, njc, nkv))
    for f in range(6):
        g.mirror_ghost_lengths(f)
    return g
