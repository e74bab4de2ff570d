"""Set-up geometry (host side): the uniform fast constructor must equal the general formulas
bit for bit on dyadic grids; general grids give positive volumes and orthonormal frames."""
import numpy as np

from gdtk_b200.geometry import geometry_2d, geometry_3d, NG
from gdtk_b200.grids import box_grid_2d, box_grid_3d, uniform_box_geometry, quad_patch_grid


def test_uniform_3d_matches_general_formulas():
    n, h = 8, 1.0 / 64
    g1 = geometry_3d(*box_grid_3d((0.25, 0.5, 0.0), (0.25 + n * h, 0.5 + n * h, n * h), n, n, n))
    g2 = uniform_box_geometry(3, n, n, n, h, h, h)
    inner = (slice(NG, NG + n),) * 3
    assert np.array_equal(g1.vol[inner], g2.vol[inner])
    for d in range(3):
        assert np.array_equal(g1.len[d], g2.len[d])
        sl = [slice(NG, NG + n)] * 3
        sl[2 - d] = slice(NG, NG + n + 1)
        assert np.array_equal(g1.face[d][(slice(None),) + tuple(sl)], g2.face[d][(slice(None),) + tuple(sl)])


def test_uniform_2d_matches_general_formulas_including_signed_zeros():
    nx, ny, dx, dy = 12, 6, 3.0 / 4096, 1.0 / 1024
    g1 = geometry_2d(*box_grid_2d(0.0, nx * dx, 0.0, ny * dy, nx, ny))
    g2 = uniform_box_geometry(2, nx, ny, 1, dx, dy)
    assert np.array_equal(g1.vol[0, NG:NG + ny, NG:NG + nx], g2.vol[0, NG:NG + ny, NG:NG + nx])
    a = g1.face[0][:, 0, NG:NG + ny, NG:NG + nx + 1]
    b = g2.face[0][:, 0, NG:NG + ny, NG:NG + nx + 1]
    assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))
    a = g1.face[1][:, 0, NG:NG + ny + 1, NG:NG + nx]
    b = g2.face[1][:, 0, NG:NG + ny + 1, NG:NG + nx]
    assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))


def test_sheared_3d_grid_frames_and_volumes():
    X, Y, Z = box_grid_3d((0, 0, 0), (1, 1, 1), 6, 5, 4)
    X = X + 0.2 * Z
    g = geometry_3d(X, Y, Z)
    v = g.vol[NG:NG + 4, NG:NG + 5, NG:NG + 6]
    assert np.all(v > 0) and abs(v.sum() - 1.0) < 1e-12
    for d in range(3):
        sl = [slice(NG, NG + 4), slice(NG, NG + 5), slice(NG, NG + 6)]
        f = g.face[d][(slice(None),) + tuple(sl)]
        n, t1, t2 = f[0:3], f[3:6], f[6:9]
        for a, b, val in ((n, n, 1), (t1, t1, 1), (t2, t2, 1), (n, t1, 0), (n, t2, 0), (t1, t2, 0)):
            assert np.allclose((a * b).sum(0), val, atol=1e-13)
        assert np.allclose(np.cross(n, t1, axis=0), t2, atol=1e-13)    # right-handed


def test_axisymmetric_volume_and_ghost_lengths():
    x, y = quad_patch_grid((0.2, 0.0), (1.0, 0.29118), (1.0, 1.0), (0.2, 1.0), 30, 40)
    g = geometry_2d(x, y, axisymmetric=True)
    vol = g.vol[0, NG:NG + 40, NG:NG + 30]
    assert np.all(vol > 0)
    # Pappus: sum of volumes per radian = integral of y dA
    area = g.areaxy[0, NG:NG + 40, NG:NG + 30]
    assert abs(vol.sum() - (area * g.pos[1][0, NG:NG + 40, NG:NG + 30]).sum()) < 1e-15
    # ghost n <- interior n (sfluidblock.d:902-933)
    assert np.array_equal(g.len[0][0, NG:NG + 40, 1], g.len[0][0, NG:NG + 40, NG])
    assert np.array_equal(g.len[0][0, NG:NG + 40, 0], g.len[0][0, NG:NG + 40, NG + 1])
    assert np.array_equal(g.len[1][0, NG + 40 + 1, NG:NG + 30], g.len[1][0, NG + 40 - 2, NG:NG + 30])


def test_roberts_cluster_function_matches_the_reference_unit_test():
    """src/geom/misc/univariatefunctions.d:814-816: RobertsFunction(false, true, 1.1)."""
    from gdtk_b200.grids import roberts_function
    cf = roberts_function(False, True, 1.1)
    assert abs(float(cf(0.1)) - 0.166167) < 1.0e-4 * 0.166167 + 1e-9
    assert abs(float(cf(0.9)) - 0.96657) < 1.0e-4 * 0.96657 + 1e-9
    assert float(cf(0.0)) == 0.0 and abs(float(cf(1.0)) - 1.0) < 1e-15
    # clustering towards the other end is the mirror image
    rev = roberts_function(True, False, 1.1)
    assert abs(float(rev(0.9)) - (1.0 - float(cf(0.1)))) < 1e-15


def test_hexahedron_grid_is_the_trilinear_map_of_its_corners():
    from gdtk_b200.grids import hex_volume_grid
    c = [[0, 0, 0], [2, 0, 0.5], [2, 1, 0.5], [0, 1, 0], [0, 0, 1], [2, 0, 1], [2, 1, 1], [0, 1, 1]]
    X, Y, Z = hex_volume_grid(c, 5, 3, 4)
    assert X.shape == (4, 3, 5)
    assert (X[0, 0, 0], Y[0, 0, 0], Z[0, 0, 0]) == (0.0, 0.0, 0.0) and (X[-1, -1, -1], Y[-1, -1, -1], Z[-1, -1, -1]) == (2.0, 1.0, 1.0)
    assert abs(Z[0, 0, -1] - 0.5) < 1e-15 and abs(Z[0, 1, 2] - 0.25) < 1e-15      # bottom face rises linearly along x
    assert np.allclose(X[:, :, 2], 1.0)
