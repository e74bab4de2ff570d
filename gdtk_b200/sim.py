"""Host-side mirror of the part of Eilmer's job-script API and time-marching
loop that sits either side of the accelerated path.

Names follow the reference so a job script reads the same:

* ``config.*`` options and defaults      src/eilmer/globalconfig.d:923-1130,1262-1295
* ``FluidBlock:new{grid=, initialState=, bcList=}``  src/eilmer/fluidblock.lua
* boundary conditions                    src/eilmer/bc.lua (WallBC_WithSlip :757,
  InFlowBC_Supersonic :1320, OutFlowBC_SimpleExtrapolate :1580, OutFlowBC_Simple :1601-1625,
  ExchangeBC_FullFace :1672)
* ``identifyBlockConnections()``         src/eilmer/prep.lua / src/geom/grid/sgrid.d
* ``determine_time_step_size`` policy    src/eilmer/simcore_gasdynamic_step.d:60-159
* step retry                             src/eilmer/simcore_gasdynamic_step.d:995-999,1545-1554
* main loop                              src/eilmer/simcore.d:1014-1331

All numerical work of a time step happens inside the C-ABI library
(include/eb200.h); this module only prepares arrays and keeps the dt policy.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .geometry import NG, geometry_2d, geometry_3d, BlockGeometry
from .gas import FlowState


class Config:
    """The ``config`` table of a job script; attribute names and defaults are Eilmer's."""

    def __init__(self, **kw):
        self.title = ""
        self.dimensions = 2
        self.axisymmetric = False
        self.flux_calculator = "adaptive_hanel_ausmdv"   # globalconfig.d:1031 (see check())
        self.interpolation_order = 2
        self.apply_limiter = True                          # :1077
        self.extrema_clipping = True                       # :1078
        self.epsilon_van_albada = 1.0e-12                  # :1079
        self.interpolate_in_local_frame = True
        self.apply_entropy_fix = True                      # :1088
        self.thermo_interpolator = "rhou"                  # :1073
        self.M_inf = 0.01                                  # :1114
        self.compression_tolerance = -0.30                 # :1123 (PJ shock detector)
        self.shear_tolerance = 0.20                        # :1108
        self.strict_shock_detector = True                  # :1064
        self.shock_detector = "PJ"                         # :1124
        self.shock_detector_smoothing = 0                  # :1127
        self.gasdynamic_update_scheme = "predictor-corrector"  # :936
        self.cfl_value = 0.5
        # config.cfl_schedule = {{t0, cfl0}, {t1, cfl1}, ...}: when given it replaces cfl_value, as in the reference's
        # write_config_file (output.lua:173-198); the .control file's cfl_scale_factor multiplies either
        self.cfl_schedule = None
        self.cfl_scale_factor = 1.0
        self.cfl_count = 10                                # :1294
        self.fixed_time_step = False
        self.dt_init = 1.0e-3                              # :1281
        self.dt_max = 1.0e-3                               # :1282
        self.max_time = 1.0e-3
        self.dt_history = 1.0e-3                           # history cells are sampled at this interval of simulated time
        self.dt_plot = 1.0e-3                              # flow solutions are written at this interval (simcore.d:1196-1216)
        self.max_step = 100
        self.max_attempts_for_step = 3                     # :947
        self.max_invalid_cells = 0                         # :1014
        self.flowstate_limits_max_velocity = 30000.0       # :79-86
        self.flowstate_limits_max_temp = 50000.0
        self.flowstate_limits_min_temp = 0.0
        self.ignore_low_T_thermo_update_failure = True     # :1005
        self.suggested_low_T_value = 200.0                 # :1006
        self.viscous = False
        self.reacting = False
        # Not an Eilmer option: selects the FMA-free kernel build (see include/eb200.h).
        self.strict_fp = False
        # Not an Eilmer option: testing knob, never use the uniform-Cartesian fast path.
        self.force_general_path = False
        self.force_generic_kernel = False  # testing knob: never use the tuned flux kernel
        # "eilmer4" (src/eilmer, the default) or "lmr": Eilmer 5's formulas where they change numbers on this path
        # (scaled van Albada epsilon, smooth-maximum sound speed in AUSMDV, no thermo fall-back)
        self.solver_variant = "eilmer4"
        self.no_tma = False                # testing knob: stage tiles with cp.async instead of TMA
        self.no_push = False               # testing knob: all ghost cells are filled by the ghost-cell kernel
        self.block_index = None          # optional {block id: (ib, jb, kb)} left by the case factories
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(f"unknown config option {k!r}")
            setattr(self, k, v)

    def cfl_at(self, t):
        """cfl_schedule.interpolate_value(SimState.time) * cfl_scale_factor (simcore_gasdynamic_step.d:77-78, with
        Schedule.interpolate_value of src/nm/schedule.d:39-56: constant outside the table, linear inside)."""
        sched = self.cfl_schedule or [(0.0, self.cfl_value)]
        times = [float(p[0]) for p in sched]
        values = [float(p[1]) for p in sched]
        if t <= times[0]:
            v = values[0]
        elif t >= times[-1]:
            v = values[-1]
        else:
            i = len(times) - 1
            while i > 0 and t < times[i]:
                i -= 1
            frac = (t - times[i]) / (times[i + 1] - times[i])
            v = (1.0 - frac) * values[i] + frac * values[i + 1]
        return v * self.cfl_scale_factor

    def check(self):
        fc = self.flux_calculator
        if fc not in _abi.FLUX_CALCULATORS:
            raise ValueError(
                f"config.flux_calculator={fc!r} is not on the accelerated path; choose one of "
                f"{sorted(_abi.FLUX_CALCULATORS)}")
        if fc.startswith("adaptive"):
            if self.shock_detector != "PJ" or self.shock_detector_smoothing != 0:
                raise ValueError("only the PJ shock detector without smoothing (the defaults) is on this path")
            if self.compression_tolerance > 0.0:
                raise ValueError("compression_tolerance should be negative!")
        if self.gasdynamic_update_scheme not in _abi.UPDATE_SCHEMES:
            raise ValueError(f"gasdynamic_update_scheme {self.gasdynamic_update_scheme!r} not supported")
        if self.interpolation_order not in (1, 2):
            raise ValueError("interpolation_order must be 1 or 2 on this path")
        if self.thermo_interpolator not in _abi.THERMO_INTERPOLATORS:
            raise ValueError(f"unknown thermo_interpolator {self.thermo_interpolator!r}")
        if self.viscous or self.reacting:
            raise ValueError("viscous / reacting flow is outside the accelerated path")
        if self.axisymmetric and self.dimensions != 2:
            raise ValueError("axisymmetric needs dimensions=2")

    def to_struct(self, gmodel, rank=0, device=0):
        self.check()
        c = _abi.Config()
        c.dimensions = self.dimensions
        c.axisymmetric = int(self.axisymmetric)
        c.flux_calculator = _abi.FLUX_CALCULATORS[self.flux_calculator]
        c.interpolation_order = self.interpolation_order
        c.apply_limiter = int(self.apply_limiter)
        c.extrema_clipping = int(self.extrema_clipping)
        c.interpolate_in_local_frame = int(self.interpolate_in_local_frame)
        c.apply_entropy_fix = int(self.apply_entropy_fix)
        c.update_scheme = _abi.UPDATE_SCHEMES[self.gasdynamic_update_scheme]
        c.max_invalid_cells = self.max_invalid_cells
        c.strict_fp = int(self.strict_fp)
        c.thermo_interpolator = _abi.THERMO_INTERPOLATORS[self.thermo_interpolator]
        if self.solver_variant not in ("eilmer4", "lmr"):
            raise ValueError(f"unknown solver_variant {self.solver_variant!r}")
        c.solver_variant = 1.0 if self.solver_variant == "lmr" else 0.0
        c.reserved_i[0] = int(self.force_general_path)
        c.reserved_i[1] = int(self.force_generic_kernel)
        c.reserved_i[2] = int(self.no_tma)
        c.reserved_i[3] = int(self.no_push)
        c.rank = rank
        c.device = device
        c.epsilon_van_albada = self.epsilon_van_albada
        c.M_inf = self.M_inf
        c.max_velocity = self.flowstate_limits_max_velocity
        c.max_temp = self.flowstate_limits_max_temp
        c.min_temp = self.flowstate_limits_min_temp
        c.suggested_low_T_value = self.suggested_low_T_value
        c.ignore_low_T_thermo_update_failure = int(self.ignore_low_T_thermo_update_failure)
        c.strict_shock_detector = int(self.strict_shock_detector)
        c.compression_tolerance = self.compression_tolerance
        c.shear_tolerance = self.shear_tolerance
        gmodel.fill_config(c)
        return c


# ---------------------------------------------------------------------------
# Boundary conditions (src/eilmer/bc.lua)

class BoundaryCondition:
    kind = None

    def params(self):
        return []


class WallBC_WithSlip(BoundaryCondition):
    kind = _abi.BC_WALL_WITH_SLIP


class WallBC_WithSlip1(BoundaryCondition):
    """bc.lua:783-806: a slip wall WITHOUT ghost-cell data (ghost_cell_data_available = false, no preReconAction).
    The faces next to it are reconstructed from one-sided stencils (onedinterp.d:117-273: l0r2 / l1r2 / l2r1 / l2r0)
    and the wall face takes the one-sided wall flux (fluxcalc.d:187-385)."""
    kind = _abi.BC_WALL_WITH_SLIP1


class UserDefinedBC(BoundaryCondition):
    """bc.lua UserDefinedBC whose ``ghostCells`` function depends on the position of the ghost cell only (the vortex
    inflow of examples/eilmer/2D/vortex-supersonic/udf-vortex-flow.lua): ``ghost_state(x, y, z) -> FlowState``.
    bc/user_defined_effects.d:237-310 calls the Lua function with the ghost-cell centres every stage; a function of
    position gives the same FlowStates every time, so it is evaluated once, at set-up, at the centres the reference
    gives its ghost cells (sfluidblock.d:897-1086: linear extrapolation from the two cells inside)."""
    kind = _abi.BC_GHOST_PROFILE

    def __init__(self, ghost_state):
        self.ghost_state = ghost_state

    def params_for(self, geom, face):
        d, hi = face // 2, face & 1
        n = (geom.nic, geom.njc, geom.nkc)
        off = (NG, NG, geom.kg)
        d1, d2 = (d + 1) % 3, (d + 2) % 3
        out = []
        for a2 in range(n[d2]):
            for a1 in range(n[d1]):
                def centre(m):                      # interior cell m layers in from the face
                    idx = [0, 0, 0]
                    idx[d], idx[d1], idx[d2] = (n[d] - 1 - m if hi else m), a1, a2
                    return geom.pos[:, idx[2] + off[2], idx[1] + off[1], idx[0] + off[0]]
                c1, c2 = centre(0), centre(1)
                g0 = 2.0 * c1 - c2                  # extrap(ghost, cell_1, cell_2)
                g1 = 2.0 * g0 - c1
                for pos in (g0, g1):
                    out += self.ghost_state(float(pos[0]), float(pos[1]), float(pos[2])).as_prims()
        return out


class InFlowBC_Supersonic(BoundaryCondition):
    kind = _abi.BC_INFLOW_SUPERSONIC

    def __init__(self, flowState):
        self.flowState = flowState

    def params(self):
        return self.flowState.as_prims()


class OutFlowBC_SimpleExtrapolate(BoundaryCondition):
    kind = _abi.BC_OUTFLOW_SIMPLE_EXTRAPOLATE


class OutFlowBC_SimpleFlux(BoundaryCondition):
    kind = _abi.BC_OUTFLOW_SIMPLE_FLUX


OutFlowBC_Simple = OutFlowBC_SimpleFlux   # bc.lua:1625


class OutFlowBC_FixedP(BoundaryCondition):
    """bc.lua:1627-1647: extrapolate, then fix the ghost-cell pressure."""
    kind = _abi.BC_OUTFLOW_FIXED_P

    def __init__(self, p_outside):
        self.p_outside = float(p_outside)

    def params(self):
        return [self.p_outside]


class OutFlowBC_FixedPT(BoundaryCondition):
    """bc.lua:1649-1670: extrapolate, then fix the ghost-cell pressure and temperature."""
    kind = _abi.BC_OUTFLOW_FIXED_PT

    def __init__(self, p_outside, T_outside):
        self.p_outside, self.T_outside = float(p_outside), float(T_outside)

    def params(self):
        return [self.p_outside, self.T_outside]


class ExchangeBC_FullFace(BoundaryCondition):
    kind = _abi.BC_EXCHANGE_FULL_FACE

    def __init__(self, otherBlock, otherFace, orientation=0, cell_map=None):
        self.otherBlock = otherBlock          # block id
        self.otherFace = otherFace if isinstance(otherFace, int) else _abi.FACE_NAMES.index(otherFace)
        self.orientation = orientation
        # int array (n2, n1, 2 layers, 3): source cell (i, j, k) in the other block of every ghost cell behind this
        # face, for connections other than opposite faces of aligned blocks (face_cell_map below); None: aligned
        self.cell_map = cell_map


# ---------------------------------------------------------------------------

class FluidBlock:
    """``FluidBlock:new{grid=..., initialState=..., bcList=...}``.

    grid: tuple of vertex arrays (x, y) with shape (njv, niv) or (x, y, z) with shape
    (nkv, njv, niv), or a ready BlockGeometry.  initialState: a FlowState, or a
    function (x, y, z) -> FlowState evaluated at cell centroids, or a dict of padded
    primitive arrays.
    """

    def __init__(self, grid, initialState, bcList=None, id=None):
        self.id = id
        self.grid = grid
        self.initialState = initialState
        self.bcList = dict(bcList or {})
        self.geom = None

    def corner_signature(self, dims):
        """Corner vertices of each face, used by identify_block_connections."""
        if isinstance(self.grid, BlockGeometry):
            return None
        P = [np.asarray(a) for a in self.grid]
        sig = {}
        if dims == 2:
            x, y = P
            pts = lambda j, i: (float(x[j, i]), float(y[j, i]))
            sig[_abi.WEST] = (pts(0, 0), pts(-1, 0))
            sig[_abi.EAST] = (pts(0, -1), pts(-1, -1))
            sig[_abi.SOUTH] = (pts(0, 0), pts(0, -1))
            sig[_abi.NORTH] = (pts(-1, 0), pts(-1, -1))
        else:
            x, y, z = P
            pt = lambda k, j, i: (float(x[k, j, i]), float(y[k, j, i]), float(z[k, j, i]))
            sig[_abi.WEST] = (pt(0, 0, 0), pt(0, -1, 0), pt(-1, 0, 0), pt(-1, -1, 0))
            sig[_abi.EAST] = (pt(0, 0, -1), pt(0, -1, -1), pt(-1, 0, -1), pt(-1, -1, -1))
            sig[_abi.SOUTH] = (pt(0, 0, 0), pt(0, 0, -1), pt(-1, 0, 0), pt(-1, 0, -1))
            sig[_abi.NORTH] = (pt(0, -1, 0), pt(0, -1, -1), pt(-1, -1, 0), pt(-1, -1, -1))
            sig[_abi.BOTTOM] = (pt(0, 0, 0), pt(0, 0, -1), pt(0, -1, 0), pt(0, -1, -1))
            sig[_abi.TOP] = (pt(-1, 0, 0), pt(-1, 0, -1), pt(-1, -1, 0), pt(-1, -1, -1))
        return sig


def _face_corners(grid, face):
    """Corner vertices of a 3D block face as P[alpha][beta], alpha/beta = low/high end of the in-face
    directions d1 = (d+1)%3 and d2 = (d+2)%3 (d = face // 2), and the cell counts (n1, n2) along them."""
    X, Y, Z = (np.asarray(a) for a in grid)
    nv = (X.shape[2], X.shape[1], X.shape[0])           # vertices along i, j, k
    d, hi = face // 2, face & 1
    d1, d2 = (d + 1) % 3, (d + 2) % 3
    P = [[None, None], [None, None]]
    for al in (0, 1):
        for be in (0, 1):
            idx = [0, 0, 0]
            idx[d] = nv[d] - 1 if hi else 0
            idx[d1] = al * (nv[d1] - 1)
            idx[d2] = be * (nv[d2] - 1)
            i, j, k = idx
            P[al][be] = (float(X[k, j, i]), float(Y[k, j, i]), float(Z[k, j, i]))
    return P, (nv[d1] - 1, nv[d2] - 1)


def face_cell_map(grid_a, face_a, grid_b, face_b, tol=1.0e-6):
    """If face_a of block A and face_b of block B coincide (their four corners match in some order), the
    source cell in B of every ghost cell behind face_a of A: int array of shape (n2, n1, NG, 3) with (i, j, k),
    indexed by the cell indices along A's in-face directions and the ghost layer.  None if they do not coincide.
    Replaces the case tables of full_face_copy.d:141-1380 by matching corners, which covers every pair of
    faces and all four rotations."""
    PA, (na1, na2) = _face_corners(grid_a, face_a)
    PB, (nb1, nb2) = _face_corners(grid_b, face_b)

    def find(p):
        for ga in (0, 1):
            for de in (0, 1):
                if _close(p, PB[ga][de], tol):
                    return ga, de
        return None
    c00, c10, c01, c11 = find(PA[0][0]), find(PA[1][0]), find(PA[0][1]), find(PA[1][1])
    if None in (c00, c10, c01, c11) or len({c00, c10, c01, c11}) != 4:
        return None
    a_runs_along_first = c10[0] != c00[0]           # does A's first in-face direction run along B's first?
    if a_runs_along_first:
        if c10[1] != c00[1] or c01[0] != c00[0]:
            return None
        if (na1, na2) != (nb1, nb2):
            return None
    else:
        if c10[0] != c00[0] or c01[1] != c00[1]:
            return None
        if (na1, na2) != (nb2, nb1):
            return None
    XB = np.asarray(grid_b[0])
    nB = (XB.shape[2] - 1, XB.shape[1] - 1, XB.shape[0] - 1)
    e, ehi = face_b // 2, face_b & 1
    e1, e2 = (e + 1) % 3, (e + 2) % 3
    out = np.zeros((na2, na1, NG, 3), dtype=np.int32)
    a1 = np.arange(na1)[None, :]
    a2 = np.arange(na2)[:, None]
    if a_runs_along_first:
        c = (nb1 - 1 - a1) if c00[0] else a1
        dd = (nb2 - 1 - a2) if c00[1] else a2
    else:
        dd = (nb2 - 1 - a1) if c00[1] else a1
        c = (nb1 - 1 - a2) if c00[0] else a2
    c, dd = np.broadcast_to(c, (na2, na1)), np.broadcast_to(dd, (na2, na1))
    for layer in range(NG):
        out[:, :, layer, e] = (nB[e] - 1 - layer) if ehi else layer
        out[:, :, layer, e1] = c
        out[:, :, layer, e2] = dd
    return out


def _close(p, q, tol):
    return all(abs(a - b) <= tol for a, b in zip(p, q))


def identify_block_connections(blocks, dims, tol=1.0e-6):
    """``identifyBlockConnections()``: connect coincident faces with ExchangeBC_FullFace.

    2D: any pair of faces whose end points coincide (either sense; the cell mapping of
    full_face_copy.d:704-870 handles the index reversal).  3D: opposite faces of aligned
    blocks (orientation 0).  Faces that already carry a boundary condition are left alone.
    """
    sigs = [b.corner_signature(dims) for b in blocks]
    for ia, A in enumerate(blocks):
        for ib, B in enumerate(blocks):
            if ib <= ia or sigs[ia] is None or sigs[ib] is None:
                continue
            for fa, pa in sigs[ia].items():
                for fb, pb in sigs[ib].items():
                    na, nb = _abi.FACE_NAMES[fa], _abi.FACE_NAMES[fb]
                    if na in A.bcList or nb in B.bcList:
                        continue
                    if dims == 2:
                        same = _close(pa[0], pb[0], tol) and _close(pa[1], pb[1], tol)
                        rev = _close(pa[0], pb[1], tol) and _close(pa[1], pb[0], tol)
                        # which sense is admissible is fixed by the face pair (full_face_copy.d:704-870)
                        is_rev = (fa in (_abi.NORTH, _abi.WEST)) == (fb in (_abi.NORTH, _abi.WEST))
                        ok = rev if is_rev else same
                    else:
                        ok = (fa ^ 1) == fb and all(_close(p, q, tol) for p, q in zip(pa, pb))
                        if not ok and all(any(_close(p, q, tol) for q in pb) for p in pa):
                            # the faces coincide some other way round: connect them through explicit cell maps
                            mab = face_cell_map(A.grid, fa, B.grid, fb, tol)
                            mba = face_cell_map(B.grid, fb, A.grid, fa, tol)
                            if mab is not None and mba is not None:
                                A.bcList[na] = ExchangeBC_FullFace(B.id, fb, 0, cell_map=mab)
                                B.bcList[nb] = ExchangeBC_FullFace(A.id, fa, 0, cell_map=mba)
                            continue
                    if ok:
                        A.bcList[na] = ExchangeBC_FullFace(B.id, fb, 0)
                        B.bcList[nb] = ExchangeBC_FullFace(A.id, fa, 0)


def full_face_source(dims, face, other_dims, other_face, t1, t2, layer):
    """Interior cell (i, j, k) of the other block that feeds ghost `layer` of this block's
    boundary-face cell with in-face indices (t1, t2).  2D: t1 runs along the boundary
    (j for east/west, i for north/south).  3D aligned: (t1, t2) are indices in the
    ((d+1)%3, (d+2)%3) directions.  Mapping of full_face_copy.d:704-870 and :958-1014."""
    onic, onjc, onkc = other_dims
    W, E, S, N = _abi.WEST, _abi.EAST, _abi.SOUTH, _abi.NORTH
    if dims == 2:
        t = t1
        rev = (face in (N, W)) == (other_face in (N, W))
        if other_face == N:
            return ((onic - t - 1) if rev else t, onjc - 1 - layer, 0)
        if other_face == E:
            return (onic - 1 - layer, (onjc - t - 1) if rev else t, 0)
        if other_face == S:
            return ((onic - t - 1) if rev else t, layer, 0)
        return (layer, (onjc - t - 1) if rev else t, 0)
    d = face // 2
    if (face ^ 1) != other_face:
        raise ValueError("3D connections must join opposite faces of aligned blocks (orientation 0)")
    idx = [0, 0, 0]
    idx[(d + 1) % 3] = t1
    idx[(d + 2) % 3] = t2
    on = (onic, onjc, onkc)
    idx[d] = on[d] - 1 - layer if (other_face & 1) else layer
    return tuple(idx)


def _as_dpp(arrays):
    ptrs = (_abi.DP * len(arrays))()
    for i, a in enumerate(arrays):
        ptrs[i] = a.ctypes.data_as(_abi.DP)
    return ptrs


class Simulation:
    """One job: config + gas model + blocks, bound to an implementation of the C ABI."""

    def __init__(self, config, gmodel, blocks, lib=None, rank=0, world_size=1, block_owner=None,
                 device=0, exchange=None):
        self.config = config
        self.gmodel = gmodel
        self.blocks = list(blocks)
        for n, b in enumerate(self.blocks):
            if b.id is None:
                b.id = n
        self.lib = lib if lib is not None else _abi.load_library()
        self.rank, self.world_size = rank, world_size
        self.block_owner = block_owner or {b.id: 0 for b in self.blocks}
        self.dims = config.dimensions
        self.nsp = gmodel.n_species
        self.nprim = _abi.NPRIM_BASE + (2 * self.nsp if self.nsp > 1 else 0)
        self.ncq = (5 if self.dims == 3 else 4) + (self.nsp if self.nsp > 1 else 0)
        self.n_stages = _abi.N_STAGES[_abi.UPDATE_SCHEMES[config.gasdynamic_update_scheme]]
        self._cfg_struct = config.to_struct(gmodel, rank, device)
        if any(isinstance(bc, WallBC_WithSlip1) for b in self.blocks for bc in b.bcList.values()):
            # walls without ghost-cell data are served by the generic kernel on general-metric blocks only
            # (include/eb200.h, EB200_BC_WALL_WITH_SLIP1): every block of the job, on every process -- whoever owns the
            # wall -- keeps its metric arrays and runs that kernel, so that N processes compute what one computes
            self._cfg_struct.reserved_i[0] = 1
            self._cfg_struct.reserved_i[1] = 1
        self.handle = self.lib.check(self.lib.init(C.byref(self._cfg_struct)), "init")
        self._exchange_cb = None
        self.time = 0.0
        self.step = 0
        self.dt_global = config.dt_init
        self.dt_allow = None
        self.cfl_max = 0.0
        self.dt_history = []
        # setHistoryPoint{ib=, i=, j=, k=}: (time, FlowState variables) of every history cell, taken each
        # config.dt_history of simulated time like write_history_cells_to_files (simcore.d:942,1235-1243)
        self.history_points = []
        self.history = {}
        self.t_history = None
        self._setup(exchange)

    # -- set-up ------------------------------------------------------------
    def _geom(self, blk):
        if blk.geom is None:
            if isinstance(blk.grid, BlockGeometry):
                blk.geom = blk.grid
            elif self.dims == 2:
                blk.geom = geometry_2d(blk.grid[0], blk.grid[1], self.config.axisymmetric)
            else:
                blk.geom = geometry_3d(*blk.grid)
        return blk.geom

    def is_local(self, blk):
        return self.block_owner[blk.id] == self.rank

    def _setup(self, exchange):
        lib, h = self.lib, self.handle
        byid = {b.id: b for b in self.blocks}
        self.byid = byid
        nfaces = 6 if self.dims == 3 else 4
        local = [b for b in self.blocks if self.is_local(b)]
        # Which blocks does this process need to know about? local ones + their neighbours.
        needed = {b.id for b in local}
        for b in local:
            for bc in b.bcList.values():
                if isinstance(bc, ExchangeBC_FullFace):
                    needed.add(bc.otherBlock)
        for b in self.blocks:
            if b.id not in needed:
                continue
            g = self._dims_of(b)
            lib.check(lib.block_create(h, b.id, g[0], g[1], g[2], self.block_owner[b.id]), "block_create")
        # ghost-cell lengths at connections come from the neighbour (full_face_copy.d:1576-1626)
        for b in local:
            g = self._geom(b)
            for f in range(nfaces):
                bc = b.bcList.get(_abi.FACE_NAMES[f])
                if isinstance(bc, ExchangeBC_FullFace):
                    self._copy_ghost_lengths(b, f, byid[bc.otherBlock], bc.otherFace)
        for b in local:
            g = self._geom(b)
            faces = (_abi.DP * 3)()
            keep = []
            for d in range(self.dims):
                a = np.ascontiguousarray(g.face[d])
                keep.append(a)
                faces[d] = a.ctypes.data_as(_abi.DP)
            arr = lambda a: np.ascontiguousarray(a).ctypes.data_as(_abi.DP)
            lib.check(lib.block_set_geometry(
                h, b.id, arr(g.vol), arr(g.areaxy), arr(g.len[0]), arr(g.len[1]),
                arr(g.len[2]) if self.dims == 3 else None, faces), "block_set_geometry")
            for f in range(nfaces):
                bc = b.bcList.get(_abi.FACE_NAMES[f]) or WallBC_WithSlip()
                p = bc.params_for(g, f) if hasattr(bc, "params_for") else bc.params()
                pa = (C.c_double * max(1, len(p)))(*p)
                if isinstance(bc, ExchangeBC_FullFace):
                    lib.check(lib.block_set_bc(h, b.id, f, bc.kind, pa, 0, bc.otherBlock, bc.otherFace,
                                               bc.orientation), "block_set_bc")
                else:
                    lib.check(lib.block_set_bc(h, b.id, f, bc.kind, pa, len(p), -1, -1, 0), "block_set_bc")
        # explicit cell maps of connections that are not between opposite faces of aligned blocks: every process
        # declares them for its own blocks and for the faces of other processes' blocks that meet its own
        def declare_map(blk, f, bc):
            m = np.ascontiguousarray(bc.cell_map, dtype=np.int32)
            lib.check(lib.block_set_face_map(h, blk.id, f, m.ctypes.data_as(C.POINTER(C.c_int)), m.size // 3), "block_set_face_map")
        local_ids = {b.id for b in local}
        for b in self.blocks:
            if b.id not in needed:
                continue
            for f in range(nfaces):
                bc = b.bcList.get(_abi.FACE_NAMES[f])
                if not isinstance(bc, ExchangeBC_FullFace) or bc.cell_map is None:
                    continue
                if b.id in local_ids:
                    declare_map(b, f, bc)
                elif bc.otherBlock in local_ids:
                    lib.check(lib.block_set_bc(h, b.id, f, bc.kind, (C.c_double * 1)(), 0, bc.otherBlock, bc.otherFace, 0), "block_set_bc")
                    declare_map(b, f, bc)
        if exchange is not None:
            self._exchange_cb = _abi.EXCHANGE_FN(exchange)
            lib.check(lib.set_exchange(h, self._exchange_cb, None), "set_exchange")
        lib.check(lib.commit(h), "commit")
        for b in local:
            prims = self._initial_prims(b)
            lib.check(lib.upload_flow(h, b.id, _as_dpp(prims), len(prims)), "upload_flow")
        self.local_blocks = local
        self.n_local_cells = sum(np.prod(self._dims_of(b)) for b in local)

    def _dims_of(self, b):
        if isinstance(b.grid, BlockGeometry):
            return (b.grid.nic, b.grid.njc, b.grid.nkc)
        shp = np.shape(b.grid[0])
        if self.dims == 2:
            return (shp[1] - 1, shp[0] - 1, 1)
        return (shp[2] - 1, shp[1] - 1, shp[0] - 1)

    def _copy_ghost_lengths(self, b, face, other, other_face):
        """Ghost cells behind a connection take the neighbour's cell lengths
        (full_face_copy.d:1576-1626), with the cell mapping of full_face_source()."""
        g, og = self._geom(b), self._geom(other)
        if og is g:
            return            # blocks sharing one (uniform) geometry object: nothing to change
        d, hi = face // 2, face & 1
        n = (g.nic, g.njc, g.nkc)
        on = (og.nic, og.njc, og.nkc)
        bc = b.bcList.get(_abi.FACE_NAMES[face])
        if getattr(bc, "cell_map", None) is not None:
            # only the length along the face normal of a ghost cell is ever used (stencils are one-dimensional):
            # it is the neighbour's cell length along ITS face normal
            e = other_face // 2
            m = bc.cell_map
            d1, d2 = (d + 1) % 3, (d + 2) % 3
            off = (NG, NG, g.kg)
            a1, a2 = np.meshgrid(np.arange(n[d1]), np.arange(n[d2]))
            for layer in range(NG):
                idx = [None, None, None]
                idx[d] = np.full(a1.shape, (n[d] + layer if hi else -1 - layer) + off[d])
                idx[d1], idx[d2] = a1 + off[d1], a2 + off[d2]
                src = (m[:, :, layer, 2] + og.kg, m[:, :, layer, 1] + NG, m[:, :, layer, 0] + NG)
                g.len[d][idx[2], idx[1], idx[0]] = og.len[e][src]
            return
        for layer in range(NG):
            if self.dims == 3:
                # aligned, orientation 0: in-face indices carry over unchanged
                dst = [slice(g.kg, g.kg + g.nkc), slice(NG, NG + g.njc), slice(NG, NG + g.nic)]
                src = [slice(og.kg, og.kg + og.nkc), slice(NG, NG + og.njc), slice(NG, NG + og.nic)]
                off = (NG, NG, g.kg)[d]
                ooff = (NG, NG, og.kg)[d]
                dst[2 - d] = off + (n[d] + layer if hi else -1 - layer)
                src[2 - d] = ooff + (on[d] - 1 - layer if (other_face & 1) else layer)
                for m in range(3):
                    g.len[m][tuple(dst)] = og.len[m][tuple(src)]
            else:
                nt = n[1] if d == 0 else n[0]
                t = np.arange(nt)
                rev = (face in (_abi.NORTH, _abi.WEST)) == (other_face in (_abi.NORTH, _abi.WEST))
                tt = (on[1 - (other_face // 2)] - t - 1) if rev else t
                if other_face // 2 == 0:     # east/west of the other block: i fixed by layer, j runs
                    si = np.full(nt, on[0] - 1 - layer if other_face == _abi.EAST else layer)
                    sj = tt
                else:
                    sj = np.full(nt, on[1] - 1 - layer if other_face == _abi.NORTH else layer)
                    si = tt
                if d == 0:
                    di = np.full(nt, n[0] + layer if hi else -1 - layer)
                    dj = t
                else:
                    dj = np.full(nt, n[1] + layer if hi else -1 - layer)
                    di = t
                for m in range(3):
                    g.len[m][0, dj + NG, di + NG] = og.len[m][0, sj + NG, si + NG]

    def _initial_prims(self, b):
        g = self._geom(b)
        shp = (g.NK, g.NJ, g.NI)
        init = b.initialState
        if hasattr(init, "padded_arrays"):        # lazy generator: arrays are made when needed and dropped
            return [np.ascontiguousarray(a, dtype=np.float64).reshape(shp) for a in init.padded_arrays(g)]
        if isinstance(init, dict):
            names = ["rho", "u", "p", "T", "a", "velx", "vely", "velz"]
            return [np.ascontiguousarray(init[k], dtype=np.float64).reshape(shp) for k in names] + \
                   [np.ascontiguousarray(a, dtype=np.float64).reshape(shp) for a in init.get("species", [])]
        if isinstance(init, (list, tuple)):
            return [np.ascontiguousarray(a, dtype=np.float64).reshape(shp) for a in init]
        if isinstance(init, FlowState):
            vals = init.as_prims()
            return [np.full(shp, v) for v in vals]
        # function of position
        prims = [np.zeros(shp) for _ in range(self.nprim)]
        for k in range(g.kg, g.kg + g.nkc):
            for j in range(NG, NG + g.njc):
                for i in range(NG, NG + g.nic):
                    fs = init(g.pos[0][k, j, i], g.pos[1][k, j, i], g.pos[2][k, j, i])
                    for v, val in enumerate(fs.as_prims()):
                        prims[v][k, j, i] = val
        return prims

    # -- data access -------------------------------------------------------
    def download_flow(self, blk_id):
        g = self._geom(self.byid[blk_id])
        out = [np.zeros((g.NK, g.NJ, g.NI)) for _ in range(self.nprim)]
        self.lib.check(self.lib.download_flow(self.handle, blk_id, _as_dpp(out), self.nprim), "download_flow")
        return out

    def download_conserved(self, blk_id):
        g = self._geom(self.byid[blk_id])
        out = [np.zeros((g.NK, g.NJ, g.NI)) for _ in range(self.ncq)]
        self.lib.check(self.lib.download_conserved(self.handle, blk_id, _as_dpp(out), self.ncq), "download_conserved")
        return out

    def interior(self, blk_id, a):
        g = self._geom(self.byid[blk_id])
        return a[g.kg:g.kg + g.nkc, NG:NG + g.njc, NG:NG + g.nic]

    # -- time marching -----------------------------------------------------
    def compute_dt(self, check_cfl):
        out = (C.c_double * 3)()
        self.lib.check(self.lib.compute_dt(self.handle, self.dt_global, self.config.cfl_at(self.time),
                                           int(check_cfl), out), "compute_dt")
        return out[0], out[1]

    def reduce_dt(self, dt_allow, cfl_max):
        """Hook for multi-process runs: min/max over ranks (MPI_Allreduce at :105-107)."""
        return dt_allow, cfl_max

    def determine_time_step_size(self):
        """simcore_gasdynamic_step.d:60-159 (global time stepping branch)."""
        cfg = self.config
        if (self.step % cfg.cfl_count) != 0:
            return
        dt_allow, cfl_max = self.compute_dt(self.step > 0)
        dt_allow, cfl_max = self.reduce_dt(dt_allow, cfl_max)
        self.cfl_max = cfl_max
        if self.step == 0:
            dt_allow = min(cfg.dt_init, dt_allow)
        self.dt_allow = dt_allow
        if dt_allow <= self.dt_global:
            self.dt_global = dt_allow
        else:
            self.dt_global = min(self.dt_global * 1.5, dt_allow)
            self.dt_global = min(self.dt_global, cfg.dt_max)

    def reduce_step_status(self, rc):
        """Hook for multi-process runs: the worst status over ranks (MPI_Allreduce of step_failed, :1545-1554)."""
        return rc

    def gasdynamic_step(self):
        """One step with the reference's retry policy (:988-999, :1545-1554).  The decision is collective: if the
        step failed on any rank, every rank restores the start-of-step state (a rank whose own step went through
        takes it back), scales dt and tries again; a fatal error on one rank ends the run on all of them."""
        nbad = C.c_int(0)
        attempt = 0
        while True:
            attempt += 1
            rc_local = self.lib.step(self.handle, self.time, self.dt_global, C.byref(nbad))
            rc = self.reduce_step_status(rc_local)
            if rc < 0:
                why = self.lib.error() if rc_local < 0 else "another rank reported a fatal error"
                raise RuntimeError(f"step failed fatally ({rc}): {why}")
            if rc == 0:
                break
            if rc_local == 0:
                self.lib.check(self.lib.undo_step(self.handle), "undo_step")
            self.dt_global *= 0.2
            if attempt >= self.config.max_attempts_for_step:
                raise RuntimeError(
                    f"Explicit update failed after {self.config.max_attempts_for_step} attempts; giving up.")
        self.time += self.dt_global
        self.step += 1
        self.dt_history.append(self.dt_global)

    def set_history_point(self, ib, i, j, k=0):
        """setHistoryPoint{ib=, i=, j=, k=} for a block owned by this process."""
        key = (int(ib), int(i), int(j), int(k))
        self.history_points.append(key)
        self.history[key] = []
        return key

    def probe_cells(self, points):
        """FlowState variables (EB200_PRIM order) of interior cells [(ib, i, j, k), ...]."""
        n = len(points)
        ids = (C.c_int * n)(*[p[0] for p in points])
        ijk = (C.c_int * (3 * n))(*[v for p in points for v in p[1:4]])
        out = np.zeros((n, self.nprim))
        self.lib.check(self.lib.probe_cells(self.handle, n, ids, ijk, out.ctypes.data_as(_abi.DP), self.nprim), "probe_cells")
        return out

    def write_history(self):
        if self.history_points:
            for key, row in zip(self.history_points, self.probe_cells(self.history_points)):
                self.history[key].append((self.time,) + tuple(float(v) for v in row))

    def run(self, max_step=None, max_time=None):
        """integrate_in_time (simcore.d:1014-1331) without file IO; history cells are sampled every
        config.dt_history (simcore.d:942,1235-1243)."""
        max_step = self.config.max_step if max_step is None else max_step
        max_time = self.config.max_time if max_time is None else max_time
        if self.t_history is None:
            self.t_history = self.time + self.config.dt_history
        while self.time < max_time and self.step < max_step:
            if not self.config.fixed_time_step:
                self.determine_time_step_size()
            self.gasdynamic_step()
            if self.history_points and self.time >= self.t_history:
                self.write_history()
                self.t_history += self.config.dt_history
        return self.step

    def run_fixed(self, nsteps, dt):
        """nsteps of fixed dt without host synchronisation between steps (bench loop)."""
        nbad = C.c_int(0)
        rc = self.lib.run_steps(self.handle, self.time, dt, nsteps, C.byref(nbad))
        if rc != 0:
            raise RuntimeError(f"run_steps returned {rc}: {self.lib.error()}")
        self.time += nsteps * dt
        self.step += nsteps

    def kernel_launches(self):
        return int(self.lib.kernel_launches(self.handle))

    def flux_kernel_time(self, reset=False):
        ms = C.c_double(0.0)
        n = C.c_longlong(0)
        self.lib.check(self.lib.flux_kernel_time(self.handle, int(reset), C.byref(ms), C.byref(n)), "flux_kernel_time")
        return ms.value, n.value

    def close(self):
        if self.handle is not None:
            self.lib.finalize(self.handle)
            self.handle = None
