"""ctypes binding of the C ABI declared in include/eb200.h.

The product library is ``gdtk_b200/csrc/libeb200.so`` (hand-written CUDA for
sm_100a).  There is no CPU fallback: if the library is missing, loading fails
loudly.  ``load_library`` takes the path and symbol prefix explicitly so that
the TESTS can bind the CPU oracle (``oracle/_build/liboracle.so``, prefix
``orc_``) through the very same host classes; product code only ever uses the
defaults.
"""
import ctypes as C
import os

NGHOST = 2
MAX_SPECIES = 8
MAX_SEGMENTS = 4
NPRIM_BASE = 8

# enum eb200_face (src/geom/elements/nomenclature.d:12-19 in the reference)
WEST, EAST, SOUTH, NORTH, BOTTOM, TOP = range(6)
FACE_NAMES = ["west", "east", "south", "north", "bottom", "top"]

# config.flux_calculator names (reference src/eilmer/globalconfig.d:293-345)
FLUX_CALCULATORS = {
    "ausmdv": 0, "hanel": 1, "ldfss0": 2, "ldfss2": 3, "ausm_plus_up": 4, "roe": 5,
    "adaptive_hanel_ausmdv": 6, "adaptive_hanel_ausm_plus_up": 7, "adaptive_ldfss0_ldfss2": 8,
    "efm": 9, "adaptive_efm_ausmdv": 10, "adaptive": 10, "hllc": 11, "hlle2": 12,
}
# config.gasdynamic_update_scheme names (reference src/eilmer/globalconfig.d:126-200)
UPDATE_SCHEMES = {
    "euler": 0, "pc": 1, "predictor-corrector": 1, "predictor_corrector": 1,
    "midpoint": 2, "classic-rk3": 3, "classic_rk3": 3, "tvd-rk3": 4, "tvd_rk3": 4,
    "denman-rk3": 5, "denman_rk3": 5, "classic-rk4": 6, "classic_rk4": 6,
}
THERMO_INTERPOLATORS = {"rhou": 0, "pt": 1, "rhop": 2, "rhot": 3}
N_STAGES = {0: 1, 1: 2, 2: 2, 3: 3, 4: 3, 5: 3, 6: 4}

GAS_IDEAL, GAS_THERMALLY_PERFECT = 0, 1

BC_WALL_WITH_SLIP = 0
BC_INFLOW_SUPERSONIC = 1
BC_OUTFLOW_SIMPLE_EXTRAPOLATE = 2
BC_OUTFLOW_SIMPLE_FLUX = 3
BC_EXCHANGE_FULL_FACE = 4
BC_OUTFLOW_FIXED_P = 5
BC_OUTFLOW_FIXED_PT = 6
BC_WALL_WITH_SLIP1 = 7
BC_GHOST_PROFILE = 8


class Species(C.Structure):
    _fields_ = [
        ("mol_mass", C.c_double),
        ("nsegments", C.c_int),
        ("T_break_points", C.c_double * (MAX_SEGMENTS + 1)),
        ("T_blend_ranges", C.c_double * MAX_SEGMENTS),
        ("coeffs", (C.c_double * 9) * MAX_SEGMENTS),
    ]


class Config(C.Structure):
    _fields_ = [
        ("dimensions", C.c_int),
        ("axisymmetric", C.c_int),
        ("gas_model", C.c_int),
        ("n_species", C.c_int),
        ("flux_calculator", C.c_int),
        ("interpolation_order", C.c_int),
        ("apply_limiter", C.c_int),
        ("extrema_clipping", C.c_int),
        ("interpolate_in_local_frame", C.c_int),
        ("apply_entropy_fix", C.c_int),
        ("update_scheme", C.c_int),
        ("max_invalid_cells", C.c_int),
        ("strict_fp", C.c_int),
        ("rank", C.c_int),
        ("device", C.c_int),
        ("reserved_i", C.c_int * 4), ("thermo_interpolator", C.c_int),
        ("epsilon_van_albada", C.c_double),
        ("M_inf", C.c_double),
        ("max_velocity", C.c_double),
        ("max_temp", C.c_double),
        ("min_temp", C.c_double),
        ("suggested_low_T_value", C.c_double),
        ("ignore_low_T_thermo_update_failure", C.c_int),
        ("strict_shock_detector", C.c_int),
        ("ideal_mol_mass", C.c_double),
        ("ideal_gamma", C.c_double),
        ("compression_tolerance", C.c_double),
        ("shear_tolerance", C.c_double),
        ("solver_variant", C.c_double),
        ("reserved_d", C.c_double * 3),
        ("species", Species * MAX_SPECIES),
    ]


DP = C.POINTER(C.c_double)
DPP = C.POINTER(DP)

EXCHANGE_FN = C.CFUNCTYPE(
    C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int),
    C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
    C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.c_void_p)

# name -> (restype, argtypes); every symbol include/eb200.h declares.
SIGNATURES = {
    "init": (C.c_int, [C.POINTER(Config)]),
    "finalize": (C.c_int, [C.c_int]),
    "last_error": (C.c_int, [C.c_char_p, C.c_int]),
    "block_create": (C.c_int, [C.c_int] * 6),
    "block_set_geometry": (C.c_int, [C.c_int, C.c_int, DP, DP, DP, DP, DP, DPP]),
    "block_set_bc": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, DP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "block_set_face_map": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_longlong]),
    "commit": (C.c_int, [C.c_int]),
    "set_exchange": (C.c_int, [C.c_int, EXCHANGE_FN, C.c_void_p]),
    "upload_flow": (C.c_int, [C.c_int, C.c_int, DPP, C.c_int]),
    "download_flow": (C.c_int, [C.c_int, C.c_int, DPP, C.c_int]),
    "download_conserved": (C.c_int, [C.c_int, C.c_int, DPP, C.c_int]),
    "download_conserved_async": (C.c_int, [C.c_int, C.c_int, DPP, C.c_int]),
    "wait_downloads": (C.c_int, [C.c_int]),
    "probe_cells": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), DP, C.c_int]),
    "compute_dt": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, DP]),
    "step": (C.c_int, [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "kernel_launches": (C.c_longlong, [C.c_int]),
    "flux_kernel_time": (C.c_int, [C.c_int, C.c_int, DP, C.POINTER(C.c_longlong)]),
    "run_steps": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "block_is_cartesian": (C.c_int, [C.c_int, C.c_int]),
    "cuda_stream": (C.c_void_p, [C.c_int]),
    "undo_step": (C.c_int, [C.c_int]),
    "p2p_export": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "p2p_import": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "describe": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "debug_face_flux": (C.c_int, [C.c_int, C.c_int, DP, DP, DP, DP, C.POINTER(C.c_int)]),
}

_HERE = os.path.dirname(os.path.abspath(__file__))
# EB200_LIBRARY: development aid to try another build of the same CUDA library
DEFAULT_LIBRARY = os.environ.get("EB200_LIBRARY", os.path.join(_HERE, "csrc", "libeb200.so"))


class Library:
    """A loaded implementation of the eb200 C ABI."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: the CUDA library is not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
                "There is no CPU fallback.")
        self.path = path
        self.prefix = prefix
        self.cdll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.cdll, prefix + name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def error(self):
        buf = C.create_string_buffer(1024)
        self.last_error(buf, 1024)
        return buf.value.decode(errors="replace")

    def check(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{self.prefix}{what} failed ({rc}): {self.error()}")
        return rc


_default = None


def load_library(path=None, prefix="eb200_"):
    """Load (once) and return the product library; tests may pass another path/prefix."""
    global _default
    if path is None:
        if _default is None:
            _default = Library(DEFAULT_LIBRARY, "eb200_")
        return _default
    return Library(path, prefix)
