"""ctypes binding of the C ABI declared in ``include/lintsampler_b200.h``.

The shared library is built in-tree by ``lintsampler_b200._build`` (nvcc,
sm_100a) and loaded from ``lintsampler_b200/_lib``.  There is no CPU fallback:
if the library is missing, or no CUDA device is present when device work is
requested, the product raises.
"""
import ctypes as C
import os

LINT_MAX_DIM = 6
LINT_GMM_MAX_COMP = 16
LINT_F32, LINT_F64 = 0, 1
LINT_CDF_PARALLEL, LINT_CDF_SEQUENTIAL = 0, 1
LINT_FLAG_NEGATIVE, LINT_FLAG_NONFINITE = 1, 2

# LINTSAMPLER_B200_LIB: load another build of the same library (kernel experiments)
LIB_PATH = os.environ.get("LINTSAMPLER_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "_lib", "liblintsampler_b200.so")


class LintGrid(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dtype", C.c_int32),
        ("ncells", C.c_int64 * LINT_MAX_DIM),
        ("vertex_densities_dev", C.c_void_p),
        ("edges_dev", C.c_void_p * LINT_MAX_DIM),
        ("cdf_dev", C.c_void_p),
        ("guide_dev", C.c_void_p),
        ("guide_bits", C.c_int32), ("reserved", C.c_int32),
        ("cell_records_dev", C.c_void_p),
    ]


class LintCells(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dtype", C.c_int32),
        ("ncells", C.c_int64),
        ("mins_dev", C.c_void_p), ("widths_dev", C.c_void_p),
        ("corners_dev", C.c_void_p), ("cdf_dev", C.c_void_p),
        ("guide_dev", C.c_void_p),
        ("guide_bits", C.c_int32), ("reserved", C.c_int32),
        ("cell_records_dev", C.c_void_p),
    ]


class LintSobol(C.Structure):
    _fields_ = [
        ("sv", (C.c_uint32 * 32) * (LINT_MAX_DIM + 1)),
        ("shift", C.c_uint32 * (LINT_MAX_DIM + 1)),
        ("first_index", C.c_uint64),
        ("sorted_bits", C.c_uint32),
        ("sorted_ainv", C.c_uint32 * 32),
    ]


class LintHalos(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ncomp", C.c_int32),
        ("rho0", C.POINTER(C.c_double)),
        ("centre", C.POINTER(C.c_double)),
        ("r_s", C.POINTER(C.c_double)),
        ("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
        ("eps", C.c_double), ("q_cut", C.c_double),
    ]


class LintGmm(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ncomp", C.c_int32),
        ("log_norm", C.POINTER(C.c_double)),
        ("mean", C.POINTER(C.c_double)),
        ("prec_chol", C.POINTER(C.c_double)),
    ]


class LintTree(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("reserved", C.c_int32),
        ("mins", C.c_double * LINT_MAX_DIM), ("maxs", C.c_double * LINT_MAX_DIM),
        ("capacity", C.c_int64),
        ("origin_dev", C.c_void_p), ("level_dev", C.c_void_p), ("parent_dev", C.c_void_p),
        ("corners_dev", C.c_void_p), ("mass_raw_dev", C.c_void_p), ("est_dev", C.c_void_p),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_U64 = C.c_uint64

# name -> (restype, argtypes); every symbol include/lintsampler_b200.h declares
PROTOTYPES = {
    "lint_abi_version": (C.c_int, []),
    "lint_last_error": (C.c_char_p, []),
    "lint_set_offset_counter": (C.c_int, [_P]),
    "lint_advance_counter": (C.c_int, [_P, _U64, _P]),
    "lint_grid_eval_gmm": (C.c_int, [C.POINTER(LintGrid), C.POINTER(LintGmm), C.c_double, _P]),
    "lint_grid_eval_halos": (C.c_int, [C.POINTER(LintGrid), C.POINTER(LintHalos), C.c_double, _P]),
    "lint_tree_est_stride": (C.c_int, []),
    "lint_tree_open": (C.c_int, [C.POINTER(LintTree), C.POINTER(LintGmm), C.POINTER(LintHalos), _P, _I64, _I64, _P,
                                 _P, _P]),
    "lint_tree_refine_worst": (C.c_int, [C.POINTER(LintTree), C.POINTER(LintGmm), C.POINTER(LintHalos), _P, _P, _P,
                                         _P, _I64, C.c_double, _I64, _P, _P, _P]),
    "lint_tree_max_refine_leaves": (C.c_int, []),
    "lint_points_eval": (C.c_int, [C.c_int, C.POINTER(LintGmm), C.POINTER(LintHalos), _P, _I64, _P, _P]),
    "lint_grid_masses": (C.c_int, [C.POINTER(LintGrid), _P, _P, _P]),
    "lint_grid_records": (C.c_int, [C.POINTER(LintGrid), _P, _P]),
    "lint_cells_record_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "lint_cells_records": (C.c_int, [C.POINTER(LintCells), _P, _P]),
    "lint_cells_check": (C.c_int, [C.POINTER(LintCells), _P, _P]),
    "lint_cdf_scratch_bytes": (C.c_size_t, [_I64]),
    "lint_sum_numpy_f64": (C.c_int, [_P, _I64, _P, _P, C.c_size_t, _P]),
    "lint_cdf_build": (C.c_int, [_P, _I64, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "lint_guide_build": (C.c_int, [_P, _I64, C.c_int, _P, _P]),
    "lint_grid_build_scratch_bytes": (C.c_size_t, [C.POINTER(LintGrid)]),
    "lint_grid_group_cells": (_I64, [C.POINTER(LintGrid)]),
    "lint_grid_build": (C.c_int, [C.POINTER(LintGrid), C.c_double, C.c_double, _P, _P, _P, C.c_int, _P, _P, _P,
                                  _P, C.c_size_t, _P]),
    "lint_grid_sample_u": (C.c_int, [C.POINTER(LintGrid), _P, _I64, _P, _P, _P]),
    "lint_cells_sample_u": (C.c_int, [C.POINTER(LintCells), _P, _I64, _P, _P, _P]),
    "lint_incell_sample_u": (C.c_int, [C.c_int, _I64, _P, _P, _P, _P, _P, _P]),
    "lint_grid_sample_philox": (C.c_int, [C.POINTER(LintGrid), _I64, _U64, _U64, C.c_double, C.c_double,
                                          _P, _P, _P]),
    "lint_cells_sample_philox": (C.c_int, [C.POINTER(LintCells), _I64, _U64, _U64, C.c_double, C.c_double,
                                           _P, _P, _P]),
    "lint_philox_uniforms": (C.c_int, [C.c_int, C.c_int, _I64, _U64, _U64, _P, _P]),
    "lint_grid_sample_sobol": (C.c_int, [C.POINTER(LintGrid), _I64, C.POINTER(LintSobol), _P, _P, _P]),
    "lint_cells_sample_sobol": (C.c_int, [C.POINTER(LintCells), _I64, C.POINTER(LintSobol), _P, _P, _P]),
    "lint_sobol_uniforms": (C.c_int, [C.c_int, _I64, C.POINTER(LintSobol), _P, _P]),
    "lint_multi_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "lint_multi_grid_sample_u": (C.c_int, [C.POINTER(LintGrid), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           _P, _I64, _P, C.c_size_t, _P, _P, _P]),
    "lint_multi_grid_sample_philox": (C.c_int, [C.POINTER(LintGrid), C.c_int, C.POINTER(C.c_double),
                                                C.POINTER(C.c_double), _I64, _U64, _U64, _P, C.c_size_t, _P, _P, _P]),
    "lint_multi_grid_sample_sobol": (C.c_int, [C.POINTER(LintGrid), C.c_int, C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), _I64, C.POINTER(LintSobol), _P, C.c_size_t, _P,
                                               _P, _P]),
    "lint_permute_rows": (C.c_int, [_P, _P, _I64, C.c_int32, C.c_int32, _U64, _P]),
    "lint_cellmajor_scratch_bytes": (C.c_size_t, [_I64, _I64]),
    "lint_cellmajor_timing": (C.c_int, [C.c_int]),
    "lint_cellmajor_last_times": (C.c_int, [C.POINTER(C.c_double)]),
    "lint_grid_sample_cellmajor": (C.c_int, [C.POINTER(LintGrid), _I64, _U64, _U64, C.c_double, C.c_double,
                                             _P, _P, C.c_size_t, _P, _P, _P]),
    "lint_cells_sample_cellmajor": (C.c_int, [C.POINTER(LintCells), _I64, _U64, _U64, C.c_double,
                                              C.c_double, _P, C.c_size_t, _P, _P, _P]),
}

_lib = None


class LintError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def load():
    """Load (once) and return the ctypes handle of the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"lintsampler_b200: CUDA library not built ({LIB_PATH}). Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python -m lintsampler_b200._build`. There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.lint_abi_version() != 2:
        raise ImportError("lintsampler_b200: ABI version mismatch; rebuild the library")
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-status entry point; raise LintError on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.lint_last_error().decode("utf-8", "replace")
        raise LintError(f"{name} failed (status {rc}): {msg}")
