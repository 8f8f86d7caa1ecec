"""ctypes binding of libehdg.so (include/ehdg.h).  No CPU fallback: a missing library or a
missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("EHDG_LIB_NAME", "libehdg.so"))

MAX_ENR = 2


class LayerExpr(C.Structure):
    _fields_ = [("c", C.c_double * 4), ("k", C.c_double * 2)]


class Problem(C.Structure):
    _fields_ = [("order", C.c_int32), ("bonus_int", C.c_int32), ("epsilon", C.c_double),
                ("beta", C.c_double * 2), ("n_enrich", C.c_int32), ("enrich_axis", C.c_int32 * MAX_ENR),
                ("enrich_k", C.c_double * MAX_ENR), ("coeff", LayerExpr), ("exact", LayerExpr)]


class Mesh(C.Structure):
    _fields_ = [("nv", C.c_int32), ("ne", C.c_int32), ("nf", C.c_int32), ("verts", C.c_void_p),
                ("tris", C.c_void_p), ("el2facet", C.c_void_p), ("facet2el", C.c_void_p)]


class Sizes(C.Structure):
    _fields_ = [("n_p", C.c_int32), ("nfl", C.c_int32), ("nb", C.c_int32), ("nvmax", C.c_int32),
                ("nfmax", C.c_int32), ("nf_poly", C.c_int32), ("nq", C.c_int32), ("nqf", C.c_int32),
                ("nq0", C.c_int32), ("nqf0", C.c_int32), ("nq_err", C.c_int32),
                ("flops_poly_elem", C.c_int64), ("flops_quad_elem", C.c_int64),
                ("flops_poly_alpha", C.c_int64), ("flops_poly_assemble", C.c_int64)]


class PlanSizes(C.Structure):
    _fields_ = [("n_poly", C.c_int64), ("n_gen", C.c_int64), ("S_doubles", C.c_int64), ("g_doubles", C.c_int64),
                ("W_doubles", C.c_int64), ("z_doubles", C.c_int64)]


class CsrSizes(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("nnz", C.c_int64)]


class GmresInfo(C.Structure):
    _fields_ = [("restart", C.c_int32), ("max_iters", C.c_int32), ("rtol", C.c_double), ("iters", C.c_int32),
                ("converged", C.c_int32), ("rel_residual", C.c_double), ("total_seconds", C.c_double),
                ("spmv_calls", C.c_int64), ("dropped_dofs", C.c_int32),
                ("use_strip_block", C.c_int32), ("strip_dofs", C.c_int32), ("strip_blocks", C.c_int32),
                ("jacobi_sweeps", C.c_int32), ("use_line_block", C.c_int32), ("line_target_dofs", C.c_int32),
                ("line_chains", C.c_int32), ("line_blocks", C.c_int32), ("line_max_block", C.c_int32),
                ("setup_seconds", C.c_double), ("precond_bytes", C.c_double), ("rel_residual_unscaled", C.c_double), ("drop_tol", C.c_double),
                ("line_max_blocks", C.c_int32), ("reserved2", C.c_int32)]


# every symbol include/ehdg.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ehdg_version": (C.c_char_p, []),
    "ehdg_last_error": (C.c_char_p, []),
    "ehdg_create": (C.c_int, [C.POINTER(Problem), C.POINTER(_P)]),
    "ehdg_destroy": (C.c_int, [_P]),
    "ehdg_get_sizes": (C.c_int, [_P, C.POINTER(Sizes)]),
    "ehdg_set_expression": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
    "ehdg_set_drop_tol": (C.c_int, [_P, C.c_double]),
    "ehdg_kernel_timing": (C.c_int, [_P, C.c_int]),
    "ehdg_kernel_times_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ehdg_mark": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, _P, _P]),
    "ehdg_plan_create": (C.c_int, [_P, C.POINTER(Mesh), _P, C.c_int, C.POINTER(_P), _P]),
    "ehdg_select_elements": (C.c_int, [_P, C.POINTER(Mesh), C.c_int32, C.c_int32, _P, _P, _P]),
    "ehdg_plan_create_select": (C.c_int, [_P, C.POINTER(Mesh), _P, C.c_int, _P, C.POINTER(_P), _P]),
    "ehdg_plan_destroy": (C.c_int, [_P]),
    "ehdg_plan_get_sizes": (C.c_int, [_P, C.POINTER(PlanSizes)]),
    "ehdg_plan_eslot": (_P, [_P]),
    "ehdg_prune": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, C.c_double, _P, _P, _P, _P]),
    "ehdg_alpha": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, _P, _P]),
    "ehdg_assemble_condense": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ehdg_edg_set_boundary_data": (C.c_int, [_P, C.c_int]),
    "ehdg_assemble_condense_csr": (C.c_int, [_P, _P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ehdg_csr_symbolic": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, C.POINTER(_P), _P]),
    "ehdg_csr_symbolic_part": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, C.c_int, C.c_int, C.POINTER(_P), _P]),
    "ehdg_csr_get_partition": (C.c_int, [_P, _P, _P]),
    "ehdg_csr_get_extent": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ehdg_csr_set_lines": (C.c_int, [_P, _P, _P]),
    "ehdg_csr_auto_lines": (C.c_int, [_P, _P, C.POINTER(Mesh), C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "ehdg_csr_lines": (_P, [_P]),
    "ehdg_csr_destroy": (C.c_int, [_P]),
    "ehdg_csr_get_sizes": (C.c_int, [_P, C.POINTER(CsrSizes)]),
    "ehdg_csr_rowptr": (_P, [_P]),
    "ehdg_csr_colind": (_P, [_P]),
    "ehdg_csr_row_start": (_P, [_P]),
    "ehdg_csr_dofmask": (_P, [_P]),
    "ehdg_csr_numeric": (C.c_int, [_P, _P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P]),
    "ehdg_dirichlet_project": (C.c_int, [_P, C.POINTER(Mesh), _P, _P]),
    "ehdg_csr_numeric_bc": (C.c_int, [_P, _P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P, _P]),
    "ehdg_backsolve_bc": (C.c_int, [_P, _P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P, _P]),
    "ehdg_gmres": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(GmresInfo), _P]),
    "ehdg_bicgstab": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(GmresInfo), _P]),
    "ehdg_dqgmres": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(GmresInfo), _P]),
    "ehdg_edg_prune": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, C.c_double, _P, _P, _P]),
    "ehdg_edg_alpha": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, _P, _P]),
    "ehdg_edg_alpha_carry": (C.c_int, [_P, C.c_int32, _P, _P]),
    "ehdg_edg_csr_symbolic": (C.c_int, [_P, C.POINTER(Mesh), _P, C.POINTER(_P), _P]),
    "ehdg_edg_assemble": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, _P, _P, _P, _P, _P]),
    "ehdg_edg_unpack": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, _P, _P, _P]),
    "ehdg_edg_l2_error": (C.c_int, [_P, C.POINTER(Mesh), _P, _P, _P, _P]),
    "ehdg_comm_unique_id": (C.c_int, [C.c_char_p]),
    "ehdg_comm_init": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int]),
    "ehdg_comm_set_partition": (C.c_int, [_P, _P, _P, _P]),
    "ehdg_comm_destroy": (C.c_int, [_P]),
    "ehdg_krylov_workspace_bytes": (C.c_int, [_P, _P, C.POINTER(GmresInfo), C.c_int, C.POINTER(C.c_int64)]),
    "ehdg_set_workspace": (C.c_int, [_P, _P, C.c_int64]),
    "ehdg_spmv": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ehdg_backsolve": (C.c_int, [_P, _P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P]),
    "ehdg_l2_error": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, _P, _P]),
    "ehdg_l2_error_select": (C.c_int, [_P, _P, C.POINTER(Mesh), _P, _P, _P, _P, _P]),
}

_lib = None


class EhdgError(RuntimeError):
    pass


def load():
    """Load libehdg.so and bind every symbol of include/ehdg.h.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EhdgError(f"{LIB_PATH} not built: run `python -m convectiondiffusionequation_b200.build` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ehdg_last_error().decode()
        raise EhdgError(f"{what} failed ({rc}): {msg}")
