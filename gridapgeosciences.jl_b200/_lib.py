"""ctypes binding of libgeosgpu.so (the C ABI declared in include/geosgpu.h).

The library is the product; this module only declares its symbols.  There is no CPU path: if the
shared library is missing the import fails, and if no B200 is visible `gg_init` fails.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB_PATH") or os.path.join(_HERE, "libgeosgpu.so")   # GG_LIB_PATH: development builds

GG_OK = 0
ERRORS = {-1: "GG_ERR_INVALID", -2: "GG_ERR_CUDA", -3: "GG_ERR_UNSUPPORTED", -4: "GG_ERR_NOMEM", -5: "GG_ERR_NOCONV", -6: "GG_ERR_COMM"}

SPACE_CG, SPACE_DG, SPACE_RT, SPACE_CGV = 0, 1, 2, 3
CONSTRAINT_NONE, CONSTRAINT_ZEROMEAN, CONSTRAINT_NOFLUX = 0, 1, 2
FORM_LAPLACE_BELTRAMI = 1
FORM_LAPLACE_BELTRAMI_EXTRINSIC = 2
FORM_HELMHOLTZ = 3
FORM_MASS_SCALAR = 4
FORM_MASS_RT = 5
FORM_DIV = 6
FORM_SOURCE = 7
FORM_MASS_VECTOR, FORM_SOURCE_VECTOR = 8, 9
FORM_SHELL_BOUNDARY_FLUX = 10
FORM_VECTOR_LAPLACIAN, FORM_VECTOR_DIV = 11, 12
PROBLEM_WAVE = 1
PROBLEM_SHALLOW_WATER = 2
PROBLEM_ADVECTION = 3
PROBLEM_THERMAL_SHALLOW_WATER = 4
PROBLEM_BOUSSINESQ = 5
Q0_ALIAS, Q0_START_OF_STEP = 0, 1


class GGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class FormArgs(C.Structure):
    _fields_ = [("test", C.c_void_p), ("trial", C.c_void_p), ("degree", C.c_int32), ("u", C.c_void_p),
                ("qdata", C.c_void_p), ("dirichlet", C.c_double), ("scale", C.c_double)]


class RKArgs(C.Structure):
    _fields_ = [("problem", C.c_int32), ("order", C.c_int32), ("degree", C.c_int32), ("n_stages", C.c_int32),
                ("A", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)), ("c", C.POINTER(C.c_double)),
                ("dt", C.c_double), ("rtol", C.c_double),
                ("gravity", C.c_double), ("tau", C.c_double), ("q0_mode", C.c_int32),
                ("coriolis", C.POINTER(C.c_double)), ("topography", C.POINTER(C.c_double)),
                ("velocity_cells", C.POINTER(C.c_double)), ("velocity_facets", C.POINTER(C.c_double)),
                ("sound_speed", C.c_double), ("buoyancy_frequency", C.c_double)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I64 = C.c_int64
_PI64 = C.POINTER(C.c_int64)
_PI32 = C.POINTER(C.c_int32)
_PI8 = C.POINTER(C.c_int8)
_PD = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/geosgpu.h declares
SIGNATURES = {
    "gg_init": (C.c_int, [C.c_int, _PP]),
    "gg_finalize": (None, [_P]),
    "gg_last_error": (C.c_char_p, [_P]),
    "gg_stream": (C.c_void_p, [_P]),
    "gg_synchronize": (C.c_int, [_P]),
    "gg_launch_count": (_I64, [_P]),
    "gg_set_geometry_classes": (C.c_int, [_P, C.c_int]),
    "gg_set_async": (C.c_int, [_P, C.c_int]),
    "gg_profile_enable": (C.c_int, [_P, C.c_int]),
    "gg_profile_query": (C.c_int, [_P, C.c_char_p, _PI64, _PD]),
    "gg_measure_fp64_peak": (C.c_int, [_P, _PD]),
    "gg_mesh_cubed_sphere": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _PP]),
    "gg_mesh_from_arrays": (C.c_int, [_P, C.c_int, _I64, _I64, _PI32, _PD, _PI32, C.c_double, C.c_double, _PP]),
    "gg_mesh_info": (C.c_int, [_P, _PI64, _PI64, _PI64]),
    "gg_mesh_get_cell_nodes": (C.c_int, [_P, _PI32]),
    "gg_mesh_get_cell_to_chart": (C.c_int, [_P, _PI32]),
    "gg_mesh_get_chart_coords": (C.c_int, [_P, _PD]),
    "gg_mesh_destroy": (None, [_P]),
    "gg_partition_range": (C.c_int, [_I64, C.c_int, C.c_int, _PI64, _PI64]),
    "gg_mesh_restrict": (C.c_int, [_P, _I64, _PI64, _PP]),
    "gg_mesh_ghost_cells": (C.c_int, [_P, _I64, _I64, _PI64, _PI64]),
    "gg_plan_create": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, _PP]),
    "gg_plan_create_shell": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _PP]),
    "gg_plan_create_strips": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, _PP]),
    "gg_plan_owned_cells": (C.c_int, [_P, _PI64]),
    "gg_plan_destroy": (None, [_P]),
    "gg_plan_info": (C.c_int, [_P, _PI64, _PI64, _PI64, _PI64]),
    "gg_plan_local_cells": (C.c_int, [_P, _PI64]),
    "gg_plan_space_info": (C.c_int, [_P, C.c_int, C.c_int, _PI64, _PI64, _PI64, _PI32, _PI64]),
    "gg_plan_space_maps": (C.c_int, [_P, C.c_int, C.c_int, _PI64, _PI32]),
    "gg_plan_space_neighbor": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _PI32, _PI64, _PI64, _PI32, _PI32, _PI32]),
    "gg_comm_create": (C.c_int, [_P, C.c_int, C.c_int, _I64, _PP]),
    "gg_comm_handle": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "gg_comm_connect": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "gg_comm_error": (C.c_int, [_P, _PI64]),
    "gg_comm_destroy": (None, [_P]),
    "gg_mesh_partitioned": (C.c_int, [_P, _P, _P, _PP]),
    "gg_space_global_dofs": (C.c_int, [_P, _PI64, _PI32]),
    "gg_vec_consistent": (C.c_int, [_P, _P]),
    "gg_space_builtin": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _PP]),
    "gg_space_from_arrays": (C.c_int, [_P, C.c_int, C.c_int, _I64, _PI32, _PI8, _PP]),
    "gg_space_info": (C.c_int, [_P, _PI64, _PI32, _PI64]),
    "gg_space_get_cell_dofs": (C.c_int, [_P, _PI32]),
    "gg_space_get_cell_signs": (C.c_int, [_P, _PI8]),
    "gg_space_change_of_basis": (C.c_int, [_P, _PD]),
    "gg_space_destroy": (None, [_P]),
    "gg_vec_create": (C.c_int, [_P, _I64, _PP]),
    "gg_vec_set": (C.c_int, [_P, _PD]),
    "gg_vec_get": (C.c_int, [_P, _PD]),
    "gg_vec_fill": (C.c_int, [_P, C.c_double]),
    "gg_vec_size": (_I64, [_P]),
    "gg_vec_device_ptr": (C.c_void_p, [_P]),
    "gg_vec_destroy": (None, [_P]),
    "gg_pattern": (C.c_int, [_P, _P, _PP]),
    "gg_pattern_skeleton": (C.c_int, [_P, _PP]),
    "gg_assemble_advection": (C.c_int, [_P, C.c_int32, _PD, _PD, _P]),
    "gg_csr_info": (C.c_int, [_P, _PI64, _PI64, _PI64]),
    "gg_csr_export": (C.c_int, [_P, _PI64, _PI64, _PD, C.c_int]),
    "gg_csr_export_csc": (C.c_int, [_P, _PI64, _PI64, _PD, C.c_int]),
    "gg_csr_device_values": (C.c_void_p, [_P]),
    "gg_csr_spmv": (C.c_int, [_P, C.c_double, _P, C.c_double, _P]),
    "gg_csr_destroy": (None, [_P]),
    "gg_assemble_matrix": (C.c_int, [C.c_int, C.POINTER(FormArgs), _P, C.c_int]),
    "gg_assemble_vector": (C.c_int, [C.c_int, C.POINTER(FormArgs), _P, C.c_int]),
    "gg_assemble_matrix_and_vector": (C.c_int, [C.c_int, C.POINTER(FormArgs), _P, _P]),
    "gg_element_matrices": (C.c_int, [C.c_int, C.POINTER(FormArgs), _I64, _I64, _PD]),
    "gg_element_vectors": (C.c_int, [C.c_int, C.POINTER(FormArgs), _I64, _I64, _PD]),
    "gg_quadrature_points": (C.c_int, [_P, C.c_int32, _PD, _PD]),
    "gg_skeleton_points": (C.c_int, [_P, C.c_int32, _PI32, _PD]),
    "gg_integrate": (C.c_int, [_P, C.c_int32, _P, _PD]),
    "gg_eval_geometry": (C.c_int, [_P, C.c_int, C.c_double, _I64, _PD, _PD, _PD, _PD, _PD, _PD]),
    "gg_solver_pcg_create": (C.c_int, [_P, C.c_double, C.c_int, _PP]),
    "gg_solver_update": (C.c_int, [_P]),
    "gg_solver_solve": (C.c_int, [_P, _P, _P, _PI32, _PD]),
    "gg_solver_destroy": (None, [_P]),
    "gg_space_interp_points": (C.c_int, [_P, _PI32, _PD, _PD]),
    "gg_space_interpolate": (C.c_int, [_P, _PD, _P]),
    "gg_norm_sq": (C.c_int, [_P, _P, C.c_int32, _PD]),
    "gg_h1_seminorm_sq": (C.c_int, [_P, _P, C.c_int32, _PD]),
    "gg_surface_laplacian": (C.c_int, [_P, C.c_int32, _PD, _PD, _P]),
    "gg_l2_error_sq": (C.c_int, [_P, _P, C.c_double, _P, C.c_int32, _PD]),
    "gg_surface_gradient": (C.c_int, [_P, C.c_int32, _PD, _P]),
    "gg_surface_divergence": (C.c_int, [_P, C.c_int32, _PD, _PD, _P]),
    "gg_chart_mean": (C.c_int, [_P, _P, C.c_double, _PD]),
    "gg_rk_create": (C.c_int, [_P, C.POINTER(RKArgs), _PP]),
    "gg_rk_info": (C.c_int, [_P, _PI64, _PI64]),
    "gg_rk_set_state": (C.c_int, [_P, _PD]),
    "gg_rk_get_state": (C.c_int, [_P, _PD]),
    "gg_rk_state_device_ptr": (C.c_void_p, [_P]),
    "gg_rk_step": (C.c_int, [_P, C.c_int32]),
    "gg_rk_residual": (C.c_int, [_P, _PD, _PD]),
    "gg_rk_stats": (C.c_int, [_P, _PI64, _PI64]),
    "gg_rk_solver_phase_ns": (C.c_int, [_P, C.c_int, _PD]),
    "gg_rk_diag_info": (C.c_int, [_P, _PI64, _PI64, _PI64]),
    "gg_rk_get_diagnostics": (C.c_int, [_P, _PD]),
    "gg_rk_diagnostics": (C.c_int, [_P, _PD, _PD]),
    "gg_rk_swe_residual": (C.c_int, [_P, _PD, _PD, _PD, _PD]),
    "gg_rk_destroy": (None, [_P]),
}

_lib = None


def load():
    """dlopen libgeosgpu.so and bind every declared symbol (raises if the library or a symbol is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                          "(there is no CPU fallback for this package)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != GG_OK:
        msg = load().gg_last_error(None)
        raise GGError(code, msg.decode() if msg else "")
