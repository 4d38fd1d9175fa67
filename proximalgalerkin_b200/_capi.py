"""ctypes binding of liblvpp_b200.so (the C ABI declared in include/lvpp_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, a
:class:`LvppError` is raised.  ``load()`` only dlopens the library; nothing here needs a GPU until a
compute entry point is called.
"""
import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "liblvpp_b200.so"

OK = 0
E_INVALID, E_CUDA, E_CAPACITY, E_COMM, E_NOGPU = -1, -2, -3, -4, -5
OBSTACLE_ARRAY, OBSTACLE_PHI_SET = 0, 1
PC_JACOBI, PC_CHEBYSHEV, PC_MG = 0, 1, 2
LINESEARCH_NONE, LINESEARCH_BT = 0, 1
LINESEARCH_L2 = 2  # host loop only (linesearch.py); the library's own Newton loops know none and bt
FORM_GRADIENT, FORM_MULTIPHASE, FORM_SIGNORINI = 1, 2, 3

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class LvppError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"liblvpp_b200 error {code}: {message}")
        self.code = code


class ObstacleDesc(C.Structure):
    """struct lvpp_obstacle_desc"""

    _fields_ = [
        ("tdim", C.c_int32),
        ("nld", C.c_int32),
        ("num_nodes", C.c_int64),
        ("num_owned", C.c_int64),
        ("num_cells", C.c_int64),
        ("num_owned_cells", C.c_int64),
        ("node_coords", c_double_p),
        ("cell_nodes", c_int32_p),
        ("nq", C.c_int32),
        ("qweights", c_double_p),
        ("phi_tab", c_double_p),
        ("dphi_tab", c_double_p),
        ("qpoints", c_double_p),
        ("num_bc", C.c_int64),
        ("bc_nodes", c_int32_p),
        ("bc_values", c_double_p),
        ("obstacle_kind", C.c_int32),
        ("phi_obs_q", c_double_p),
        ("f", C.c_double),
        ("obstacle_period", C.c_double),
        ("obstacle_origin", C.c_double),
        ("num_neighbors", C.c_int32),
        ("neighbor_ranks", c_int32_p),
        ("send_ptr", c_int64_p),
        ("send_nodes", c_int32_p),
        ("recv_ptr", c_int64_p),
        ("recv_nodes", c_int32_p),
    ]


class NewtonOpts(C.Structure):
    """struct lvpp_newton_opts (defaults = PETSc SNES defaults)"""

    _fields_ = [
        ("snes_rtol", C.c_double),
        ("snes_atol", C.c_double),
        ("snes_stol", C.c_double),
        ("snes_divtol", C.c_double),
        ("snes_max_it", C.c_int32),
        ("ksp_rtol", C.c_double),
        ("ksp_atol", C.c_double),
        ("ksp_max_it", C.c_int32),
        ("pc_type", C.c_int32),
        ("pc_degree", C.c_int32),
        ("snes_linesearch", C.c_int32),
        ("ksp_restart", C.c_int32),
        ("psi_increase_max", C.c_double),
        ("psi_free_below", C.c_double),
    ]

    @classmethod
    def defaults(cls):
        return cls(1e-8, 1e-50, 1e-8, 1e4, 50, 1e-12, 1e-50, 100000, PC_JACOBI, 0, LINESEARCH_NONE, 0, 0.0, -1e300)


class IntegralDesc(C.Structure):
    """struct lvpp_integral_desc"""

    _fields_ = [
        ("num_entities", C.c_int64),
        ("nld", C.c_int32),
        ("nv", C.c_int32),
        ("dofs", c_int32_p),
        ("vertices", c_int32_p),
        ("to_nnz", c_int64_p),
        ("nq", C.c_int32),
        ("reserved0", C.c_int32),
        ("qweights", c_double_p),
        ("tab_a", c_double_p),
        ("dtab_a", c_double_p),
        ("tab_b", c_double_p),
    ]


class FormDesc(C.Structure):
    """struct lvpp_form_desc"""

    _fields_ = [
        ("form", C.c_int32),
        ("gdim", C.c_int32),
        ("num_dofs", C.c_int64),
        ("num_vertices", C.c_int64),
        ("vertex_coords", c_double_p),
        ("indptr", c_int64_p),
        ("indices", c_int32_p),
        ("num_bc", C.c_int64),
        ("bc_dofs", c_int64_p),
        ("bc_values", c_double_p),
        ("num_integrals", C.c_int32),
        ("num_params", C.c_int32),
        ("integrals", C.POINTER(IntegralDesc)),
        ("params", c_double_p),
        ("coef0", c_double_p),
        ("coef1", c_double_p),
        ("num_blocks", C.c_int64),
        ("block_ptr", c_int64_p),
        ("block_dofs", c_int32_p),
    ]


class Stats(C.Structure):
    """struct lvpp_stats"""

    _fields_ = [
        ("num_rows", C.c_int64),
        ("local_rows", C.c_int64),
        ("nnz", C.c_int64),
        ("scalar_nnz", C.c_int64),
        ("sell_slots", C.c_int64),
        ("krylov_iterations", C.c_int64),
        ("newton_steps", C.c_int64),
        ("residual_evals", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("device_bytes", C.c_int64),
        ("t_assembly_ms", C.c_double),
        ("t_krylov_ms", C.c_double),
        ("last_spmv_ms", C.c_double),
        ("spmv_sampled_ms", C.c_double),
        ("spmv_samples", C.c_int64),
        ("fine_op_launches", C.c_int64),
        ("vcycles", C.c_int64),
        ("mg_levels", C.c_int32),
        ("reserved0", C.c_int32),
        ("smooth_sampled_ms", C.c_double),
        ("smooth_samples", C.c_int64),
        ("packed_op_launches", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


H = C.c_void_p
VP = C.c_void_p  # device pointers travel as integers

# name -> (restype, argtypes); every symbol include/lvpp_b200.h declares
SIGNATURES = {
    "lvpp_last_error": (C.c_char_p, []),
    "lvpp_version": (C.c_int, []),
    "lvpp_device_count": (C.c_int, []),
    "lvpp_create": (C.c_int, [C.POINTER(ObstacleDesc), C.POINTER(H)]),
    "lvpp_destroy": (C.c_int, [H]),
    "lvpp_get_stats": (C.c_int, [H, C.POINTER(Stats)]),
    "lvpp_get_csr_pattern": (C.c_int, [H, c_int64_p, c_int32_p]),
    "lvpp_set_alpha": (C.c_int, [H, C.c_double]),
    "lvpp_set_forcing": (C.c_int, [H, C.c_double]),
    "lvpp_set_bc_values": (C.c_int, [H, C.c_int64, c_int32_p, c_double_p]),
    "lvpp_get_last_iterate_host": (C.c_int, [H, c_double_p]),
    "lvpp_set_previous": (C.c_int, [H, VP]),
    "lvpp_assemble_residual": (C.c_int, [H, VP, VP, c_double_p]),
    "lvpp_assemble_jacobian": (C.c_int, [H, VP]),
    "lvpp_get_jacobian_values": (C.c_int, [H, VP]),
    "lvpp_spmv": (C.c_int, [H, VP, VP]),
    "lvpp_linear_solve": (C.c_int, [H, VP, VP, C.POINTER(NewtonOpts), c_int32_p, c_int32_p, c_double_p]),
    "lvpp_newton_solve": (C.c_int, [H, VP, C.POINTER(NewtonOpts), c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "lvpp_newton_begin": (C.c_int, [H, VP, c_double_p]),
    "lvpp_newton_begin_same_iterate": (C.c_int, [H, VP, c_double_p]),
    "lvpp_newton_step": (C.c_int, [H, VP, C.POINTER(NewtonOpts), c_double_p, c_int32_p, c_int32_p]),
    "lvpp_observables": (C.c_int, [H, VP, c_double_p]),
    "lvpp_newton_solve_host": (C.c_int, [H, c_double_p, C.POINTER(NewtonOpts), c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "lvpp_set_previous_host": (C.c_int, [H, c_double_p]),
    "lvpp_timer_start": (C.c_int, [H]),
    "lvpp_timer_stop": (C.c_int, [H, c_double_p]),
    "lvpp_time_spmv": (C.c_int, [H, VP, VP, C.c_int32, C.c_int32, c_double_p]),
    "lvpp_time_assembly": (C.c_int, [H, VP, VP, C.c_int32, c_double_p, c_double_p, c_double_p]),
    "lvpp_comm_unique_id": (C.c_int, [c_uint8_p]),
    "lvpp_comm_init": (C.c_int, [H, c_uint8_p, C.c_int32, C.c_int32]),
    "lvpp_halo_forward": (C.c_int, [H, VP]),
    "lvpp_form_create": (C.c_int, [C.POINTER(FormDesc), C.POINTER(H)]),
    "lvpp_form_destroy": (C.c_int, [H]),
    "lvpp_form_set_param": (C.c_int, [H, C.c_int32, C.c_double]),
    "lvpp_form_set_aux": (C.c_int, [H, C.c_int32, VP]),
    "lvpp_form_set_bc_values": (C.c_int, [H, c_double_p]),
    "lvpp_form_assemble_residual": (C.c_int, [H, VP, VP, c_double_p]),
    "lvpp_form_get_jacobian_values": (C.c_int, [H, VP]),
    "lvpp_form_spmv": (C.c_int, [H, VP, VP]),
    "lvpp_form_linear_solve": (C.c_int, [H, VP, VP, C.POINTER(NewtonOpts), c_int32_p, c_int32_p, c_double_p]),
    "lvpp_form_newton_solve": (C.c_int, [H, VP, C.POINTER(NewtonOpts), c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "lvpp_form_increment_sq": (C.c_int, [H, VP, VP, c_double_p]),
    "lvpp_form_get_stats": (C.c_int, [H, C.POINTER(Stats)]),
    "lvpp_form_time_kernels": (C.c_int, [H, VP, C.c_int32, c_double_p, c_double_p]),
}

_lib = None


def load():
    """dlopen the in-tree library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("LVPP_B200_LIB", LIB_PATH))
    if not path.exists():
        raise LvppError(
            E_INVALID,
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for this path)",
        )
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != OK:
        msg = load().lvpp_last_error()
        raise LvppError(code, msg.decode() if msg else "unknown")


def as_ptr(arr, ctype):
    """ctypes pointer to a C-contiguous numpy array (caller keeps the array alive)."""
    return arr.ctypes.data_as(C.POINTER(ctype))
