"""ctypes binding of libmetafem_b200.so -- the same C ABI the Julia side binds with ccall.

There is no CPU fallback: if the shared library is missing or no CUDA device is present, the
product path raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmetafem_b200.so")

MFB_OK, MFB_NOT_CONVERGED = 0, 1
MFB_IDRS, MFB_BICGSTABL_GS, MFB_BICGSTABL, MFB_GMRES, MFB_CGS, MFB_CGS2, MFB_TFQMR, MFB_LSQR, MFB_IDRS_ORIGINAL = range(9)
PR_JACOBI, PR_JACOBI_COLUMN, PR_IDENTITY = 0, 1, 2
PL_IDENTITY, PL_JACOBI, PL_JACOBI_ROW, PL_ILU = 0, 1, 2, 3
VEC_X, VEC_DX, VEC_X_STAR, VEC_RESIDUE = 0, 1, 2, 3
MAT_K_LINEAR, MAT_K_TOTAL = 0, 1
NUMBERING_SORTED, NUMBERING_REFERENCE = 0, 1


class MfbError(RuntimeError):
    pass


class BlockDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("bg_ID", C.c_int32), ("linear_kernel", C.c_char_p),
                ("nonlinear_kernel", C.c_char_p), ("n_cp_vars", C.c_int32), ("cp_var_names", C.POINTER(C.c_char_p)),
                ("n_globals", C.c_int32), ("global_names", C.POINTER(C.c_char_p)),
                ("threads_per_block", C.c_int32), ("smem_bytes", C.c_int32), ("has_nonlinear_K", C.c_int32),
                ("eval_kernel", C.c_char_p), ("n_qp_in", C.c_int32), ("qp_in_names", C.POINTER(C.c_char_p)),
                ("n_qp_out", C.c_int32), ("qp_out_names", C.POINTER(C.c_char_p))]


class J2Params(C.Structure):
    _fields_ = [("lam", C.c_double), ("mu", C.c_double), ("Eb", C.c_double), ("Ep", C.c_double), ("f_res", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("passes", C.c_int32), ("iterations", C.c_int32), ("spmv_count", C.c_int32),
                ("converged", C.c_int32), ("residual", C.c_double), ("initial_residual", C.c_double)]


_P = C.c_void_p
_SIGS = {
    "mfb_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "mfb_destroy": (C.c_int, [_P]),
    "mfb_last_error": (C.c_char_p, [_P]),
    "mfb_set_stream": (C.c_int, [_P, _P]),
    "mfb_launch_count": (C.c_int64, [_P]),
    "mfb_synchronize": (C.c_int, [_P]),
    "mfb_measure_fp64_peak": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mfb_profile_enable": (C.c_int, [_P, C.c_int]),
    "mfb_profile_get": (C.c_int, [_P, _P, _P]),
    "mfb_mesh_set": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mfb_mesh_build_second_order": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int, C.c_int64, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P,
                                              C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mfb_mesh_build_get": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "mfb_mesh_build_device_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mfb_total_mesh_build": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int64, _P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]),
    "mfb_total_mesh_get": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mfb_facets_set": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P]),
    "mfb_boundary_group_set": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "mfb_pattern_build": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mfb_pattern_get": (C.c_int, [_P, _P, _P, _P, _P]),
    "mfb_sparse_ids_get": (C.c_int, [_P, C.c_int, _P]),
    "mfb_field_set": (C.c_int, [_P, C.c_char_p, _P]),
    "mfb_global_set": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "mfb_vector_set": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "mfb_vector_get": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "mfb_matrix_get": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "mfb_kernel_compile": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(BlockDesc)]),
    "mfb_compile_log": (C.c_char_p, [_P]),
    "mfb_kernel_check": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "mfb_assemble_linear": (C.c_int, [_P, _P, C.c_int]),
    "mfb_assemble_nonlinear": (C.c_int, [_P, _P, C.c_int, C.c_double, C.c_double]),
    "mfb_qp_array": (C.c_int, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "mfb_qp_set": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "mfb_qp_get": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "mfb_eval_qp_args": (C.c_int, [_P, C.c_double, C.c_double]),
    "mfb_j2_init": (C.c_int, [_P, C.c_char_p, C.c_double, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
    "mfb_j2_iterate_stress": (C.c_int, [_P, C.c_char_p, C.POINTER(J2Params), C.POINTER(C.c_int64)]),
    "mfb_j2_update_states": (C.c_int, [_P, C.c_char_p]),
    "mfb_j2_yield_count": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "mfb_spmv": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64]),
    "mfb_spmv_variant_bench": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mfb_krylov_solve": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, _P,
                                   C.POINTER(SolveInfo)]),
    "mfb_krylov_solve_ex": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                      _P, C.POINTER(SolveInfo)]),
    "mfb_ilu_selftest": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int32), _P, C.c_int64]),
    "mfb_initialize_dx": (C.c_int, [_P, C.c_double, _P, C.c_int]),
    "mfb_update_x_star": (C.c_int, [_P, _P, C.c_int]),
    "mfb_update_dx": (C.c_int, [_P, _P, C.c_int, C.c_double]),
    "mfb_commit_step": (C.c_int, [_P]),
    "mfb_residue_norm": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mfb_write_vtk": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_char_p), _P, _P, C.c_double, C.c_int]),
    "mfb_comm_unique_id": (C.c_int, [_P]),
    "mfb_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mfb_interface_set": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the C-ABI library and declare every prototype of include/metafem_b200.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MfbError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                           "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


def kernel_check(src, cubin_path=None):
    """NVRTC compile-only check (works without a GPU). Returns the log; raises on failure."""
    buf = C.create_string_buffer(1 << 16)
    rc = load().mfb_kernel_check(src.encode(), cubin_path.encode() if cubin_path else None, buf, len(buf))
    if rc != 0:
        raise MfbError(buf.value.decode(errors="replace"))
    return buf.value.decode(errors="replace")


def exported_symbols():
    return sorted(_SIGS)


def ptr(a):
    """Pointer of a numpy array / torch tensor / raw int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


class Context:
    """RAII wrapper of mfb_ctx."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = _P()
        rc = self.lib.mfb_create(C.byref(self.h), device)
        if rc != 0:
            raise MfbError(f"mfb_create failed with code {rc}: no usable CUDA device (there is no CPU fallback)")

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc < 0:
            raise MfbError(f"{name}: {self.lib.mfb_last_error(self.h).decode(errors='replace')}")
        return rc

    def close(self):
        if self.h:
            self.lib.mfb_destroy(self.h)
            self.h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
