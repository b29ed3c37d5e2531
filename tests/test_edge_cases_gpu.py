"""Edge cases and error behaviour of the C ABI on a device: call-order errors, bad arguments, NVRTC failures, the
"not converged" warning, a single-element mesh, an empty boundary group, repeated assembly (idempotence)."""
import ctypes as C

import numpy as np
import pytest

from helpers import build_case, product_from_oracle, rel
from oracle import assembly as oasm, solver as osv

pytestmark = pytest.mark.gpu


def _both(dom, fd):
    import metafem_b200 as m
    osv.update_Time(dom)
    osv.initialize_dx(dom)
    oasm.K_linear_func(dom)
    osv.update_x_star(dom)
    oasm.K_nonlinear_func(dom)
    td = fd.time_discretization
    m.api.update_Time(fd.globalfield, td)
    gam, al = np.array(td.gamma_params), np.array(td.alpha_params)
    fd.ctx.call("mfb_initialize_dx", fd.globalfield.dt, m.lib.ptr(gam), len(gam))
    fd.K_linear_func(td, fem_domain=fd)
    fd.ctx.call("mfb_update_x_star", m.lib.ptr(al), len(al))
    fd.K_nonlinear_func(td, fem_domain=fd)


def test_call_order_and_argument_errors(built_lib):
    import metafem_b200 as m
    L = m.lib
    ctx = L.Context(0)
    try:
        nnz, unit = C.c_int64(0), C.c_int64(0)
        mapping = np.zeros(2, np.int32)
        # pattern before mesh, assembly before kernels, solve before pattern: MFB_ERR_STATE (-4) with a message
        assert ctx.lib.mfb_pattern_build(ctx.h, 1, 0, 1, L.ptr(mapping), C.byref(nnz), C.byref(unit)) == -4
        assert b"mfb_mesh_set" in ctx.lib.mfb_last_error(ctx.h)
        kp = np.ones(1)
        assert ctx.lib.mfb_assemble_nonlinear(ctx.h, L.ptr(kp), 1, 0.0, 1.0) == -4
        info = L.SolveInfo()
        assert ctx.lib.mfb_krylov_solve(ctx.h, L.MFB_IDRS, 4, 10, 1, 1e-8, 1, None, C.byref(info)) == -4
        assert ctx.lib.mfb_mesh_set(ctx.h, 0, 0, 0, 0, None, None, None, None, None, None) == -3      # MFB_ERR_ARG
        with pytest.raises(L.MfbError):
            ctx.call("mfb_field_set", b"s", None)
    finally:
        ctx.close()
    assert built_lib.mfb_create(C.byref(C.c_void_p()), 10 ** 6) == -3                                  # no such device


def test_bad_kernel_source_and_unknown_method(built_lib):
    import metafem_b200 as m
    dom, spec, mesh = build_case("thermal", (2, 1, 1))
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    try:
        m.assemble_Global_Variables(fd)
        desc = (m.lib.BlockDesc * 1)()
        desc[0].kind = 0
        desc[0].nonlinear_kernel = b"nope"
        rc = fd.ctx.lib.mfb_kernel_compile(fd.ctx.h, b'#include "mfb_skeleton.cuh"\nint x = ;', 1, desc)
        assert rc == -2 and b"error" in fd.ctx.lib.mfb_last_error(fd.ctx.h).lower()                  # MFB_ERR_NVRTC + log
        m.compile_Updater_GPU(1, fd)
        _both(dom, fd)
        info = m.lib.SolveInfo()
        assert fd.ctx.lib.mfb_krylov_solve(fd.ctx.h, 99, 4, 10, 1, 1e-8, 1, None, C.byref(info)) == -3
        assert fd.ctx.lib.mfb_krylov_solve(fd.ctx.h, m.lib.MFB_IDRS, 40, 10, 1, 1e-8, 1, None, C.byref(info)) == -3   # s too large
        with pytest.raises(ValueError):
            m.iterative_Solve(fd, Sv_func="conjugate_gradient!")
        # one iteration cannot converge: MFB_NOT_CONVERGED (1) is a warning, the result is still delivered (02_Preconditioner.jl:66-68)
        delta = np.empty(dom.globalfield.basicfield_size)
        rc = fd.ctx.lib.mfb_krylov_solve(fd.ctx.h, m.lib.MFB_BICGSTABL_GS, 2, 1, 1, 1e-14, 1, m.lib.ptr(delta), C.byref(info))
        assert rc == 1 and info.converged == 0 and np.isfinite(delta).all() and info.residual > 1e-14
        # wrong vector length
        assert fd.ctx.lib.mfb_vector_get(fd.ctx.h, m.lib.VEC_X, m.lib.ptr(delta), 3) == -3
    finally:
        fd.close()


@pytest.mark.parametrize("name", ["thermal", "neo_hookean"])
def test_single_cube_mesh(built_lib, name):
    """Smallest meshes (1 hex20 element / 5 tet10 elements): every node is on the boundary, every row is short."""
    import metafem_b200 as m
    dom, spec, mesh = build_case(name, (1, 1, 1), size=(1.0, 1.0, 1.0))
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    try:
        m.assemble_Global_Variables(fd)
        m.compile_Updater_GPU(1, fd)
        _both(dom, fd)
        gf = dom.globalfield
        K_I, K_J, K_J_ptr, _ = fd.get_pattern()
        assert np.array_equal(K_I, gf.K_I) and np.array_equal(K_J, gf.K_J) and np.array_equal(K_J_ptr, gf.K_J_ptr)
        assert rel(fd.get_vector(m.lib.VEC_RESIDUE), gf.residue) < 1e-12
        assert rel(fd.get_matrix(m.lib.MAT_K_TOTAL), gf.K_total[gf.K_val_ids - 1]) < 1e-12
    finally:
        fd.close()


def test_empty_boundary_group_and_repeated_assembly(built_lib):
    """A boundary group without facets contributes nothing (and does not fail); assembling twice gives the same result
    up to the order of the atomic additions (K_linear/K_total/residue are re-initialised by every call)."""
    import metafem_b200 as m
    dom, spec, mesh = build_case("neo_hookean", (3, 2, 2))
    mesh.bg_fIDs[2] = mesh.bg_fIDs[2][:0]                        # the traction group loses all its facets
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    try:
        m.assemble_Global_Variables(fd)
        m.compile_Updater_GPU(1, fd)
        _both(dom, fd)
        gf = dom.globalfield
        r1, K1 = fd.get_vector(m.lib.VEC_RESIDUE), fd.get_matrix(m.lib.MAT_K_TOTAL)
        assert rel(r1, gf.residue) < 1e-12 and rel(K1, gf.K_total[gf.K_val_ids - 1]) < 1e-12
        fd.K_linear_func(fd.time_discretization, fem_domain=fd)
        fd.K_nonlinear_func(fd.time_discretization, fem_domain=fd)
        assert rel(fd.get_vector(m.lib.VEC_RESIDUE), r1) < 1e-14 and rel(fd.get_matrix(m.lib.MAT_K_TOTAL), K1) < 1e-14
    finally:
        fd.close()
