"""BASELINE.json configs at their FULL sizes on the CUDA path, checked through size-independent properties (the oracle
cannot run these sizes in seconds): pattern sortedness and counts, rigid-body null space, symmetry, SpMV linearity,
true-residual check of the solve, return-map consistency of the J2 history update."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _domain(spec, n, groups, **kw):
    import metafem_b200 as m
    from metafem_jl_b200.frontend import mesh as fmesh
    tables = fmesh.box_tables((1.0, 1.0, 1.0), (n, n, n), "CUBE", groups=groups, numbering="scattered")
    fd = m.FEM_Domain(tables, spec, **kw)
    return fd, tables


def _prepare(fd):
    import metafem_b200 as m
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
    td = fd.time_discretization
    m.api.update_Time(fd.globalfield, td)
    gam, al = np.array(td.gamma_params), np.array(td.alpha_params)
    fd.ctx.call("mfb_initialize_dx", fd.globalfield.dt, m.lib.ptr(gam), len(gam))
    fd.K_linear_func(td, fem_domain=fd)
    fd.ctx.call("mfb_update_x_star", m.lib.ptr(al), len(al))
    fd.K_nonlinear_func(td, fem_domain=fd)


def _spmv(fd, which, x):
    import metafem_b200 as m
    y = np.empty_like(x)
    fd.ctx.call("mfb_spmv", which, m.lib.ptr(x), m.lib.ptr(y), len(x))
    return y


def test_linear_elasticity_1M_dof(built_lib):
    """configs[1]: 3-D linear elasticity, 43^3 hex20 = 1 004 784 DOF, nnz = 9 * 18 866 380."""
    import metafem_b200 as m
    from metafem_jl_b200.frontend import weakform as wf
    E, nu = 1.0, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    fd, tables = _domain(wf.linear_elasticity(lam, mu, 1000.0 * E, fixed_bg=1, traction_bgs=((2, "sl"),)), 43, ("left", "right"))
    try:
        fd.controlpoints["sl1"][:] = 0.01
        fd.globalfield.converge_tol = 1e-8
        _prepare(fd)
        gf = fd.globalfield
        N = tables.variable_size
        assert gf.basicfield_size == 1004784 and gf.nnz == 9 * 18866380
        # pattern: sorted by row then column, row pointer consistent (bit-exactness against the oracle is tested at small sizes)
        K_I, K_J, K_J_ptr, K_val_ids = fd.get_pattern()
        assert K_J_ptr[0] == 1 and K_J_ptr[-1] == gf.nnz + 1 and np.all(np.diff(K_J_ptr) > 0)
        assert np.all(np.diff(K_I) >= 0)
        same_row = np.diff(K_I) == 0
        assert np.all(np.diff(K_J)[same_row] > 0)
        assert np.array_equal(K_I[K_J_ptr[:-1] - 1], np.arange(1, gf.basicfield_size + 1, dtype=np.int32))
        del K_I, K_J, K_val_ids
        rng = np.random.default_rng(3)
        x, y = rng.standard_normal(gf.basicfield_size), rng.standard_normal(gf.basicfield_size)
        Kx, Ky = _spmv(fd, m.lib.MAT_K_TOTAL, x), _spmv(fd, m.lib.MAT_K_TOTAL, y)
        # linearity and symmetry of the assembled operator
        assert np.linalg.norm(_spmv(fd, m.lib.MAT_K_TOTAL, 2.0 * x - 3.0 * y) - (2.0 * Kx - 3.0 * Ky)) < 1e-12 * np.linalg.norm(Kx)
        assert abs(y @ Kx - x @ Ky) < 1e-11 * abs(y @ Kx)
        # rigid translations are in the null space of the stiffness except on the penalty face (x = 0)
        free = np.tile(tables.x[0] > 1e-9, 3)
        scale = np.abs(Kx).max()
        for c in range(3):
            t = np.zeros(gf.basicfield_size)
            t[c * N:(c + 1) * N] = 1.0
            assert np.abs(_spmv(fd, m.lib.MAT_K_TOTAL, t)[free]).max() < 1e-10 * scale
        # solve with the cantilever script's solver and check the TRUE residual with the library's own SpMV
        res = fd.get_vector(m.lib.VEC_RESIDUE)
        delta = m.iterative_Solve(fd, Sv_func="idrs!", maxiter=2000, max_pass=20, s=8, want_delta=True)
        assert fd.last_solve["converged"], fd.last_solve
        r = res - _spmv(fd, m.lib.MAT_K_TOTAL, delta)
        assert np.linalg.norm(r) / np.sqrt(len(r)) < 1.01 * gf.converge_tol
    finally:
        fd.close()


def test_thermo_elasticity_4M_dof(built_lib):
    """configs[3]: coupled thermo-mechanical assembly, 62^3 hex20 = 3 953 124 DOF, 16 blocks, nnz = 16 * 56 312 245."""
    import metafem_b200 as m
    from metafem_jl_b200.frontend import weakform as wf
    fd, tables = _domain(wf.thermo_elasticity(fixed_bg=1, thermal_bg=2), 62, ("left", "right"))
    try:
        N = tables.variable_size
        fd.controlpoints["T"][:] = 20.0 * np.cos(tables.x[1])
        fd.controlpoints["Te"][:] = 300.0
        for i, b in enumerate(("d1", "d2", "d3")):
            fd.controlpoints[b][:] = 1e-3 * np.sin(1.3 * tables.x[(i + 1) % 3] + 0.2 * i) * tables.x[0]
        fd.globalfield.dt, fd.globalfield.converge_tol = 1.0, 1e-6
        _prepare(fd)
        gf = fd.globalfield
        assert gf.basicfield_size == 3953124 and gf.nnz == 16 * 56312245 and gf.max_time_level == 1
        rng = np.random.default_rng(4)
        x, y = rng.standard_normal(gf.basicfield_size), rng.standard_normal(gf.basicfield_size)
        Kx, Ky = _spmv(fd, m.lib.MAT_K_TOTAL, x), _spmv(fd, m.lib.MAT_K_TOTAL, y)
        assert np.linalg.norm(_spmv(fd, m.lib.MAT_K_TOTAL, x + 0.5 * y) - (Kx + 0.5 * Ky)) < 1e-12 * np.linalg.norm(Kx)
        # the weak form is linear: K_total == K_linear, and the residual is affine in x_star
        assert np.array_equal(_spmv(fd, m.lib.MAT_K_LINEAR, x), Kx)
        # variable order is the string sort [T, d1, d2, d3] (02_LocalAssembly.jl:93): a uniform translation of d leaves the
        # mechanical rows untouched away from the penalty face; a uniform temperature does not (thermal strain)
        free = tables.x[0] > 1e-9
        t = np.zeros(gf.basicfield_size)
        t[1 * N:2 * N] = 1.0
        Kt = _spmv(fd, m.lib.MAT_K_TOTAL, t).reshape(4, N)
        # stiffness part of d1 vanishes; what remains at K_params = (1, 1/dt) is the damping term rho c / dt * M
        r_d1 = fd.get_vector(m.lib.VEC_RESIDUE)
        assert np.isfinite(r_d1).all() and np.isfinite(Kt).all()
        assert np.abs(Kt[2:, free]).max() < 1e-9 * np.abs(Kt[1]).max()       # no coupling d1 -> d2, d3 from a translation
        res = fd.get_vector(m.lib.VEC_RESIDUE)
        delta = m.iterative_Solve(fd, Sv_func="bicgstabl_GS!", maxiter=2000, max_pass=20, s=8, want_delta=True)
        assert fd.last_solve["converged"], fd.last_solve
        r = res - _spmv(fd, m.lib.MAT_K_TOTAL, delta)
        assert np.linalg.norm(r) / np.sqrt(len(r)) < 1.01 * gf.converge_tol
    finally:
        fd.close()


def test_j2_plasticity_8M_dof(built_lib):
    """configs[4]: history variables at the quadrature points, 88^3 hex20 = 8 388 339 DOF, 18.4 M quadrature points."""
    import metafem_b200 as m
    from metafem_jl_b200.frontend import weakform as wf
    fd, tables = _domain(wf.j2_plasticity(fixed_bg=1, traction_bg=2), 88, ("left", "right"))
    try:
        for i, b in enumerate(("d1", "d2", "d3")):
            fd.controlpoints[b][:] = 4e-4 * np.sin(1.3 * tables.x[(i + 1) % 3] + 0.2 * i) * tables.x[0] * 4.0
        fd.controlpoints["sl1"][:] = 120.0
        fd.globalfield.dt, fd.globalfield.converge_tol = 1.0, 1e-3
        m.assemble_Global_Variables(fd)
        m.compile_Updater_GPU(1, fd)
        gf = fd.globalfield
        assert gf.basicfield_size == 8388339 and gf.nnz == 9 * 160558465 and gf.max_time_level == 2
        st = m.api.J2MaterialState(fd, Y_initial=100.0, lam=0.0, mu=50e3, Eb=12.5e3, Ep=25e3, f_res=1.0)
        fd.callbacks["strain_updater"] = st
        _prepare(fd)
        n_qp = fd.qp_shape[0] * fd.qp_shape[1]
        assert n_qp == 27 * 88 ** 3
        n1 = st.n_yielded
        assert 0 < n1 < n_qp, "the state must be partly plastic"
        r1 = fd.get_vector(m.lib.VEC_RESIDUE)
        # the trial evaluation does not touch the committed state: evaluating again gives the same points and residual
        fd.K_nonlinear_func(fd.time_discretization, fem_domain=fd)
        assert st.n_yielded == n1
        r2 = fd.get_vector(m.lib.VEC_RESIDUE)
        assert np.linalg.norm(r2 - r1) < 1e-12 * np.linalg.norm(r1)        # atomics order only
        # radial return with linear hardening lands ON the yield surface: after committing, the same strain is elastic
        st.update_States()
        fd.K_nonlinear_func(fd.time_discretization, fem_domain=fd)
        assert st.n_yielded == 0
        Y = st.state("Y")
        assert Y.min() >= 100.0 and (Y > 100.0).sum() == n1
        # and the residual did not change: the committed plastic strain is the one the trial evaluation used
        r3 = fd.get_vector(m.lib.VEC_RESIDUE)
        assert np.linalg.norm(r3 - r1) < 1e-10 * np.linalg.norm(r1)
    finally:
        fd.close()
