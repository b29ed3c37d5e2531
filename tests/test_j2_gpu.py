"""J2 plasticity (examples/hypo_elastic_plasticity/J2Plasticity.jl): INTEGRATION_POINT_VAR words, the two-phase
nonlinear update around the quadrature-point callback, and the built-in return map -- CUDA path vs oracle."""
import numpy as np
import pytest

from helpers import build_case, product_from_oracle, rel, j2_states, J2_PARAMS
from oracle import assembly as oasm, solver as osv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["j2", "j2_fused"])
def pair(request, built_lib):
    """"j2": two-phase update around the callback; "j2_fused": the built-in return map inlined into the residual kernel."""
    import metafem_b200 as m
    dom, spec, mesh = build_case(request.param, (4, 2, 2), size=(4.0, 1.0, 1.0))
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
    ost, pst = j2_states(dom, fd)
    yield dom, fd, ost, pst
    fd.close()


def _assemble_both(dom, fd):
    import metafem_b200 as m
    osv.update_Time(dom)
    osv.initialize_dx(dom)
    oasm.K_linear_func(dom)
    osv.update_x_star(dom)
    oasm.K_nonlinear_func(dom)
    td = fd.time_discretization
    m.api.update_Time(fd.globalfield, td)
    gam = np.array(td.gamma_params)
    fd.ctx.call("mfb_initialize_dx", fd.globalfield.dt, m.lib.ptr(gam), len(gam))
    fd.K_linear_func(td, fem_domain=fd)
    al = np.array(td.alpha_params)
    fd.ctx.call("mfb_update_x_star", m.lib.ptr(al), len(al))
    fd.K_nonlinear_func(td, fem_domain=fd)


def test_j2_assembly_parity(pair):
    """Argument arrays, return map, residual and (linear) tangent at a state that yields in part of the points;
    then commit the state and do it again from the hardened state."""
    import metafem_b200 as m
    dom, fd, ost, pst = pair
    gf = dom.globalfield
    for rnd in range(2):
        _assemble_both(dom, fd)
        assert 0 < ost.n_yielded < ost.Y.size, "the test state must be partly plastic"
        assert pst.n_yielded == ost.n_yielded
        call = dom.spec["blocks"][0]["qp_calls"][0]
        for k in range(6):
            assert rel(fd.qp_get(call["outs"][k]), ost.ep_eval[k]) < 1e-12
        assert rel(fd.get_vector(m.lib.VEC_RESIDUE), gf.residue) < 1e-12
        assert rel(fd.get_matrix(m.lib.MAT_K_LINEAR), gf.K_linear[gf.K_val_ids - 1]) < 1e-12
        assert rel(fd.get_matrix(m.lib.MAT_K_TOTAL), gf.K_total[gf.K_val_ids - 1]) < 1e-12
        ost.update_States()
        pst.update_States()
        for k in range(6):
            assert rel(pst.state(f"ep{k + 1}"), ost.ep[k]) < 1e-12
            assert rel(pst.state(f"b{k + 1}"), ost.b[k]) < 1e-12
        assert rel(pst.state("Y"), ost.Y) < 1e-14
        # a different displacement state for the second round
        dom.globalfield.x[:gf.basicfield_size] *= 1.3
        fd.set_vector(m.lib.VEC_X, dom.globalfield.x)


def test_j2_argument_arrays(pair):
    """Phase A of the two-phase update fills e11 e12 e13 e22 e23 e33 at the quadrature points (reference order)."""
    import metafem_b200 as m
    dom, fd, ost, pst = pair
    if pst.fused:
        pytest.skip("no argument arrays when the callback is fused into the kernel")
    _assemble_both(dom, fd)
    blk = dom.spec["blocks"][0]
    cx = oasm._block_context(dom, blk)
    env = oasm._declare_vars(dom, blk, cx, "nonlinear")
    call = blk["qp_calls"][0]
    for a, n in zip(call["args"], call["arg_names"]):
        assert rel(fd.qp_get(n), oasm._eval(a, env, cx.w.shape)) < 1e-12


def test_j2_host_callback_equals_builtin(pair):
    """The generic callback path (host function on the argument arrays) gives the same residual as the built-in kernel."""
    import metafem_b200 as m
    from oracle import j2 as oj2
    dom, fd, ost, pst = pair
    if pst.fused:
        pytest.skip("the fused kernel has no callback seam")
    fd.K_nonlinear_func(fd.time_discretization, fem_domain=fd)
    r_builtin = fd.get_vector(m.lib.VEC_RESIDUE)
    user = oj2.MaterialState(fd.qp_shape, **J2_PARAMS)
    for k in range(6):
        user.ep[k][:] = pst.state(f"ep{k + 1}")
        user.b[k][:] = pst.state(f"b{k + 1}")
    user.Y[:] = pst.state("Y")
    fd.callbacks["strain_updater"] = m.api.HostCallback(user)
    try:
        fd.K_nonlinear_func(fd.time_discretization, fem_domain=fd)
    finally:
        fd.callbacks["strain_updater"] = pst
    assert rel(fd.get_vector(m.lib.VEC_RESIDUE), r_builtin) < 1e-13


def test_j2_load_steps(pair):
    """Incremental loading with history variables: a few update_OneStep! + update_States! cycles per load level, both
    sides with their own Krylov solver -> solver-tolerance agreement of displacements and plastic strain."""
    import metafem_b200 as m
    dom, fd, ost, pst = pair
    dom.linear_solver = lambda d: osv.iterative_Solve(d, osv.bicgstabl_GS, maxiter=2000, max_pass=20, s=8)
    fd.linear_solver = lambda d: m.iterative_Solve(d, Sv_func="bicgstabl_GS", maxiter=2000, max_pass=20, s=8)
    for k in dom.cp:
        if k[0] == "d" and k[1] in "123":
            dom.cp[k][:] = 0.0
            fd.controlpoints[k][:] = 0.0
    oasm.assemble_X(dom)
    m.assemble_X(fd)
    ost.reset()
    pst.reset()
    dom.globalfield.t = fd.globalfield.t = 0.0
    for load in (80.0, 130.0, 60.0):
        dom.cp["sl1"][:] = load
        fd.controlpoints["sl1"][:] = load
        for _ in range(6):
            ho = osv.update_OneStep(dom, max_iter=3)
            hp = m.update_OneStep(fd.time_discretization, max_iter=3, fem_domain=fd)
            ost.update_States()
            pst.update_States()
            assert abs(ho[0] - hp[0]) <= 1e-6 * abs(ho[0]) + 10 * dom.globalfield.converge_tol
    xo, xp = dom.globalfield.x, fd.get_vector(m.lib.VEC_X)
    n = dom.globalfield.basicfield_size
    assert np.linalg.norm(xp[:n] - xo[:n]) / np.linalg.norm(xo[:n]) < 1e-3
    assert ost.Y.max() > J2_PARAMS["Y_initial"] + 1.0, "the load path must have hardened some points"
    assert np.abs(pst.state("Y") - ost.Y).max() < 1e-2 * (ost.Y.max() - J2_PARAMS["Y_initial"])
    assert np.abs(pst.state("ep1") - ost.ep[0]).max() < 1e-2 * np.abs(ost.ep[0]).max()


@pytest.mark.parametrize("case", ["j2", "j2_fused"])
def test_j2_cuda_path_matches_analytical_table(built_lib, case):
    """The script's loading loop (J2Plasticity.jl:264-283) entirely on the CUDA path -- update_OneStep!, built-in return
    map, update_States!, dessemble_X! -- against the script's analytical load-displacement table (:223-228, group 1:
    isotropic hardening, loading into yield and partial unloading)."""
    import metafem_b200 as m
    EY = 100e3
    s_tests, d1_ana = [40, 80, 100, 120, 140, 180, 200, 180, 100], [4, 8, 10, 16, 22, 34, 40, 38, 30]
    dom, spec, mesh = build_case(case, (5, 2, 2), size=(10.0, 1.0, 1.0))
    for k in dom.cp:
        dom.cp[k][:] = 0.0
    fd = product_from_oracle(dom)
    try:
        fd.globalfield.converge_tol, fd.globalfield.dt = 1e-3, 1.0
        m.assemble_Global_Variables(fd)
        m.compile_Updater_GPU(1, fd)
        st = m.api.J2MaterialState(fd, Y_initial=100.0, lam=0.0, mu=EY / 2, Eb=0.0, Ep=EY / 2, f_res=1.0)
        fd.callbacks["strain_updater"] = st
        fd.linear_solver = lambda d: m.iterative_Solve(d, Sv_func="bicgstabl_GS!", maxiter=2000, max_pass=20, s=8)
        right = np.abs(mesh.x[0] - 10.0) < 1e-6
        for s, ana in zip(s_tests, d1_ana):
            fd.controlpoints["sl1"][:] = s
            for counter in range(300):
                m.update_OneStep(fd.time_discretization, max_iter=3, fem_domain=fd)
                m.dessemble_X(fd)
                st.update_States()
                if np.abs(fd.controlpoints["d1_t1"]).max() < 1e-4:
                    break
            else:
                pytest.fail(f"load {s}: the viscous relaxation loop did not settle")
            d1 = fd.controlpoints["d1"][right].mean()
            assert abs(d1 - ana * 1e-3) < 0.03 * abs(ana * 1e-3) + 4e-4, (s, d1, ana * 1e-3)
    finally:
        fd.close()
