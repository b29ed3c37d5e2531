"""All exported Krylov methods and preconditioner modes of iterative_Solve! on the CUDA path (through the C ABI), against
a direct solve and the oracle's restatement of the same method: solver-tolerance parity (random shadow vectors differ)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

from helpers import build_case, product_from_oracle
from oracle import assembly as oasm, solver as osv

pytestmark = pytest.mark.gpu


def _make_system(name):
    import metafem_b200 as m
    dom, spec, mesh = build_case(name, (3, 2, 2))
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
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
    gf = dom.globalfield
    A = oasm.csr_from_globalfield(gf)
    return dom, fd, A, spl.spsolve(A.tocsc(), gf.residue)


@pytest.fixture(scope="module", params=["thermal", "linear_elasticity"])
def system(request, built_lib):
    dom, fd, A, exact = _make_system(request.param)
    yield dom, fd, A, exact
    fd.close()


METHODS = [("idrs", dict(s=8)), ("bicgstabl_GS", dict(s=4)), ("bicgstabl", dict(s=4)), ("gmres", dict(s=20)), ("cgs", {}),
           ("cgs2", {}), ("tfqmr", dict(checkiter=20)), ("lsqr", {})]


def _check(dom, fd, A, exact, delta, odelta):
    gf = dom.globalfield
    assert fd.last_solve["converged"], fd.last_solve
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01
    scale = np.linalg.norm(exact)
    err_p, err_o = np.linalg.norm(delta - exact) / scale, np.linalg.norm(odelta - exact) / scale
    assert err_p < max(100 * err_o, 1e-5), (err_p, err_o)


@pytest.mark.parametrize("name,kw", METHODS, ids=[m[0] for m in METHODS])
def test_method_parity(system, name, kw):
    import metafem_b200 as m
    dom, fd, A, exact = system
    maxiter = 20000 if name == "lsqr" else 4000
    delta = m.iterative_Solve(fd, Sv_func=name + "!", maxiter=maxiter, max_pass=10, want_delta=True, **kw)
    odelta = osv.iterative_Solve(dom, getattr(osv, name), max_pass=10, maxiter=maxiter, **kw)
    _check(dom, fd, A, exact, delta, odelta)


MODES = [("Pr_Jacobi_column", "Identity", osv.Pr_Jacobi_column, None), ("Identity", "Identity", None, None),
         ("Pr_Jacobi", "Pl_Jacobi", osv.Pr_Jacobi, osv.Pl_Jacobi),
         ("Identity", "Pl_Jacobi_row", None, lambda A: osv.Pl_Jacobi(A, normalized_by_row=True))]


@pytest.mark.parametrize("pr,pl,opr,opl", MODES, ids=[f"{a}-{b}" for a, b, _, _ in MODES])
def test_preconditioner_modes(system, pr, pl, opr, opl):
    import metafem_b200 as m
    dom, fd, A, exact = system
    delta = m.iterative_Solve(fd, Sv_func="bicgstabl_GS", Pr_func=pr, Pl_func=pl, maxiter=4000, max_pass=10, s=4, want_delta=True)
    odelta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4, Pr_func=opr, Pl_func=opl)
    _check(dom, fd, A, exact, delta, odelta)


def test_idrs_original_follows_the_reference_operation_by_operation(system):
    """idrs_original! (04_IDRs.jl:97-169) is exported but "not used": it is restated as written (including :147-148, which
    replaces r by A Q), so the check is not convergence but that the CUDA path and the oracle produce the same iterate after
    the same number of iterations FROM THE SAME shadow vectors (the library's seeded FEM_rand, reproduced in numpy)."""
    import metafem_b200 as m
    dom, fd, A, exact = system
    gf = dom.globalfield
    N, nv, s, seed, iters = dom.mesh.variable_size, len(dom.spec["basic_vars"]), 4, 4321, 3
    delta = m.iterative_Solve(fd, Sv_func="idrs_original!", maxiter=iters, max_pass=1, s=s, seed=seed, want_delta=True)
    assert fd.last_solve["iterations"] == iters
    P = [osv.fem_rand_seeded(N, nv, seed, 1 * 64 + k) for k in range(s)]
    odelta = osv.iterative_Solve(dom, osv.idrs_original, max_pass=1, maxiter=iters, s=s, P=P)
    assert dom.last_solve["iters"] == [iters]
    assert np.linalg.norm(delta - odelta) <= 1e-8 * np.linalg.norm(odelta), np.linalg.norm(delta - odelta) / np.linalg.norm(odelta)


def test_pl_ilu_factorisation_property_and_solve(system):
    """Pl_ILU on the CUDA path: (L U)_ij = A_ij on the pattern (the defining property of a zero-fill incomplete factorisation,
    checked by a device kernel on K_total), a few dozen dependency levels, and the left-preconditioned bicgstabl_GS! reaches the
    same solution as the direct solve in fewer iterations than with the Jacobi scaling alone."""
    import ctypes as C
    import metafem_b200 as m
    dom, fd, A, exact = system
    defect, levels = C.c_double(1.0), C.c_int32(0)
    fd.ctx.call("mfb_ilu_selftest", C.byref(defect), C.byref(levels), None, 0)
    assert defect.value < 1e-12 and 1 <= levels.value < 200, (defect.value, levels.value)
    # U^-1 L^-1 (A v) ~ v to the extent the factorisation is complete: a contraction, not an identity; it must at least beat Jacobi
    delta = m.iterative_Solve(fd, Sv_func="bicgstabl_GS", Pl_func="Pl_ILU", maxiter=4000, max_pass=10, s=4, want_delta=True)
    it_ilu = fd.last_solve["iterations"]
    odelta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4, Pl_func=osv.Pl_ILU)
    _check(dom, fd, A, exact, delta, odelta)
    m.iterative_Solve(fd, Sv_func="bicgstabl_GS", maxiter=4000, max_pass=10, s=4)
    assert it_ilu < fd.last_solve["iterations"], (it_ilu, fd.last_solve)


def test_pl_ilu_four_variables(built_lib):
    """Pl_ILU with 4 x 4 node blocks (thermo-elasticity: displacement + temperature): the factorisation property on the pattern and a
    left-preconditioned solve whose TRUE residual (numpy, the oracle's matrix) meets the tolerance (the 1- and 3-variable block sizes run in the test above)."""
    import ctypes as C
    import metafem_b200 as m
    dom, fd, A, exact = _make_system("thermo_elasticity")
    try:
        defect, levels = C.c_double(1.0), C.c_int32(0)
        P = osv.Pl_ILU_block(A, 4)
        v = np.random.default_rng(12).standard_normal(A.shape[0])
        want = P(v.copy())
        got = np.ascontiguousarray(v.copy())
        fd.ctx.call("mfb_ilu_selftest", C.byref(defect), C.byref(levels), m.lib.ptr(got), len(got))
        assert defect.value < 1e-12 and levels.value == P.levels, (defect.value, levels.value, P.levels)
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
        delta = m.iterative_Solve(fd, Sv_func="bicgstabl_GS", Pl_func="Pl_ILU", maxiter=4000, max_pass=10, s=4, want_delta=True)
        it_ilu = fd.last_solve["iterations"]
        assert fd.last_solve["converged"], fd.last_solve
        gf = dom.globalfield
        r = gf.residue - A @ delta
        assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01
        m.iterative_Solve(fd, Sv_func="bicgstabl_GS", maxiter=4000, max_pass=10, s=4)
        assert it_ilu <= fd.last_solve["iterations"], (it_ilu, fd.last_solve)
    finally:
        fd.close()


def test_pl_ilu_application_matches_the_oracle_restatement(system):
    """The preconditioner itself, not only what it does to a solve: U^-1 L^-1 v from the CUDA sweeps (mfb_ilu_selftest) against the
    oracle's restatement of the library's algorithm (oracle/solver.py::Pl_ILU_block: block ILU(0) in the colour-class order of the
    node hash), and the same dependency depth. The CUDA sweeps read FP32-rounded factors (FP64 accumulation): 1e-4 relative; a
    different elimination order, a missed update or a wrong level schedule would be an O(1) difference."""
    import ctypes as C
    import metafem_b200 as m
    import os
    dom, fd, A, exact = system
    nv = len(dom.spec["basic_vars"])
    # the library's A/B switches: MFB_ILU_ORDER=hash (plain hash order), MFB_ILU_FP64=1 (factors of the sweeps in doubles: 1e-10)
    order = "hash" if os.environ.get("MFB_ILU_ORDER", "c")[0] == "h" else "color"
    tol = 1e-10 if os.environ.get("MFB_ILU_FP64") == "1" or os.environ.get("MFB_ILU_UNPACKED") == "1" else 1e-4
    P = osv.Pl_ILU_block(A, nv, order=order)
    rng = np.random.default_rng(11)
    v = rng.standard_normal(A.shape[0])
    want = P(v.copy())
    got = np.ascontiguousarray(v.copy())
    defect, levels = C.c_double(1.0), C.c_int32(0)
    fd.ctx.call("mfb_ilu_selftest", C.byref(defect), C.byref(levels), m.lib.ptr(got), len(got))
    assert levels.value == P.levels, (levels.value, P.levels)
    assert np.abs(got - want).max() <= tol * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
