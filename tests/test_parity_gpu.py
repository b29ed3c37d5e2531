"""GPU parity tests proper: CUDA path (through the C ABI) vs the oracle on the same seeded inputs."""
import numpy as np
import pytest

from helpers import build_case, product_from_oracle, rel
from oracle import assembly as oasm, solver as osv

pytestmark = pytest.mark.gpu

CASES = [("thermal", (3, 2, 2)), ("linear_elasticity", (3, 2, 2)), ("neo_hookean", (3, 2, 2)),
         ("neo_hookean", (5, 4, 3)), ("thermal", (4, 4, 3)), ("thermo_elasticity", (3, 2, 2))]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}-{'x'.join(map(str, c[1]))}")
def pair(request, built_lib):
    import metafem_b200 as m
    name, n = request.param
    dom, spec, mesh = build_case(name, n)
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
    yield dom, fd
    fd.close()


def test_pattern_bit_exact(pair):
    """CSR sparsity pattern and DOF numbering: bit-exact (canonical CSR of 03_GlobalAssembly.jl:77-168)."""
    dom, fd = pair
    gf = dom.globalfield
    K_I, K_J, K_J_ptr, K_val_ids = fd.get_pattern()
    assert fd.globalfield.nnz == len(gf.K_I)
    assert fd.globalfield.sparse_unitsize == dom.mesh.sparse_unitsize
    assert np.array_equal(K_I, gf.K_I)
    assert np.array_equal(K_J, gf.K_J)
    assert np.array_equal(K_J_ptr, gf.K_J_ptr)
    assert np.array_equal(K_val_ids, np.arange(1, len(K_I) + 1))


def test_x_roundtrip(pair):
    import metafem_b200 as m
    dom, fd = pair
    assert np.array_equal(fd.get_vector(m.lib.VEC_X), dom.globalfield.x)


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


def test_assembly_parity(pair):
    """Residual and assembled entries: 1e-12 relative (norm-wise; atomics change only the summation order)."""
    import metafem_b200 as m
    dom, fd = pair
    _assemble_both(dom, fd)
    gf = dom.globalfield
    TOL = 1e-12
    assert rel(fd.get_vector(m.lib.VEC_X_STAR), gf.x_star) == 0.0
    assert rel(fd.get_vector(m.lib.VEC_RESIDUE), gf.residue) < TOL
    assert rel(fd.get_matrix(m.lib.MAT_K_LINEAR), gf.K_linear[gf.K_val_ids - 1]) < TOL
    Kt_ref = gf.K_total[gf.K_val_ids - 1]
    Kt = fd.get_matrix(m.lib.MAT_K_TOTAL)
    assert rel(Kt, Kt_ref) < TOL
    # per-row check as well: every CSR row within 1e-11 of its own norm
    rows = gf.K_I - 1
    num = np.sqrt(np.bincount(rows, (Kt - Kt_ref) ** 2))
    den = np.sqrt(np.bincount(rows, Kt_ref ** 2))
    assert np.all(num <= 1e-11 * den + 1e-300)


def test_spmv_parity(pair):
    import metafem_b200 as m
    dom, fd = pair
    _assemble_both(dom, fd)
    A = oasm.csr_from_globalfield(dom.globalfield)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(A.shape[0])
    y = np.empty_like(x)
    fd.ctx.call("mfb_spmv", m.lib.MAT_K_TOTAL, m.lib.ptr(x), m.lib.ptr(y), len(x))
    assert rel(y, A @ x) < 1e-13


@pytest.mark.parametrize("method,s", [("idrs", 8), ("bicgstabl_GS", 4)])
def test_solve_parity(pair, method, s):
    """Solutions agree within the solver tolerance (random shadow vectors differ: tolerance-level only)."""
    import scipy.sparse.linalg as spl
    import metafem_b200 as m
    dom, fd = pair
    _assemble_both(dom, fd)
    gf = dom.globalfield
    A = oasm.csr_from_globalfield(gf)
    exact = spl.spsolve(A.tocsc(), gf.residue)
    delta = m.iterative_Solve(fd, Sv_func=method, maxiter=3000, max_pass=10, s=s, want_delta=True)
    info = fd.last_solve
    assert info["converged"], info
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01
    ofun = osv.idrs if method == "idrs" else osv.bicgstabl_GS
    odelta = osv.iterative_Solve(dom, ofun, max_pass=10, maxiter=3000, s=s)
    scale = np.linalg.norm(exact)
    err_o = np.linalg.norm(odelta - exact) / scale
    err_p = np.linalg.norm(delta - exact) / scale
    # both meet the same residual tolerance; their distance to the exact solution is bounded by cond(A)*tol
    assert err_p < max(100 * err_o, 1e-5), (err_p, err_o)


def test_newton_step_parity(pair):
    """update_OneStep! end to end: same Newton history (to tolerance) and same x afterwards."""
    import metafem_b200 as m
    dom, fd = pair
    name = "bicgstabl_GS"
    dom.linear_solver = lambda d: osv.iterative_Solve(d, osv.bicgstabl_GS, maxiter=3000, max_pass=10, s=4)
    fd.linear_solver = lambda d: m.iterative_Solve(d, Sv_func=name, maxiter=3000, max_pass=10, s=4)
    oasm.assemble_X(dom)
    m.assemble_X(fd)
    dom.globalfield.t = 0.0
    fd.globalfield.t = 0.0
    ho = osv.update_OneStep(dom, max_iter=7)
    hp = m.update_OneStep(fd.time_discretization, max_iter=7, fem_domain=fd)
    assert len(ho) == len(hp), (ho, hp)
    assert abs(ho[0] - hp[0]) <= 1e-12 * abs(ho[0])
    xo = dom.globalfield.x
    xp = fd.get_vector(m.lib.VEC_X)
    # converged Newton states agree to the accuracy the tolerances allow (penalty BCs make K ill-conditioned)
    assert ho[-1] < dom.globalfield.converge_tol and hp[-1] < dom.globalfield.converge_tol
    assert np.linalg.norm(xp - xo) / np.linalg.norm(xo) < 1e-4


def test_sparse_ids_by_el_match_the_oracle(pair):
    """elements.sparse_IDs_by_el (03_GlobalAssembly.jl:111-118): the oracle's hash-order entry IDs, taken through K_val_ids to CSR
    positions, are what mfb_sparse_ids_get returns for every variable block."""
    dom, fd = pair
    gf, mesh = dom.globalfield, dom.mesh
    inv = np.empty(len(gf.K_val_ids), np.int64)
    inv[gf.K_val_ids - 1] = np.arange(1, len(inv) + 1)            # reference entry ID -> position in the canonical CSR
    for m in range(len(dom.spec["sparse_mapping"])):
        sid = mesh.sparse_IDs_by_el.astype(np.int64) + m * mesh.sparse_unitsize
        assert np.array_equal(fd.get_sparse_IDs_by_el(m), inv[sid - 1])


def test_assembly_is_bit_reproducible(pair):
    """Deterministic scatter (pairwise accumulators + fixed-order fold, coloured boundary launches): assembling the same state
    twice gives the same bits in K_linear, K_total and the residual -- the reference's atomics do not (06_FEM_Kernel.jl:10,36,74)."""
    import metafem_b200 as m
    dom, fd = pair
    _assemble_both(dom, fd)
    first = [fd.get_matrix(m.lib.MAT_K_LINEAR), fd.get_matrix(m.lib.MAT_K_TOTAL), fd.get_vector(m.lib.VEC_RESIDUE)]
    for _ in range(3):
        td = fd.time_discretization
        fd.K_linear_func(td, fem_domain=fd)
        fd.K_nonlinear_func(td, fem_domain=fd)
        again = [fd.get_matrix(m.lib.MAT_K_LINEAR), fd.get_matrix(m.lib.MAT_K_TOTAL), fd.get_vector(m.lib.VEC_RESIDUE)]
        for a, b in zip(first, again):
            assert np.array_equal(a, b)


def test_solve_is_bit_reproducible(pair):
    """... and with fixed-order reductions in the Krylov kernels the whole solve repeats: same iteration count, same bits."""
    import metafem_b200 as m
    dom, fd = pair
    _assemble_both(dom, fd)
    name, s = ("idrs", 8) if dom.spec["basic_vars"] == ["T"] else ("bicgstabl_GS", 4)
    runs = []
    for _ in range(2):
        d = m.iterative_Solve(fd, Sv_func=name, maxiter=3000, max_pass=10, s=s, seed=99, want_delta=True)
        runs.append((fd.last_solve["iterations"], d.copy()))
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1])
