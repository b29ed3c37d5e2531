"""CPU checks of the oracle's Krylov restatements (all exported Sv_func! choices and preconditioner modes) on a small
assembled system: every method must reach the absolute tolerance of iterative_Solve! and agree with a direct solve."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

from helpers import build_case
from oracle import assembly as oasm, solver as osv


@pytest.fixture(scope="module")
def system():
    dom, spec, mesh = build_case("thermal", (3, 2, 2))
    oasm.assemble_Global_Variables(dom)
    osv.update_Time(dom)
    osv.initialize_dx(dom)
    oasm.K_linear_func(dom)
    osv.update_x_star(dom)
    oasm.K_nonlinear_func(dom)
    gf = dom.globalfield
    A = oasm.csr_from_globalfield(gf)
    return dom, A, spl.spsolve(A.tocsc(), gf.residue)


METHODS = [("idrs", dict(s=8)), ("bicgstabl_GS", dict(s=4)), ("bicgstabl", dict(s=4)), ("gmres", dict(s=20)), ("cgs", {}),
           ("cgs2", {}), ("tfqmr", dict(checkiter=20)), ("lsqr", {})]


@pytest.mark.parametrize("name,kw", METHODS, ids=[m[0] for m in METHODS])
def test_oracle_methods_converge(system, name, kw):
    dom, A, exact = system
    gf = dom.globalfield
    delta = osv.iterative_Solve(dom, getattr(osv, name), max_pass=10, maxiter=4000, **kw)
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01, dom.last_solve
    assert np.linalg.norm(delta - exact) / np.linalg.norm(exact) < 1e-5


@pytest.mark.parametrize("pr,pl", [(osv.Pr_Jacobi_column, None), (None, None), (osv.Pr_Jacobi, osv.Pl_Jacobi),
                                   (None, lambda A: osv.Pl_Jacobi(A, normalized_by_row=True))],
                         ids=["Pr_column", "Pr_Identity", "Pl_Jacobi", "Pl_row"])
def test_oracle_preconditioner_modes(system, pr, pl):
    dom, A, exact = system
    gf = dom.globalfield
    delta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4, Pr_func=pr, Pl_func=pl)
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01, dom.last_solve
    assert np.linalg.norm(delta - exact) / np.linalg.norm(exact) < 1e-5


def test_oracle_pl_ilu(system):
    """Pl_ILU (02_Preconditioner.jl:179-194): the restated zero-fill factorisation reproduces A on its pattern, and the
    left-preconditioned solve reaches the tolerance in fewer iterations than with Jacobi alone."""
    import scipy.sparse as sps
    dom, A, exact = system
    gf = dom.globalfield
    P = osv.Pl_ILU(A)
    LU = (P.L @ P.U).tocsr()
    mask = sps.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), shape=A.shape)
    assert abs(LU.multiply(mask) - A).max() < 1e-12 * abs(A).max()
    delta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4, Pl_func=osv.Pl_ILU)
    it_ilu = sum(dom.last_solve["iters"])
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01, dom.last_solve
    osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4)
    assert it_ilu < sum(dom.last_solve["iters"])


@pytest.mark.parametrize("order", ["color", "hash"])
def test_oracle_block_ilu_in_the_library_order(system, order):
    """The restatement of the LIBRARY's Pl_ILU (block ILU(0), elimination by colour classes of a greedy colouring in hash
    priority): (L U)_ij = A_ij on the block pattern, the dependency depth is bounded by the number of colours, and the
    left-preconditioned solve reaches the tolerance in no more iterations than the reference-order scalar ILU needs plus a
    margin (the order is chosen for parallelism; it must not cost convergence)."""
    dom, A, exact = system
    gf = dom.globalfield
    nv = len(dom.spec["basic_vars"])
    P = osv.Pl_ILU_block(A, nv, order=order)
    assert P.product_defect() < 1e-12
    if order == "color":
        assert P.levels <= P.n_colors
    delta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4, Pl_func=lambda M: osv.Pl_ILU_block(M, nv, order=order))
    it_block = sum(dom.last_solve["iters"])
    r = gf.residue - A @ delta
    assert np.linalg.norm(r) / np.sqrt(len(r)) < gf.converge_tol * 1.01, dom.last_solve
    osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=10, maxiter=4000, s=4)
    assert it_block <= sum(dom.last_solve["iters"]), (it_block, dom.last_solve)
