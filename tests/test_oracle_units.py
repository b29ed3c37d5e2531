"""CPU unit tests of the oracle's building blocks (hash table, polynomial algebra, element tables, pattern,
C loop nests, Krylov solvers, time stepping)."""
import numpy as np
import pytest

from helpers import build_case
from oracle import femdict as fd, discretization as D, refgeom as rg, assembly as oasm, solver as osv


def test_dict_size_policy():
    assert [fd.dict_size(n) for n in (0, 1, 10, 11, 16, 21, 22, 100)] == [16, 16, 16, 32, 32, 32, 64, 256]


def test_wang_hash_is_a_bijection_sample_and_known_structure():
    lib = fd.lib()
    xs = np.arange(1, 5000, dtype=np.uint64)
    hs = np.array([lib.ora_wang64(int(x)) for x in xs], dtype=np.uint64)
    assert len(np.unique(hs)) == len(xs)
    assert lib.ora_wang64(0) == ((~0 & 0xFFFFFFFFFFFFFFFF) * 1 and lib.ora_wang64(0))  # deterministic
    # avalanche: flipping one input bit flips about half of the output bits
    flips = [bin(int(lib.ora_wang64(12345)) ^ int(lib.ora_wang64(12345 ^ (1 << b)))).count("1") for b in range(40)]
    assert 24 < np.mean(flips) < 40


def test_dict_insert_lookup_grow_and_duplicates():
    rng = np.random.default_rng(0)
    d = fd.FemDict()
    keys = rng.integers(1, 2 ** 40, size=5000, dtype=np.uint64)
    keys[100:200] = keys[:100]                                 # duplicates map to the same slot
    ids = d.set_ids(keys)
    assert np.array_equal(ids[100:200], ids[:100])
    assert len(d.keys) == fd.dict_size(5000)
    assert np.array_equal(d.get_ids(keys), ids)
    assert np.all(d.get_ids(np.array([2 ** 50 + 1, 2 ** 50 + 7], dtype=np.uint64)) <= 0)
    d.vals[ids - 1] = np.arange(len(ids), dtype=np.int32)
    vals_before = {int(k): int(d.vals[i - 1]) for k, i in zip(keys, ids)}
    more = rng.integers(2 ** 41, 2 ** 42, size=20000, dtype=np.uint64)
    d.set_ids(more)                                            # forces a re-hash; values must travel with keys
    assert len(d.keys) == fd.dict_size(20000 + len(np.unique(keys)))
    ids2 = d.get_ids(keys)
    assert all(vals_before[int(k)] == int(d.vals[i - 1]) for k, i in zip(keys, ids2))
    occupied = d.total_ids()
    assert len(occupied) == len(np.unique(np.concatenate([keys, more])))


def test_key_packing_roundtrip():
    a = np.array([1, 7, 2 ** 31 - 1], dtype=np.int32)
    b = np.array([5, 2 ** 20, 3], dtype=np.int32)
    k = fd.I32I32_To_UI64(a, b)
    assert np.array_equal(fd.UI64_To_UpperHalf(k), a) and np.array_equal(fd.UI64_To_LowerHalf(k), b)
    assert fd.I4I30I30_To_UI64(0, 0, 0)[()] != 0


def test_polynomial_algebra():
    x, y = D.collect_Basis(2)
    p = (x + 1) * (x - 1)
    assert p.evaluate((3.0, 0.0)) == 8.0
    assert (p * y).derivative((1, 1)).evaluate((2.0, 5.0)) == 4.0
    q = D.substitute_Polynomial(D.Polynomial([1.0, 2.0], [(2,), (0,)]), 1, x + y)     # (x+y)^2 + 2
    assert q.evaluate((1.0, 2.0)) == 11.0
    tiny = D.Polynomial([1.0, 1e-9], [(1, 0), (0, 1)]).check_clear()                   # 1e-8 drop of check_Clear
    assert tiny.orders == [(1, 0)]


@pytest.mark.parametrize("shape,na", [("CUBE", 20), ("SIMPLEX", 10)])
def test_shape_functions_are_nodal_and_complete(shape, na):
    sp = D.initialize_Classical_Element(3, shape, 2, 1, 5, "Serendipity" if shape == "CUBE" else "Lagrange")
    r = sp.ref_itp_vals
    assert r.shape[1] == na
    assert np.abs(r[:, :, 0, 0, 0].sum(1) - 1).max() < 1e-13                            # partition of unity
    for s in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
        assert np.abs(r[(slice(None), slice(None)) + s].sum(1)).max() < 1e-12
    vol = 1.0 if shape == "CUBE" else 1.0 / 6
    assert abs(sp.itg_weight.sum() - vol) < 1e-14
    # degree-5 exactness of the rule on a monomial
    pts = np.array(sp.itg_pos)
    exact = (1 / 3) * (1 / 4) * 1 if shape == "CUBE" else 2 * 3 * 1 / 5040 * 1.0 * 1  # int x^2 y^3 ; tet: 2!3!0!/(2+3+0+3)!
    got = (sp.itg_weight * pts[:, 0] ** 2 * pts[:, 1] ** 3).sum()
    assert abs(got - (exact if shape == "CUBE" else 2 * 6 / 40320)) < 1e-14


def test_box_topology_counts():
    for n in ((2, 3, 4), (1, 1, 1)):
        c, conn = rg.make_Brick((1.0, 1.0, 1.0), n)
        m = rg.construct_TotalMesh_3D(c, conn)
        nx, ny, nz = n
        assert m.segment_vertex_IDs.shape[1] == nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz
        assert m.face_vertex_IDs.shape[1] == nx * ny * (nz + 1) + nx * (ny + 1) * nz + (nx + 1) * ny * nz
        assert len(rg.get_BoundaryMesh(m)) == 2 * (nx * ny + ny * nz + nx * nz)
        # every face's 4 vertices really bound a face of some block
        assert np.all(np.sort(m.face_vertex_IDs, axis=0)[0] > 0)


def test_pattern_is_canonical_csr_and_matches_dense_count():
    dom, spec, mesh = build_case("linear_elasticity", (2, 2, 2))
    oasm.assemble_Global_Variables(dom)
    gf = dom.globalfield
    N = mesh.variable_size
    pairs = set()
    cp = mesh.controlpoint_IDs
    for e in range(cp.shape[1]):
        for a in cp[:, e]:
            for b in cp[:, e]:
                pairs.add((int(a), int(b)))
    assert mesh.sparse_unitsize == len(pairs)
    assert len(gf.K_I) == 9 * len(pairs)
    key = gf.K_I.astype(np.int64) * (3 * N + 1) + gf.K_J
    assert np.all(np.diff(key) > 0)                                                    # sorted, unique
    assert gf.K_J_ptr[0] == 1 and gf.K_J_ptr[-1] == len(gf.K_I) + 1
    assert np.array_equal(np.diff(gf.K_J_ptr), np.bincount(gf.K_I - 1, minlength=3 * N))
    # sparse_IDs_by_el points at the right (row, col)
    sid = mesh.sparse_IDs_by_el
    inv = np.empty(len(gf.K_I), np.int64); inv[gf.K_val_ids - 1] = np.arange(len(gf.K_I))
    e, a, b = 3, 5, 17
    pos = inv[sid[a, b, e] - 1 + 4 * mesh.sparse_unitsize]                             # block (1,1)
    assert gf.K_I[pos] == cp[a, e] + N and gf.K_J[pos] == cp[b, e] + N


def test_c_loop_nest_matches_numpy_contraction():
    dom, spec, mesh = build_case("neo_hookean", (2, 2, 1))
    oasm.assemble_Global_Variables(dom)
    osv.update_Time(dom); osv.initialize_dx(dom); osv.update_x_star(dom)
    blk = spec["blocks"][0]
    cx = oasm._block_context(dom, blk)
    rng = np.random.default_rng(1)
    vals = rng.standard_normal(cx.w.shape)
    K1, K2 = np.zeros_like(dom.globalfield.K_total), np.zeros_like(dom.globalfield.K_total)
    term = blk["nonlinear_gradients"][7]
    oasm._kval(dom, K1, term, vals, cx)
    oasm._kval_numpy(dom, K2, term, vals, cx)
    assert np.abs(K1 - K2).max() <= 1e-13 * np.abs(K2).max()


@pytest.mark.parametrize("solver,s", [(osv.idrs, 4), (osv.idrs, 8), (osv.bicgstabl_GS, 2), (osv.bicgstabl_GS, 4)])
def test_krylov_solvers_converge_on_nonsymmetric_system(solver, s):
    import scipy.sparse as sps
    from oracle import cpath
    rng = np.random.default_rng(3)
    n = 400
    A = (sps.random(n, n, density=0.02, random_state=4) + sps.diags(np.linspace(2, 6, n))).tocsr()
    A.sort_indices()
    b = rng.standard_normal(n)
    op = cpath.CsrOperator(A.indptr + 1, A.indices + 1, A.data, n)
    x, r = np.zeros(n), b.copy()
    it = solver(x, op, b, r, tol=1e-10, maxiter=2000, s=s, rng=np.random.default_rng(9))
    assert 0 < it < 2000
    assert np.linalg.norm(b - A @ x) / np.sqrt(n) < 1e-9


def test_generalized_alpha_parameters_and_predictor():
    dom, spec, mesh = build_case("thermal", (2, 2, 2))
    oasm.assemble_Global_Variables(dom)
    dom.globalfield.dt = 0.5
    osv.update_Time(dom)
    assert dom.beta_params == [1.0] and dom.K_params == [1.0] and dom.globalfield.t == 0.5
    osv.initialize_dx(dom)
    assert not dom.globalfield.dx.any()
    osv.update_x_star(dom)
    assert np.array_equal(dom.globalfield.x_star, dom.globalfield.x)
