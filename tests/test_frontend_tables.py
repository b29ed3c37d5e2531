"""CPU tests: the package's closed-form element tables / mesh builder against the oracle's restatement of the
reference's polynomial-algebra path (two independent derivations of the same tables)."""
import numpy as np

import metafem_b200  # noqa: F401
from metafem_jl_b200.frontend import elements, mesh as fmesh
from oracle import discretization as D, refgeom as rg, femmesh as fm


def test_hex20_tables_match_oracle():
    sp = D.initialize_Classical_Element(3, "CUBE", 2, 1, 5, "Serendipity")
    t = elements.hex20_tables()
    sl = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    for s in sl:
        assert np.abs(t.ref_itp_vals[(slice(None), slice(None)) + s] - sp.ref_itp_vals[(slice(None), slice(None)) + s]).max() < 1e-13
    assert np.abs(t.itg_weight - sp.itg_weight).max() < 1e-15
    for f in range(6):
        for s in sl:
            a = t.bdy_ref_itp_vals[(slice(None), slice(None)) + s + (f,)]
            b = sp.bdy_ref_itp_vals[f][(slice(None), slice(None)) + s]
            assert np.abs(a - b).max() < 1e-13
        assert np.abs(t.bdy_itg_weights[:, f] - sp.bdy_itg_weights[f]).max() < 1e-15
        assert np.array_equal(t.bdy_tangent_directions[..., f], sp.bdy_tangent_directions[f])


def test_tet10_tables_match_oracle():
    sp = D.initialize_Classical_Element(3, "SIMPLEX", 2, 1, 5)
    t = elements.tet10_tables()
    sl = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    for s in sl:
        assert np.abs(t.ref_itp_vals[(slice(None), slice(None)) + s] - sp.ref_itp_vals[(slice(None), slice(None)) + s]).max() < 1e-13
    assert np.abs(t.itg_weight - sp.itg_weight).max() < 1e-16
    for f in range(4):
        for s in sl:
            a = t.bdy_ref_itp_vals[(slice(None), slice(None)) + s + (f,)]
            b = sp.bdy_ref_itp_vals[f][(slice(None), slice(None)) + s]
            assert np.abs(a - b).max() < 1e-13
        assert np.abs(t.bdy_itg_weights[:, f] - sp.bdy_itg_weights[f]).max() < 1e-15
        assert np.abs(t.bdy_tangent_directions[..., f] - sp.bdy_tangent_directions[f]).max() < 1e-15


def _canon(cp, x):
    """element -> sorted rows of node coordinates per local node (numbering-independent description)."""
    return np.round(x.T[cp.T - 1], 12)


def test_box_mesh_matches_oracle_up_to_midedge_numbering():
    for shape, n in (("CUBE", (3, 2, 2)), ("SIMPLEX", (2, 2, 3))):
        size = (1.5, 1.0, 1.0)
        c, conn = rg.make_Brick(size, n, shape)
        m = rg.construct_TotalMesh_3D(c, conn)
        fids = rg.get_BoundaryMesh(m)
        cen = rg.face_centroids(m, fids)
        omesh = fm.mesh_Classical(m, [fids[np.abs(cen[0]) < 1e-9], fids[np.abs(cen[0] - size[0]) < 1e-9]], shape)
        for numbering in ("sorted", "scattered"):
            t = fmesh.box_tables(size, n, shape, groups=("left", "right"), numbering=numbering)
            assert t.x.shape == omesh.x.shape
            nv = c.shape[1]
            assert np.array_equal(t.x[:, :nv], omesh.x[:, :nv])
            # same geometry element by element, local node by local node
            assert np.array_equal(_canon(t.controlpoint_IDs, t.x), _canon(omesh.controlpoint_IDs, omesh.x))
            # same boundary facets (as sets of (element, local face)) per group
            for g in (1, 2):
                a = {(int(t.facet_element_ID[f - 1]), int(t.facet_element_eindex[f - 1])) for f in t.bg_fIDs[g]}
                b = {(int(omesh.facet_element_ID[f - 1]), int(omesh.facet_element_eindex[f - 1])) for f in omesh.bg_fIDs[g]}
                assert a == b
