"""Device-side construction of the second-order mesh tables: against the numpy front end
(bit-exact with its "sorted" mid-edge numbering) and against the oracle's restatement of the reference's hash-ordered
construction (same control points and elements up to the documented permutation of the mid-edge IDs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,n", [("CUBE", (4, 3, 2)), ("SIMPLEX", (3, 2, 2)), ("CUBE", (1, 1, 1)), ("CUBE", (24, 20, 16))])
def test_device_tables_equal_front_end(built_lib, shape, n):
    import metafem_b200 as m
    from metafem_jl_b200.frontend import mesh as fmesh
    size = (1.5, 1.0, 0.75)
    coors, conn = fmesh.make_Brick(size, n, shape)
    ref = fmesh.second_order_tables(coors, conn, [lambda c: np.ones(c.shape[1], bool)], numbering="sorted")
    ctx = m.lib.Context(0)
    try:
        cp, x, f_el, f_eidx, cen = fmesh.second_order_tables_device(ctx, coors, conn)
    finally:
        ctx.close()
    assert np.array_equal(cp, ref.controlpoint_IDs)
    assert np.array_equal(x, ref.x)                        # mid-edge nodes: 0.5 a + 0.5 b, the same expression
    # one group holding every boundary face: the front end lists them in the same (local face, element) order
    assert np.array_equal(f_el, ref.facet_element_ID) and np.array_equal(f_eidx, ref.facet_element_eindex)
    nx, ny, nz = n
    if shape == "CUBE":
        assert len(f_el) == 2 * (nx * ny + ny * nz + nx * nz)
    on = np.zeros(len(f_el), bool)
    for d in range(3):
        on |= (np.abs(cen[d]) < 1e-12) | (np.abs(cen[d] - size[d]) < 1e-12)
    assert on.all()                                        # every centroid lies on the box surface


def test_device_tables_match_oracle_mesh(built_lib):
    """Oracle = the reference's construction (hash-ordered segments): identical vertex control points, identical set of
    mid-edge points, identical elements when compared through coordinates."""
    import metafem_b200 as m
    from metafem_jl_b200.frontend import mesh as fmesh
    from oracle import refgeom as rg, femmesh as fm
    size, n = (1.5, 1.0, 1.0), (3, 2, 2)
    coors, conn = rg.make_Brick(size, n, "CUBE")
    tm = rg.construct_TotalMesh_3D(coors, conn)
    omesh = fm.mesh_Classical(tm, [rg.get_BoundaryMesh(tm)], "CUBE")
    ctx = m.lib.Context(0)
    try:
        cp, x, f_el, f_eidx, cen = fmesh.second_order_tables_device(ctx, coors, conn)
    finally:
        ctx.close()
    nv = coors.shape[1]
    assert x.shape == omesh.x.shape
    assert np.array_equal(x[:, :nv], omesh.x[:, :nv])                      # vertex control points: input order on both sides
    assert np.array_equal(cp[:8], omesh.controlpoint_IDs[:8])            # ... and the same corner IDs in every element
    # element by element, local node by local node: same coordinates
    assert np.array_equal(x[:, cp - 1], omesh.x[:, omesh.controlpoint_IDs - 1])
    assert len(f_el) == len(omesh.facet_element_ID)
    assert sorted(zip(f_el.tolist(), f_eidx.tolist())) == sorted(zip(omesh.facet_element_ID.tolist(), omesh.facet_element_eindex.tolist()))
