"""construct_TotalMesh_3D's first-order geometry tables built on the device (mfb_total_mesh_build):
the REFERENCE numbering mode (FEM_Dict with sequential insertion) must equal the oracle's restatement bit for bit -- segments,
faces, block incidences, boundary faces --, the SORTED mode must describe the same mesh."""
import numpy as np
import pytest

from oracle import refgeom as rg

pytestmark = pytest.mark.gpu

CASES = [("CUBE", (4, 3, 2)), ("SIMPLEX", (3, 2, 2)), ("CUBE", (7, 5, 6)), ("SIMPLEX", (5, 4, 3))]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}-{'x'.join(map(str, c[1]))}")
def meshes(request, built_lib):
    import metafem_b200 as m
    from metafem_jl_b200.frontend import mesh as fmesh
    shape, n = request.param
    c, conn = rg.make_Brick((1.0, 0.9, 0.7), n, shape)
    om = rg.construct_TotalMesh_3D(c, conn)
    ctx = m.lib.Context(0)
    ref = fmesh.total_mesh_device(ctx, c.shape[1], conn, "reference")
    srt = fmesh.total_mesh_device(ctx, c.shape[1], conn, "sorted")
    ctx.close()
    return om, ref, srt, conn


def test_reference_numbering_is_bit_exact(meshes):
    om, ref, _, _ = meshes
    assert np.array_equal(ref["segment_vertex_IDs"], om.segment_vertex_IDs)
    assert np.array_equal(ref["block_segment_IDs"], om.block_segment_IDs)
    assert np.array_equal(ref["block_face_IDs"], om.block_face_IDs)
    assert np.array_equal(ref["face_segment_IDs"], om.face_segment_IDs)
    assert np.array_equal(ref["face_vertex_IDs"], om.face_vertex_IDs)
    assert np.array_equal(ref["boundary_face_IDs"], rg.get_BoundaryMesh(om))


def test_boundary_hosts(meshes):
    """every boundary face is face `eindex` of its host block (what specify_eindex finds)."""
    om, ref, _, _ = meshes
    f, b, e = ref["boundary_face_IDs"], ref["boundary_face_block"], ref["boundary_face_eindex"]
    assert len(f) > 0 and np.array_equal(om.block_face_IDs[e - 1, b - 1], f)


def test_sorted_numbering_describes_the_same_mesh(meshes):
    om, _, srt, conn = meshes
    ns, nf = om.segment_vertex_IDs.shape[1], om.face_vertex_IDs.shape[1]
    assert srt["segment_vertex_IDs"].shape == (2, ns) and srt["face_vertex_IDs"].shape[1] == nf
    seg = lambda t: {tuple(v) for v in t.T}
    assert seg(srt["segment_vertex_IDs"]) == seg(om.segment_vertex_IDs)              # same (max, next) pairs
    faces = lambda t: {tuple(sorted(v)) for v in t.T}
    assert faces(srt["face_vertex_IDs"]) == faces(om.face_vertex_IDs)
    # IDs are ranks of the sorted keys: segment table sorted by (max vertex, next vertex)
    sv = srt["segment_vertex_IDs"].astype(np.int64)
    key = sv[0] * (1 << 30) + sv[1]
    assert np.all(np.diff(key) > 0)
    # incidences are consistent: block b's local segment j joins the two vertices the topology table says
    bsv = np.array(rg.B_S_V["CUBE" if conn.shape[0] == 8 else "SIMPLEX"]) - 1
    for j, (a, b) in enumerate(bsv):
        got = np.sort(srt["segment_vertex_IDs"][:, srt["block_segment_IDs"][j] - 1], axis=0)
        assert np.array_equal(got, np.sort(conn[[a, b]], axis=0))
    assert len(srt["boundary_face_IDs"]) == len(rg.get_BoundaryMesh(om))
