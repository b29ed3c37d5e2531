"""ORACLE: mesh_Classical node placement + update_Mesh geometry tables.

reference: src/mesh/unstructured_mesh/2_Interface.jl:7-39,98-108 (mesh_Classical, update_Mesh)
           src/mesh/unstructured_mesh/3_InitializeMesh.jl:70-178 (allocate_Basic_WP_Mesh_3D, specify_eindex)
           src/mesh/unstructured_mesh/4_Update_Integrator.jl:2-75,90-157,173-227 (Jacobians, itg vals, normals)
IDs are 1-based values in 0-based numpy storage. Arrays are stored element-major for numpy convenience:
  integral_vals[e, s, a, q]  with s = 0:N, 1:d/dx1, 2:d/dx2, 3:d/dx3   (reference: [q, a, sd1, sd2, sd3, e])
  integral_weights[e, q]                                                   (reference: [q, e])
"""
import numpy as np

from .discretization import initialize_Classical_Element

SLOT = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]   # sd_IDs - 1 of the four slots in use


class WPMesh:
    pass


def mesh_Classical(ref_mesh, boundarys, shape, itp_type="Serendipity", itp_order=2, itg_order=5, max_sd_order=1):
    """Returns WPMesh with controlpoints (x: (3,N)), elements.controlpoint_IDs (n_a, n_el), facets."""
    sp = initialize_Classical_Element(3, shape, itp_order, max_sd_order, itg_order, itp_type)
    st = sp.element_structure
    mesh = WPMesh()
    mesh.space = sp
    nv = ref_mesh.x.shape[1]
    nb = ref_mesh.block_vertex_IDs.shape[1]
    ns = ref_mesh.segment_vertex_IDs.shape[1]
    cps = st.segment_cp_ids.shape[0]
    n_a = sp.itp_func_num
    cp_ids = np.zeros((n_a, nb), np.int32)
    # vertices -> control points 1..nv (3_InitializeMesh.jl:79-93), cp_per_vertex == 1
    for j in range(st.vertex_cp_ids.shape[1]):
        cp_ids[st.vertex_cp_ids[0, j] - 1] = ref_mesh.block_vertex_IDs[j]
    x = np.zeros((3, nv + cps * ns))
    x[:, :nv] = ref_mesh.x
    # segment control points (:95-117)
    sv = ref_mesh.segment_vertex_IDs - 1
    for i in range(1, cps + 1):
        batch = nv + np.arange(ns) * cps + (i - 1)
        for d in range(3):
            x[d, batch] = (st.segment_cp_pos[0, i - 1] * ref_mesh.x[d][sv[0]]
                           + st.segment_cp_pos[1, i - 1] * ref_mesh.x[d][sv[1]])
    for j in range(st.segment_cp_ids.shape[1]):
        sid = ref_mesh.block_segment_IDs[j]
        aligned = ref_mesh.segment_vertex_IDs[0, sid - 1] == ref_mesh.block_vertex_IDs[st.segment_start_vertex[j] - 1]
        for i in range(1, cps + 1):
            a_id = nv + (sid - 1) * cps + i
            n_id = nv + sid * cps - (i - 1)
            cp_ids[st.segment_cp_ids[i - 1, j] - 1] = np.where(aligned, a_id, n_id)
    assert st.face_cp_ids.shape[0] == 0 and len(st.block_cp_ids) == 0
    mesh.x = x
    mesh.controlpoint_IDs = cp_ids
    mesh.variable_size = x.shape[1]
    # facets (:119-130, kernel :165-178)
    nf_total = ref_mesh.face_vertex_IDs.shape[1]
    face_facet = np.zeros(nf_total, np.int32)
    mesh.bg_fIDs = {}
    nfac = 0
    for bg_index, boundary in enumerate(boundarys, start=1):
        boundary = np.asarray(boundary, dtype=np.int64)
        ids = np.arange(nfac + 1, nfac + len(boundary) + 1, dtype=np.int32)
        assert np.all(face_facet[boundary - 1] == 0), "a face belongs to two boundary groups (reference mis-handles it)"
        mesh.bg_fIDs[bg_index] = ids
        face_facet[boundary - 1] = ids
        nfac += len(boundary)
    el_ID = np.zeros(nfac, np.int32)
    eindex = np.zeros(nfac, np.int32)
    for j in range(ref_mesh.block_face_IDs.shape[0]):
        fac = face_facet[ref_mesh.block_face_IDs[j] - 1]
        for e in np.nonzero(fac)[0]:                     # element order = sequential CAS winner
            f = fac[e] - 1
            if el_ID[f] == 0:
                el_ID[f] = e + 1
                eindex[f] = j + 1
    mesh.facet_element_ID, mesh.facet_element_eindex = el_ID, eindex
    return mesh


def _inv_Jac_3D(J):
    """4_Update_Integrator.jl:90-121. J[..., i, X] = d x_i / d X. Returns det, inv with inv[..., m, s]."""
    j = lambda a, b: J[..., a - 1, b - 1]
    det = (j(1, 1) * j(2, 2) * j(3, 3) - j(1, 1) * j(2, 3) * j(3, 2) - j(1, 2) * j(2, 1) * j(3, 3)
           + j(1, 2) * j(2, 3) * j(3, 1) + j(1, 3) * j(2, 1) * j(3, 2) - j(1, 3) * j(2, 2) * j(3, 1))
    inv = np.empty_like(J)
    inv[..., 0, 0] = (j(2, 2) * j(3, 3) - j(2, 3) * j(3, 2)) / det
    inv[..., 0, 1] = (j(1, 3) * j(3, 2) - j(1, 2) * j(3, 3)) / det
    inv[..., 0, 2] = (j(1, 2) * j(2, 3) - j(2, 2) * j(1, 3)) / det
    inv[..., 1, 0] = (j(2, 3) * j(3, 1) - j(3, 3) * j(2, 1)) / det
    inv[..., 1, 1] = (j(1, 1) * j(3, 3) - j(1, 3) * j(3, 1)) / det
    inv[..., 1, 2] = (j(1, 3) * j(2, 1) - j(1, 1) * j(2, 3)) / det
    inv[..., 2, 0] = (j(2, 1) * j(3, 2) - j(2, 2) * j(3, 1)) / det
    inv[..., 2, 1] = (j(1, 2) * j(3, 1) - j(3, 2) * j(1, 1)) / det
    inv[..., 2, 2] = (j(1, 1) * j(2, 2) - j(2, 1) * j(1, 2)) / det
    return det, inv


def _ref4(ref):
    """(q, a, 2,2,2) -> (4, a, q): N, dX1, dX2, dX3."""
    return np.stack([ref[:, :, s[0], s[1], s[2]].T for s in SLOT])


def _geometry(ref, xe):
    """xe: (n, a, 3) node coordinates. Returns J (n,q,3,3), det (n,q), inv, itg_vals (n,4,a,q)."""
    r4 = _ref4(ref)                                   # (4, a, q)
    J = np.einsum("Xaq,nai->nqiX", r4[1:], xe)        # jacobian[i, X, q, e], :9
    det, inv = _inv_Jac_3D(J)
    iv = np.empty((xe.shape[0], 4) + r4.shape[1:])
    iv[:, 0] = r4[0]
    # update_Basic_itgval_1_3D (:125-154): vals[sd] = ((ref[X1]*inv[1,sd] + ref[X2]*inv[2,sd]) + ref[X3]*inv[3,sd])
    for s in range(3):
        iv[:, 1 + s] = (r4[1][None] * inv[:, None, :, 0, s] + r4[2][None] * inv[:, None, :, 1, s]) \
            + r4[3][None] * inv[:, None, :, 2, s]
    return J, det, inv, iv


def update_Mesh(mesh):
    """update_BasicElements_3D + update_BasicBoundary_3D."""
    sp = mesh.space
    cp = mesh.controlpoint_IDs - 1                   # (a, e)
    xe = mesh.x.T[cp.T]                               # (e, a, 3)
    _, det, _, iv = _geometry(sp.ref_itp_vals, xe)
    mesh.integral_vals = iv
    mesh.integral_weights = sp.itg_weight[None, :] * det          # :30 (no abs)
    nfac = len(mesh.facet_element_ID)
    nqb = sp.bdy_itg_func_num
    mesh.facet_integral_vals = np.zeros((nfac, 4, sp.itp_func_num, nqb))
    mesh.facet_integral_weights = np.zeros((nfac, nqb))
    mesh.facet_normal_directions = np.zeros((nfac, 3, nqb))
    for eidx in range(1, len(sp.bdy_ref_itp_vals) + 1):
        f = np.nonzero(mesh.facet_element_eindex == eidx)[0]
        if len(f) == 0:
            continue
        xe = mesh.x.T[cp.T[mesh.facet_element_ID[f] - 1]]
        J, _, _, iv = _geometry(sp.bdy_ref_itp_vals[eidx - 1], xe)
        bt = sp.bdy_tangent_directions[eidx - 1]      # (q, 3, 2)
        # update_Basic_Tangent_3D :173-199
        t = np.empty((len(f), nqb, 3, 2))
        for k in range(2):
            for i in range(3):
                t[:, :, i, k] = (J[:, :, i, 0] * bt[None, :, 0, k] + J[:, :, i, 1] * bt[None, :, 1, k]) \
                    + J[:, :, i, 2] * bt[None, :, 2, k]
        # update_Basic_Normal_3D :210-227
        rn1 = t[..., 1, 0] * t[..., 2, 1] - t[..., 2, 0] * t[..., 1, 1]
        rn2 = -t[..., 0, 0] * t[..., 2, 1] + t[..., 2, 0] * t[..., 0, 1]
        rn3 = t[..., 0, 0] * t[..., 1, 1] - t[..., 1, 0] * t[..., 0, 1]
        ld = np.sqrt(rn1 ** 2. + rn2 ** 2. + rn3 ** 2.)
        mesh.facet_normal_directions[f] = np.stack([rn1 / ld, rn2 / ld, rn3 / ld], axis=1)
        mesh.facet_integral_vals[f] = iv
        mesh.facet_integral_weights[f] = sp.bdy_itg_weights[eidx - 1][None, :] * ld   # :71
    return mesh
