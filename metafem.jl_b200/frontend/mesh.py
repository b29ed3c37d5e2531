"""Synthetic-mesh front end for benchmarks: make_Brick + second-order node placement + boundary facets.

Stand-in for make_Brick / construct_TotalMesh / mesh_Classical (reference src/mesh/ref_geometry/201_Helper_TM.jl:36-78,
002_Initialization.jl:113-217, src/mesh/unstructured_mesh/3_InitializeMesh.jl:70-178). Vertex, element and
vertex-node numbering follow the reference exactly (they are input-order based). Mid-edge node IDs are
assigned by the sorted (max vertex, min vertex) key instead of the reference's hash-slot order;
``numbering="scattered"`` applies a seeded pseudo-random permutation to them, which reproduces the
locality of the reference's hash order for honest SpMV/assembly timings.
"""
import numpy as np

from . import elements
from ..api import MeshTables


def make_Brick(x, n, shape="CUBE"):
    """201_Helper_TM.jl:36-78."""
    nx, ny, nz = n
    dx = [x[0] / nx, x[1] / ny, x[2] / nz]
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    coors = np.stack([dx[0] * I.ravel(), dx[1] * J.ravel(), dx[2] * K.ravel()])
    i, j, k = [a.ravel() for a in np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")]
    s1, s2 = (ny + 1) * (nz + 1), nz + 1
    v = lambda di, dj, dk: (i - 1 + di) * s1 + (j - 1 + dj) * s2 + k + dk
    cc = np.stack([v(0, 0, 0), v(1, 0, 0), v(1, 1, 0), v(0, 1, 0), v(0, 0, 1), v(1, 0, 1), v(1, 1, 1), v(0, 1, 1)]).astype(np.int32)
    if shape == "CUBE":
        return coors, cc
    ne = nx * ny * nz
    conn = np.zeros((4, 5 * ne), np.int32)
    odd = ((i + j + k) % 2 == 1)
    fw, bw = np.nonzero(odd)[0], np.nonzero(~odd)[0]
    fsel = [[1, 2, 4, 5], [3, 4, 2, 7], [8, 7, 5, 4], [6, 5, 7, 2], [4, 7, 5, 2]]
    bsel = [[5, 8, 6, 1], [2, 1, 6, 3], [7, 6, 8, 3], [4, 1, 3, 8], [1, 3, 8, 6]]
    for d in range(5):
        conn[:, fw + d * ne] = cc[np.array(fsel[d]) - 1][:, fw]
        conn[:, bw + d * ne] = cc[np.array(bsel[d]) - 1][:, bw]
    return coors, conn


def second_order_tables(coors, connections, boundary_selectors, numbering="sorted", seed=1234, itg_order=5):
    """Build MeshTables for hex20 / tet10 from a first-order mesh.

    boundary_selectors: list of callables centroid(3, n_faces) -> bool mask; group k+1 gets the boundary faces
    selected by selector k (faces must not be shared between groups).
    """
    connections = np.asarray(connections, dtype=np.int64)
    vpb, nb = connections.shape
    et = elements.hex20_tables(itg_order) if vpb == 8 else elements.tet10_tables()
    nv = coors.shape[1]
    # ---- mid-edge nodes -------------------------------------------------------------------
    segs = np.array(et.segment_vertices) - 1                                   # (ns, 2)
    va, vb = connections[segs[:, 0]], connections[segs[:, 1]]                    # (ns, nb)
    key = np.maximum(va, vb) * (nv + 1) + np.minimum(va, vb)
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    n_edges = len(uniq)
    edge_ids = np.arange(n_edges, dtype=np.int64)
    if numbering == "scattered":
        edge_ids = np.random.default_rng(seed).permutation(n_edges)
    elif numbering != "sorted":
        raise ValueError(numbering)
    cp = np.zeros((et.n_a, nb), np.int32)
    for j, loc in enumerate(et.vertex_cp_ids):
        cp[loc - 1] = connections[j]
    eid = edge_ids[inv].reshape(key.shape)
    for j, loc in enumerate(et.segment_cp_ids):
        cp[loc - 1] = nv + eid[j] + 1
    x = np.zeros((3, nv + n_edges))
    x[:, :nv] = coors
    hi, lo = uniq // (nv + 1) - 1, uniq % (nv + 1) - 1
    x[:, nv + edge_ids] = 0.5 * coors[:, hi] + 0.5 * coors[:, lo]
    # ---- boundary facets ------------------------------------------------------------------
    fvs = np.array(et.face_vertices) - 1                                          # (nf, vpf)
    fv = connections[fvs]                                                         # (nf, vpf, nb)
    fkey_parts = np.sort(fv, axis=1)
    fkey = np.zeros(fkey_parts.shape[0::2], dtype=np.int64)
    for c in range(fkey_parts.shape[1]):
        fkey = fkey * (nv + 1) + fkey_parts[:, c, :]
    if fkey_parts.shape[1] == 4:      # 4 x 21 bits would overflow for huge meshes: drop the largest vertex (3 identify a quad)
        fkey = np.zeros_like(fkey)
        for c in range(3):
            fkey = fkey * (nv + 1) + fkey_parts[:, c, :]
    _, finv, cnt = np.unique(fkey.ravel(), return_inverse=True, return_counts=True)
    on_bdy = (cnt[finv] == 1).reshape(fkey.shape)                                 # (nf, nb)
    f_idx, e_idx = np.nonzero(on_bdy)
    cen = coors[:, fv[f_idx, :, e_idx] - 1].mean(axis=2)                          # (3, n_bfaces)
    facet_el, facet_eidx, bg = [], [], {}
    start = 0
    for g, sel in enumerate(boundary_selectors, start=1):
        m = sel(cen)
        k = int(m.sum())
        facet_el.append(e_idx[m] + 1)
        facet_eidx.append(f_idx[m] + 1)
        bg[g] = np.arange(start + 1, start + k + 1, dtype=np.int32)
        start += k
    facet_el = np.concatenate(facet_el) if facet_el else np.zeros(0, np.int32)
    facet_eidx = np.concatenate(facet_eidx) if facet_eidx else np.zeros(0, np.int32)
    return MeshTables(controlpoint_IDs=cp, x=x, ref_itp_vals=et.ref_itp_vals, itg_weight=et.itg_weight,
                      bdy_ref_itp_vals=et.bdy_ref_itp_vals, bdy_itg_weights=et.bdy_itg_weights,
                      bdy_tangent_directions=et.bdy_tangent_directions, facet_element_ID=facet_el,
                      facet_element_eindex=facet_eidx, bg_fIDs=bg)


def box_tables(size, n, shape="CUBE", groups=("left", "right"), numbering="sorted", seed=1234):
    """Structured box + boundary groups named by side: left/right (x), front/back (y), bottom/top (z), all."""
    coors, conn = make_Brick(size, n, shape)
    eps = 1e-6 * max(size)
    side = {"left": (0, 0.0), "right": (0, size[0]), "front": (1, 0.0), "back": (1, size[1]),
            "bottom": (2, 0.0), "top": (2, size[2])}

    def selector(name):
        if name == "all":
            return lambda c: np.ones(c.shape[1], bool)
        d, v = side[name]
        return lambda c: np.abs(c[d] - v) < eps

    return second_order_tables(coors, conn, [selector(g) for g in groups], numbering=numbering, seed=seed)


def second_order_tables_device(ctx, coors, connections, itg_order=5):
    """The same tables as ``second_order_tables(..., numbering="sorted")`` built ON THE DEVICE by
    ``mfb_mesh_build_second_order`` (segments and boundary faces through radix sorts instead of the reference's GPU hash).
    Returns (controlpoint_IDs, x, bfacet_element_ID, bfacet_element_eindex, bfacet_centroids); the script then selects its
    boundary groups on the centroids exactly as it does with get_BoundaryMesh (static_Neo_Hookean.jl:19-34)."""
    import ctypes as C
    from .. import lib as L
    connections = np.ascontiguousarray(np.asarray(connections, dtype=np.int32).T)          # [n_el][vpb] == column-major [vpb, n_el]
    n_el, vpb = connections.shape
    et = elements.hex20_tables(itg_order) if vpb == 8 else elements.tet10_tables()
    nv = coors.shape[1]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    seg, vcp, scp, fv = i32(et.segment_vertices), i32(et.vertex_cp_ids), i32(et.segment_cp_ids), i32(et.face_vertices)
    x = [np.ascontiguousarray(coors[d], dtype=np.float64) for d in range(3)]
    N, nbf = C.c_int64(0), C.c_int64(0)
    ctx.call("mfb_mesh_build_second_order", nv, L.ptr(x[0]), L.ptr(x[1]), L.ptr(x[2]), vpb, n_el, L.ptr(connections), len(seg),
             L.ptr(seg), L.ptr(vcp), L.ptr(scp), fv.shape[0], fv.shape[1], L.ptr(fv), C.byref(N), C.byref(nbf))
    cp = np.empty((n_el, et.n_a), np.int32)
    xo = np.empty((3, N.value))
    f_el, f_eidx, cen = np.empty(nbf.value, np.int32), np.empty(nbf.value, np.int32), np.empty((nbf.value, 3))
    ctx.call("mfb_mesh_build_get", L.ptr(cp), L.ptr(xo[0]), L.ptr(xo[1]), L.ptr(xo[2]), L.ptr(f_el), L.ptr(f_eidx), L.ptr(cen))
    return np.asfortranarray(cp.T), xo, f_el, f_eidx, cen.T


def total_mesh_device(ctx, n_vert, connections, numbering="sorted"):
    """construct_TotalMesh_3D's tables (002_Initialization.jl:113-217) built ON THE DEVICE by ``mfb_total_mesh_build``:
    segment / face tables, block incidences, boundary faces with their host block and local face number.
    numbering: "sorted" (rank of the key; deterministic) or "reference" (the reference's hash-slot order, sequential insertion).
    Returns a dict of 1-based int32 arrays shaped like the reference's column-major tables."""
    import ctypes as C
    from .. import lib as L
    conn = np.ascontiguousarray(np.asarray(connections, dtype=np.int32).T)          # [n_blocks][vpb] == column-major [vpb, n_blocks]
    nb, vpb = conn.shape
    ns, nf, nbf = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    mode = {"sorted": L.NUMBERING_SORTED, "reference": L.NUMBERING_REFERENCE}[numbering]
    ctx.call("mfb_total_mesh_build", int(n_vert), vpb, nb, L.ptr(conn), mode, C.byref(ns), C.byref(nf), C.byref(nbf))
    nsb, nfb, vpf = (6, 4, 3) if vpb == 4 else (12, 6, 4)
    out = dict(segment_vertex_IDs=np.empty((ns.value, 2), np.int32), block_segment_IDs=np.empty((nb, nsb), np.int32),
               face_vertex_IDs=np.empty((nf.value, vpf), np.int32), face_segment_IDs=np.empty((nf.value, vpf), np.int32),
               block_face_IDs=np.empty((nb, nfb), np.int32), boundary_face_IDs=np.empty(nbf.value, np.int32),
               boundary_face_block=np.empty(nbf.value, np.int32), boundary_face_eindex=np.empty(nbf.value, np.int32))
    ctx.call("mfb_total_mesh_get", *[L.ptr(out[k]) for k in ("segment_vertex_IDs", "block_segment_IDs", "face_vertex_IDs",
                                                            "face_segment_IDs", "block_face_IDs", "boundary_face_IDs",
                                                            "boundary_face_block", "boundary_face_eindex")])
    for k in ("segment_vertex_IDs", "block_segment_IDs", "face_vertex_IDs", "face_segment_IDs", "block_face_IDs"):
        out[k] = out[k].T                                                              # the reference's [rows, n] orientation
    return out
