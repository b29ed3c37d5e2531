"""ORACLE: first-order geometry (reference src/mesh/ref_geometry/*), numpy restatement.

All ID arrays hold 1-based IDs exactly as the reference stores them; array storage is 0-based,
so ID k lives at index k-1. Tables are modelled without holes (the reference tables have none for
any in-scope script: allocate_by_length! hands out the first free slots, 05_GPU_Table.jl:54-62).
"""
import re
import numpy as np

from .femdict import FemDict, I4I30I30_To_UI64

# 002_Initialization.jl:1-8
B_S_V = {"SIMPLEX": [[1, 2], [2, 3], [3, 1], [1, 4], [2, 4], [3, 4]],
         "CUBE": [[1, 2], [2, 3], [3, 4], [4, 1], [1, 5], [2, 6], [3, 7], [4, 8], [5, 6], [6, 7], [7, 8], [8, 5]]}
B_F_S = {"SIMPLEX": [[1, 2, 3], [1, 5, 4], [2, 6, 5], [3, 4, 6]],
         "CUBE": [[1, 2, 3, 4], [1, 6, 9, 5], [2, 7, 10, 6], [3, 8, 11, 7], [4, 8, 12, 5], [9, 10, 11, 12]]}


def make_Brick(x, n, shape="CUBE"):
    """201_Helper_TM.jl:36-78. Returns coors (3, nv) float64 and connections (8|4, nel) int32, 1-based."""
    nx, ny, nz = n
    dx = [x[0] / nx, x[1] / ny, x[2] / nz]
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    coors = np.stack([dx[0] * I.ravel(), dx[1] * J.ravel(), dx[2] * K.ravel()]).astype(np.float64)
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    s1, s2 = (ny + 1) * (nz + 1), (nz + 1)
    cc = np.stack([(i - 1) * s1 + (j - 1) * s2 + k, i * s1 + (j - 1) * s2 + k, i * s1 + j * s2 + k,
                   (i - 1) * s1 + j * s2 + k, (i - 1) * s1 + (j - 1) * s2 + k + 1, i * s1 + (j - 1) * s2 + k + 1,
                   i * s1 + j * s2 + k + 1, (i - 1) * s1 + j * s2 + k + 1]).astype(np.int32)
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


class TotalMesh3D:
    """Geo_TotalMesh3D (001_Types.jl:37-42) with the fields the path reads."""
    pass


def construct_TotalMesh_3D(coors, connections):
    """002_Initialization.jl:113-217 (segments :137-165, faces :167-215)."""
    connections = np.asarray(connections, dtype=np.int32)
    vpb, nb = connections.shape
    mesh_type = {4: "SIMPLEX", 8: "CUBE"}[vpb]
    vpf = 3 if mesh_type == "SIMPLEX" else 4
    bsv, bfs = B_S_V[mesh_type], B_F_S[mesh_type]
    m = TotalMesh3D()
    m.mesh_type = mesh_type
    m.x = np.asarray(coors, dtype=np.float64).copy()          # (3, nv)
    m.block_vertex_IDs = connections.copy()
    m.block_segment_IDs = np.zeros((len(bsv), nb), np.int32)
    m.block_face_IDs = np.zeros((len(bfs), nb), np.int32)
    cols = np.arange(nb)

    seg_dict = FemDict()
    seg_v = np.zeros((2, 0), np.int32)
    for b_s_pos, s_v_pos in enumerate(bsv):
        s_vIDs = connections[np.array(s_v_pos) - 1]           # (2, nb)
        max_pos = np.argmax(s_vIDs, axis=0)                    # first max, like findmax
        max_vIDs = s_vIDs[max_pos, cols]
        next_pos = (max_pos + 1) % 2
        next_vIDs = s_vIDs[next_pos, cols]
        keys = I4I30I30_To_UI64(0, max_vIDs, next_vIDs)
        slots = seg_dict.set_ids(keys)
        tot = seg_dict.total_ids()
        na = seg_dict.vals[tot - 1] == 0
        n_new = int(na.sum())
        new_ids = np.arange(seg_v.shape[1] + 1, seg_v.shape[1] + n_new + 1, dtype=np.int32)
        seg_dict.vals[tot[na] - 1] = new_ids
        seg_v = np.concatenate([seg_v, np.zeros((2, n_new), np.int32)], axis=1)
        local = seg_dict.vals[slots - 1]
        m.block_segment_IDs[b_s_pos] = local
        seg_v[0, local - 1] = max_vIDs
        seg_v[1, local - 1] = next_vIDs
    m.segment_vertex_IDs = seg_v

    fac_dict = FemDict()
    f_v = np.zeros((vpf, 0), np.int32)
    f_s = np.zeros((vpf, 0), np.int32)
    for b_f_pos, f_s_pos in enumerate(bfs):
        f_sIDs = m.block_segment_IDs[np.array(f_s_pos) - 1]   # (vpf, nb)
        rot = len(f_s_pos)
        max_pos = np.argmax(f_sIDs, axis=0)
        max_sIDs = f_sIDs[max_pos, cols]
        prev_pos = (max_pos - 1) % rot
        next_pos = (max_pos + 1) % rot
        is_forward = f_sIDs[next_pos, cols] >= f_sIDs[prev_pos, cols]
        next_pos = np.where(is_forward, next_pos, prev_pos)
        next_sIDs = f_sIDs[next_pos, cols]
        keys = I4I30I30_To_UI64(0, max_sIDs, next_sIDs)
        slots = fac_dict.set_ids(keys)
        tot = fac_dict.total_ids()
        na = fac_dict.vals[tot - 1] == 0
        n_new = int(na.sum())
        new_ids = np.arange(f_v.shape[1] + 1, f_v.shape[1] + n_new + 1, dtype=np.int32)
        fac_dict.vals[tot[na] - 1] = new_ids
        f_v = np.concatenate([f_v, np.zeros((vpf, n_new), np.int32)], axis=1)
        f_s = np.concatenate([f_s, np.zeros((vpf, n_new), np.int32)], axis=1)
        local = fac_dict.vals[slots - 1]
        m.block_face_IDs[b_f_pos] = local
        f_s[0, local - 1] = max_sIDs
        last_sIDs, last_pos = max_sIDs, max_pos
        for i in range(vpf):
            a = seg_v[0, last_sIDs - 1]
            is_first = (a == seg_v[0, next_sIDs - 1]) | (a == seg_v[1, next_sIDs - 1])
            f_v[i, local[is_first] - 1] = seg_v[0, last_sIDs[is_first] - 1]
            f_v[i, local[~is_first] - 1] = seg_v[1, last_sIDs[~is_first] - 1]
            if i == vpf - 1:
                break
            f_s[i + 1, local - 1] = next_sIDs
            last_sIDs, last_pos = next_sIDs, next_pos
            prev_pos = (last_pos - 1) % rot
            nxt = (last_pos + 1) % rot
            next_pos = np.where(is_forward, nxt, prev_pos)
            next_sIDs = f_sIDs[next_pos, cols]
    m.face_vertex_IDs, m.face_segment_IDs = f_v, f_s
    return m


def get_BoundaryMesh(m):
    """002_Initialization.jl:285-290: faces referenced by exactly one block (1-based IDs, ascending)."""
    cnt = np.bincount(m.block_face_IDs.ravel(), minlength=m.face_vertex_IDs.shape[1] + 1)
    return np.nonzero(cnt[1:] == 1)[0].astype(np.int32) + 1


def face_centroids(m, face_IDs):
    """Mean of the face's vertices, as every example script computes it (e.g. static_Neo_Hookean.jl:21-26)."""
    v = m.face_vertex_IDs[:, np.asarray(face_IDs) - 1] - 1
    return np.stack([m.x[d][v].sum(axis=0) / v.shape[0] for d in range(3)])


def read_MPHTXT(path):
    """102_Read_MPHTXT.jl:4-42 (first element block after the points; 0-based file IDs -> 1-based)."""
    lines = [l.strip() for l in open(path) if l.strip() and not l.startswith("#")]
    coors = conn = None
    start_vid = 0
    i = 0
    while i < len(lines):
        t = lines[i].split(" ")
        if len(t) >= 6 and t[2:6] == ["number", "of", "mesh", "points"]:
            nv = int(t[0])
            start_vid = int(lines[i + 1].split(" ")[0])
            coors = np.array([[float(s) for s in lines[i + 2 + r].split()] for r in range(nv)]).T
            i += 2 + nv
            continue
        if len(t) >= 5 and t[2:5] == ["number", "of", "elements"] and conn is None and coors is not None:
            ne = int(t[0])
            conn = np.array([[int(s) for s in lines[i + 1 + r].split()] for r in range(ne)], dtype=np.int32).T
            break
        i += 1
    return coors, (conn - (start_vid - 1)).astype(np.int32)


def read_INP(path):
    """101_Read_INP.jl:13-54: *NODE and *ELEMENT blocks, node labels compacted to 1..n."""
    vids, coors, els = [], [], []
    mode = None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("**"):
            continue
        if line.startswith("*"):
            key = line.split(",")[0].strip().upper()
            mode = {"*NODE": "n", "*ELEMENT": "e"}.get(key)
            if vids and els and mode is None:
                break
            continue
        if not line.strip():
            mode = None
            continue
        t = [s for s in re.split(r",\s*", line.strip()) if s != ""]
        if mode == "n":
            vids.append(int(t[0])); coors.append([float(s) for s in t[1:]])
        elif mode == "e":
            els.append([int(s) for s in t[1:]])
    vids = np.array(vids)
    local = np.zeros(vids.max() + 1, np.int32)
    local[vids] = np.arange(1, len(vids) + 1)
    return np.array(coors).T, local[np.array(els, dtype=np.int64).T].astype(np.int32)
