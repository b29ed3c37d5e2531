"""Reference-element tables in closed form (hex20 serendipity on [0,1]^3, tet10 Lagrange).

The production front end is MetaFEM's Julia code (src/mesh/spatial_discretization/*), which builds
these tables with its polynomial algebra; this module is the stand-in used by the benchmarks and
by tests that must not touch the oracle. Conventions follow the reference:
local node order (101_Structures.jl:129-196,224-247; 102_Interpolations.jl:46-62,69-113),
Gauss points with the first coordinate fastest (103_Integrations.jl:14-19), tet rule of degree 5
(:70-78,145-206), face numbering and tangent frames (:21-58, 208-238).
Outputs use the reference's array shapes: ref_itp_vals (n_q, n_a, 2, 2, 2) etc.
"""
import itertools
import math

import numpy as np


def _prod_first_fastest(ranges):
    for t in itertools.product(*reversed(ranges)):
        yield tuple(reversed(t))


# ---- hex20 -------------------------------------------------------------------------------------
def hex20_nodes():
    nodes = [tuple(float(c) for c in cs) for cs in _prod_first_fastest([range(2)] * 3)]
    for e in range(3):
        minor = [d for d in range(3) if d != e]
        for mc in _prod_first_fastest([range(2)] * 2):
            p = [0.5] * 3
            p[minor[0]], p[minor[1]] = float(mc[0]), float(mc[1])
            nodes.append(tuple(p))
    return nodes


def hex20_eval(pts):
    """N (n_pts, 20) and dN (n_pts, 20, 3) at points of [0,1]^3."""
    pts = np.asarray(pts, dtype=np.float64)
    n = len(pts)
    N = np.zeros((n, 20))
    dN = np.zeros((n, 20, 3))
    x = pts
    a = 0
    for cs in _prod_first_fastest([range(2)] * 3):
        Lf = [x[:, d] if cs[d] else 1 - x[:, d] for d in range(3)]
        dL = [1.0 if cs[d] else -1.0 for d in range(3)]
        s = [1 - 2 * c for c in cs]
        f = 1 - 2 * sum(cs) - 2 * sum(s[d] * x[:, d] for d in range(3))
        P = Lf[0] * Lf[1] * Lf[2]
        N[:, a] = P * f
        for d in range(3):
            others = [Lf[k] for k in range(3) if k != d]
            dN[:, a, d] = dL[d] * others[0] * others[1] * f + P * (-2 * s[d])
        a += 1
    for e in range(3):
        minor = [d for d in range(3) if d != e]
        for mc in _prod_first_fastest([range(2)] * 2):
            L1 = x[:, minor[0]] if mc[0] else 1 - x[:, minor[0]]
            L2 = x[:, minor[1]] if mc[1] else 1 - x[:, minor[1]]
            d1 = 1.0 if mc[0] else -1.0
            d2 = 1.0 if mc[1] else -1.0
            b = 4 * x[:, e] * (1 - x[:, e])
            N[:, a] = b * L1 * L2
            dN[:, a, e] = 4 * (1 - 2 * x[:, e]) * L1 * L2
            dN[:, a, minor[0]] = b * d1 * L2
            dN[:, a, minor[1]] = b * L1 * d2
            a += 1
    return N, dN


# ---- tet10 -------------------------------------------------------------------------------------
def tet10_lattice():
    return [ijk for ijk in _prod_first_fastest([range(3)] * 3) if sum(ijk) <= 2]


def tet10_eval(pts):
    pts = np.asarray(pts, dtype=np.float64)
    lam = np.stack([pts[:, 0], pts[:, 1], pts[:, 2], 1 - pts.sum(axis=1)], axis=1)      # (n, 4)
    dlam = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, -1]], dtype=np.float64)   # (4, 3)

    def P(m, l):
        return (np.ones_like(l), 2 * l, l * (2 * l - 1))[m]

    def dP(m, l):
        return (np.zeros_like(l), 2 * np.ones_like(l), 4 * l - 1)[m]

    lat = tet10_lattice()
    N = np.zeros((len(pts), 10))
    dN = np.zeros((len(pts), 10, 3))
    for a, (i, j, k) in enumerate(lat):
        ms = (i, j, k, 2 - i - j - k)
        vals = [P(m, lam[:, b]) for b, m in enumerate(ms)]
        N[:, a] = vals[0] * vals[1] * vals[2] * vals[3]
        for b, m in enumerate(ms):
            rest = np.ones(len(pts))
            for c in range(4):
                if c != b:
                    rest = rest * vals[c]
            g = dP(m, lam[:, b]) * rest
            dN[:, a, :] += g[:, None] * dlam[b][None, :]
    return N, dN


# ---- quadrature --------------------------------------------------------------------------------
def gauss_cube(itg_order, dim):
    go = int(math.ceil((itg_order + 1) / 2))
    xg, wg = np.polynomial.legendre.leggauss(go)
    xg, wg = xg / 2.0 + 0.5, wg / 2.0
    pts = [tuple(xg[i] for i in ids) for ids in _prod_first_fastest([range(go)] * dim)]
    w = [float(np.prod([wg[i] for i in ids])) for ids in _prod_first_fastest([range(go)] * dim)]
    return np.array(pts), np.array(w)


_TRI5 = ((0.10128650732345633880098736191512383, 0.12593918054482715259568394550018133),
         (0.47014206410511508977044120951344760, 0.13239415278850618073764938783315200))
_TET5 = ((0.31088591926330060979734573376345783, 0.11268792571801585079918565233328633),
         (0.09273525031089122640232391373703061, 0.07349304311636194954371020548632750),
         (-0.04550370412564964949188052627933943, 0.04254602077708146643806942812025744))


def gauss_triangle_bary():
    """Degree-5 rule, 7 points, barycentric triples, weights summing to 1/... (reference: unit sum * 1)."""
    pts, w = [], []
    for a, wt in _TRI5:
        for i in range(3):
            p = [a] * 3
            p[i] = 1 - 2 * a
            pts.append(tuple(p)); w.append(wt)
    pts.append((1 / 3,) * 3); w.append(9.0 / 40.0)
    return np.array(pts), np.array(w)


def gauss_tet():
    """Degree-5 rule, 14 points on the unit tetrahedron (x, y, z) = barycentric 2..4; weights sum to 1/6."""
    pts, w = [], []
    for a, wt in _TET5:
        if a >= 0:
            for i in range(4):
                p = [a] * 4
                p[i] = 1 - 3 * a
                pts.append(tuple(p)); w.append(wt)
        else:
            b = -a
            for (i, j) in _prod_first_fastest([range(4)] * 2):
                if i >= j:
                    continue
                p = [b] * 4
                p[i] = p[j] = 0.5 - b
                pts.append(tuple(p)); w.append(wt)
    pts = np.array(pts)
    return pts[:, 1:4], np.array(w) / 6


# ---- tables in the reference's shapes ----------------------------------------------------------
def _pack(N, dN):
    """-> (n_q, n_a, 2, 2, 2) with slot (1,0,0)=d/dX1, (0,1,0)=d/dX2, (0,0,1)=d/dX3; mixed slots unused (0)."""
    out = np.zeros(N.shape + (2, 2, 2))
    out[:, :, 0, 0, 0] = N
    out[:, :, 1, 0, 0] = dN[:, :, 0]
    out[:, :, 0, 1, 0] = dN[:, :, 1]
    out[:, :, 0, 0, 1] = dN[:, :, 2]
    return out


class ElementTables:
    pass


def hex20_tables(itg_order=5):
    t = ElementTables()
    t.shape, t.n_a = "CUBE", 20
    pts, w = gauss_cube(itg_order, 3)
    t.ref_itp_vals, t.itg_weight = _pack(*hex20_eval(pts)), w
    fpts, fw = gauss_cube(itg_order, 2)
    nq = len(fw)
    face_ids = [[5, 3], [2, 4], [1, 6]]
    bref, bw, bt = [None] * 6, [None] * 6, [None] * 6
    for nd in range(3):
        tdim = [(i + nd) % 3 for i in (1, 2)]
        for outward in (0, 1):
            fid = face_ids[nd][outward] - 1
            p = np.zeros((nq, 3))
            p[:, tdim[0]], p[:, tdim[1]], p[:, nd] = fpts[:, 0], fpts[:, 1], float(outward)
            tan = np.zeros((nq, 3, 2))
            tan[:, tdim[0], 0] = -1.0 if outward == 0 else 1.0
            tan[:, tdim[1], 1] = 1.0
            bref[fid], bw[fid], bt[fid] = _pack(*hex20_eval(p)), fw.copy(), tan
    t.bdy_ref_itp_vals = np.stack(bref, axis=-1)
    t.bdy_itg_weights = np.stack(bw, axis=-1)
    t.bdy_tangent_directions = np.stack(bt, axis=-1)
    # topology used by the mesh builder (1-based like the reference tables)
    t.vertex_cp_ids = [1, 2, 4, 3, 5, 6, 8, 7]
    t.segment_vertices = [[1, 2], [2, 3], [3, 4], [4, 1], [1, 5], [2, 6], [3, 7], [4, 8], [5, 6], [6, 7], [7, 8], [8, 5]]
    t.segment_cp_ids = [9 + o for o in (0, 5, 1, 4, 8, 9, 11, 10, 2, 7, 3, 6)]
    t.face_vertices = [[1, 2, 3, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 4, 8, 7], [4, 1, 5, 8], [5, 6, 7, 8]]
    return t


def tet10_tables():
    t = ElementTables()
    t.shape, t.n_a = "SIMPLEX", 10
    pts, w = gauss_tet()
    t.ref_itp_vals, t.itg_weight = _pack(*tet10_eval(pts)), w
    bary, bw0 = gauss_triangle_bary()
    a, b, c = bary[:, 0], bary[:, 1], bary[:, 2]
    z = np.zeros_like(a)
    fp = [np.stack([b, c, z], 1), np.stack([b, z, c], 1), np.stack([b, c, a], 1), np.stack([z, b, c], 1)]
    bws = [bw0 * 0.5 for _ in range(4)]
    bws[2] = bws[2] * math.sqrt(3)
    tans = [np.array([[-1., 0., 0.], [0., 1., 0.]]), np.array([[0., 0., -1.], [1., 0., 0.]]),
            np.array([[-1., 1., 0.]]) / math.sqrt(2), np.array([[0., -1., 0.], [0., 0., 1.]])]
    tans[2] = np.stack([np.array([-1., 1., 0.]) / math.sqrt(2), np.array([-1., -1., 2.]) / math.sqrt(6)])
    bt = []
    for f in range(4):
        tan = np.zeros((len(bw0), 3, 2))
        tan[:, :, 0], tan[:, :, 1] = tans[f][0], tans[f][1]
        bt.append(tan)
    t.bdy_ref_itp_vals = np.stack([_pack(*tet10_eval(p)) for p in fp], axis=-1)
    t.bdy_itg_weights = np.stack(bws, axis=-1)
    t.bdy_tangent_directions = np.stack(bt, axis=-1)
    t.vertex_cp_ids = [1, 3, 6, 10]
    t.segment_vertices = [[1, 2], [2, 3], [3, 1], [1, 4], [2, 4], [3, 4]]
    t.segment_cp_ids = [2, 5, 4, 7, 8, 9]
    t.face_vertices = [[1, 2, 3], [1, 2, 4], [2, 3, 4], [3, 1, 4]]
    return t
