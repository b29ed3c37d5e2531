"""ORACLE: reference-element tables (reference src/mesh/spatial_discretization/*, src/misc/03_Polynomial.jl).

Restates the reference's sparse multivariate polynomial algebra (including the 1e-8 coefficient
drop of check_Clear), the hex/tet shape functions, Gauss tables and boundary frames, and evaluates
ref_itp_vals[q, a, d1, d2, d3] the way evaluate_Itp_Funcs does.
"""
import itertools
import math
import numpy as np


class Polynomial:
    """03_Polynomial.jl:2-10. factors: list[float]; orders: list[tuple[int]] (kept in insertion order)."""

    def __init__(self, factors, orders):
        self.factors = [float(f) for f in factors]
        self.orders = [tuple(int(o) for o in od) for od in orders]

    @staticmethod
    def zero(dim):
        return Polynomial([0.0], [(0,) * dim])

    @property
    def dim(self):
        return len(self.orders[0])

    def copy(self):
        return Polynomial(list(self.factors), list(self.orders))

    def check_clear(self):                       # :61-74
        keep = [abs(f) >= 1e-8 for f in self.factors]
        if not any(keep):
            self.factors, self.orders = [0.0], [(0,) * self.dim]
        else:
            self.factors = [f for f, k in zip(self.factors, keep) if k]
            self.orders = [o for o, k in zip(self.orders, keep) if k]
        return self

    def __add__(self, other):
        if isinstance(other, (int, float)):      # :20-35
            ans = self.copy()
            if other == 0:
                return ans
            od = (0,) * self.dim
            if od in self.orders:
                ans.factors[self.orders.index(od)] += other
                ans.check_clear()
            else:
                ans.factors.append(float(other)); ans.orders.append(od)
            return ans
        ans = self.copy()                        # :39-51
        for f, od in zip(other.factors, other.orders):
            if od in self.orders:
                ans.factors[self.orders.index(od)] += f
            else:
                ans.factors.append(f); ans.orders.append(od)
        return ans.check_clear()

    __radd__ = __add__

    def __neg__(self):
        return Polynomial([-f for f in self.factors], list(self.orders))

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if isinstance(other, (int, float)):      # :76-85
            if other == 0:
                return Polynomial.zero(self.dim)
            return Polynomial([f * other for f in self.factors], list(self.orders))
        ans = Polynomial.zero(self.dim)          # :90-106
        for f1, o1 in zip(self.factors, self.orders):
            for f2, o2 in zip(other.factors, other.orders):
                f = f1 * f2
                od = tuple(a + b for a, b in zip(o1, o2))
                if od in ans.orders:
                    ans.factors[ans.orders.index(od)] += f
                else:
                    ans.factors.append(f); ans.orders.append(od)
        return ans.check_clear()

    __rmul__ = __mul__

    def __truediv__(self, num):                  # :88
        return self * (1.0 / num)

    def __pow__(self, n):                        # :108-114
        ans = Polynomial([1.0], [(0,) * self.dim])
        for _ in range(n):
            ans = ans * self
        return ans

    def derivative(self, orders):                # :127-142
        p = self.copy()
        for i in range(len(p.factors)):
            od = p.orders[i]
            if min(a - b for a, b in zip(od, orders)) < 0:
                p.factors[i] = 0.0
                continue
            for d in range(self.dim):
                p.factors[i] *= math.factorial(od[d]) / math.factorial(od[d] - orders[d])
            p.orders[i] = tuple(a - b for a, b in zip(od, orders))
        return p.check_clear()

    def evaluate(self, pos):                     # :144-151
        s = 0.0
        for f, od in zip(self.factors, self.orders):
            t = 1.0
            for x, o in zip(pos, od):
                t *= float(x) ** o
            s += t * f
        return s


def basis_Tup(dim_id, dim_num, wi=1, wo=0):      # 03_Polynomial.jl:13-17 (dim_id 1-based)
    t = [wo] * dim_num
    t[dim_id - 1] = wi
    return tuple(t)


def substitute_Polynomial(p_src, src_dim, p_template):   # :116-125
    dim1, dim2 = p_src.dim, p_template.dim
    ans = Polynomial.zero(dim2)
    for f, od in zip(p_src.factors, p_src.orders):
        core = p_template ** od[src_dim - 1]
        base = Polynomial([f], [tuple(0 if (x == src_dim or x > dim1) else od[x - 1] for x in range(1, dim2 + 1))])
        ans = ans + core * base
    return ans


def collect_Basis(dim):
    return [Polynomial([1.0], [basis_Tup(i, dim)]) for i in range(1, dim + 1)]


def _product(ranges):
    """Iterators.product order: FIRST index fastest."""
    for t in itertools.product(*reversed(ranges)):
        yield tuple(reversed(t))


def init_Interpolation_Lagrange_1D(order):       # 102_Interpolations.jl:3-23
    pos = [i / order for i in range(order + 1)]
    out = []
    for a in range(order + 1):
        p = Polynomial.zero(1) + 1
        for b in range(order + 1):
            if a == b:
                continue
            den = pos[a] - pos[b]
            p = p * Polynomial([1.0 / den, -pos[b] / den], [(1,), (0,)])
        out.append(p)
    return out


def init_Interpolation_Simplex_Lagrange(order, dim):     # :46-62
    f1d = [Polynomial([1.0], [(0,)])] + [
        substitute_Polynomial(init_Interpolation_Lagrange_1D(o)[-1], 1, Polynomial([order / o], [(1,)]))
        for o in range(1, order + 1)]
    vol = [Polynomial([1.0], [basis_Tup(i, dim)]) for i in range(1, dim + 1)]
    vol.append(Polynomial([1.0] + [-1.0] * dim, [(0,) * dim] + [basis_Tup(i, dim) for i in range(1, dim + 1)]))
    tmpl = [[substitute_Polynomial(f, 1, vol[i]) for f in f1d] for i in range(dim + 1)]
    out = []
    for ipos in _product([range(order + 1)] * dim):
        last = order - sum(ipos)
        if last < 0:
            continue
        p = tmpl[0][ipos[0]]
        for i in range(1, dim):
            p = p * tmpl[i][ipos[i]]
        out.append(p * tmpl[dim][last])
    return out


def init_Interpolation_Cube_Serendipity(order, dim):     # :69-113 (order <= 2 branch and edges)
    xs = collect_Basis(dim)
    out = []
    assert order <= 2
    for coors in _product([range(2)] * dim):
        p = None
        for c, x in zip(coors, xs):
            t = (1 - c) - x
            p = t if p is None else p * t
        for i in range(1, order):
            s = [1 - 2 * c for c in coors]
            lin = None
            for si, x in zip(s, xs):
                lin = x * si if lin is None else lin + x * si
            p = p * ((sum(a * b for a, b in zip(s, coors)) + i / order) - lin)
        p = p / p.evaluate(coors)
        out.append(p)
    for edge in range(1, dim + 1):
        minor = [i for i in range(1, dim + 1) if i != edge]
        for mc in _product([range(2)] * (dim - 1)):
            base = None
            for c, d in zip(mc, minor):
                t = (1 - c) - xs[d - 1]
                base = t if base is None else base * t
            for ip in range(1, order):
                p = None
                for i in range(order + 1):
                    if i == ip:
                        continue
                    t = xs[edge - 1] - (i / order)
                    p = t if p is None else p * t
                p = p * base
                coor = [ip / order] * dim
                for c, d in zip(mc, minor):
                    coor[d - 1] = c
                p = p / p.evaluate(tuple(coor))
                out.append(p)
    return out


def evaluate_Itp_Funcs(itp_funcs, max_sd_order, itg_pos):   # 01_Classical_DIscretization.jl:83-98
    dim = itp_funcs[0].dim
    g = max_sd_order + 1
    vals = np.zeros((len(itg_pos), len(itp_funcs)) + (g,) * dim)
    for d_orders in _product([range(g)] * dim):
        batch = [f.derivative(d_orders) for f in itp_funcs]
        for q, pos in enumerate(itg_pos):
            for a, f in enumerate(batch):
                vals[(q, a) + tuple(d_orders)] = f.evaluate(pos)
    return vals


# ---- quadrature, 103_Integrations.jl ----
_G_POS = ((0.0,), (-1.0 / math.sqrt(3.0), 1.0 / math.sqrt(3.0)), (-math.sqrt(3.0 / 5.0), 0.0, math.sqrt(3.0 / 5.0)),
          (-math.sqrt(3.0 / 7.0 + 2.0 / 7.0 * math.sqrt(6.0 / 5.0)), -math.sqrt(3.0 / 7.0 - 2.0 / 7.0 * math.sqrt(6.0 / 5.0)),
           math.sqrt(3.0 / 7.0 - 2.0 / 7.0 * math.sqrt(6.0 / 5.0)), math.sqrt(3.0 / 7.0 + 2.0 / 7.0 * math.sqrt(6.0 / 5.0))))
_G_W = ((2.0,), (1.0, 1.0), (5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0),
        ((18.0 - math.sqrt(30.0)) / 36.0, (18.0 + math.sqrt(30.0)) / 36.0,
         (18.0 + math.sqrt(30.0)) / 36.0, (18.0 - math.sqrt(30.0)) / 36.0))
G_POS_SHIFTED = tuple(tuple(x / 2.0 + 0.5 for x in t) for t in _G_POS)    # :1-12
G_W_SHIFTED = tuple(tuple(x / 2.0 for x in t) for t in _G_W)


def init_Domain_Integration_Cube_Gauss(itg_order, dim):   # :14-19
    go = int(math.ceil((itg_order + 1) / 2))
    P, W = G_POS_SHIFTED[go - 1], G_W_SHIFTED[go - 1]
    pos = [tuple(P[i] for i in ids) for ids in _product([range(go)] * dim)]
    w = []
    for ids in _product([range(go)] * dim):
        t = 1.0
        for i in ids:
            t *= W[i]
        w.append(t)
    return pos, np.array(w)


def init_Boundary_Integration_Cube_Gauss(itg_order, dim):  # :21-58
    pos, bw = init_Domain_Integration_Cube_Gauss(itg_order, dim - 1)
    face_ids = {2: [[4, 2], [1, 3]], 3: [[5, 3], [2, 4], [1, 6]]}[dim]
    nq, nf = len(pos), 2 * dim
    bpos = [[None] * nq for _ in range(nf)]
    btan = [np.zeros((nq, dim, dim - 1)) for _ in range(nf)]
    for nd in range(1, dim + 1):
        tdim = [(i + nd - 1) % dim + 1 for i in range(1, dim)]
        for outward in (0, 1):
            fid = face_ids[nd - 1][outward]
            raw = np.zeros((nq, dim, dim - 1))
            for i in range(dim - 1):
                raw[:, tdim[i] - 1, i] = 1
            if dim == 2:
                if outward + nd != 2:
                    raw *= -1
            elif outward == 0:
                raw[:, :, 0] *= -1
            btan[fid - 1] = raw
            for q in range(nq):
                p = [0.0] * dim
                for t, v in zip(tdim, pos[q]):
                    p[t - 1] = v
                p[nd - 1] = float(outward)
                bpos[fid - 1][q] = tuple(p)
    return bpos, [bw.copy() for _ in range(nf)], btan


_TRI_POS = (((0.10128650732345633880098736191512383,), (0.47014206410511508977044120951344760,), ()),)
_TRI_W = ((0.12593918054482715259568394550018133, 0.13239415278850618073764938783315200, 9.0 / 40.0),)
_TET_POS = (((0.31088591926330060979734573376345783,), (0.09273525031089122640232391373703061,),
             (-0.04550370412564964949188052627933943,)),)
_TET_W = ((0.11268792571801585079918565233328633, 0.07349304311636194954371020548632750,
           0.04254602077708146643806942812025744),)


def _init_Integration_Triangle_Gauss(itg_order):          # :80-113 (order <= 5 table only)
    assert itg_order <= 5, "oracle restates the order<=5 triangle table only"
    pos, w = [], []
    for p, wt in zip(_TRI_POS[0], _TRI_W[0]):
        if len(p) == 0:
            pos.append((1 / 3,) * 3); w.append(wt)
        elif len(p) == 1:
            a = p[0]
            for i in range(1, 4):
                pos.append(basis_Tup(i, 3, 1 - 2 * a, a)); w.append(wt)
    return pos, np.array(w)


def _init_Integration_Tetrahedron_Gauss(itg_order):       # :145-201 (order <= 5 table only)
    assert itg_order <= 5, "oracle restates the order<=5 tetrahedron table only"
    pos, w = [], []
    for p, wt in zip(_TET_POS[0], _TET_W[0]):
        a = p[0]
        if a >= 0:
            for i in range(1, 5):
                pos.append(basis_Tup(i, 4, 1 - 3 * a, a)); w.append(wt)
        else:
            b = -a
            for (i, j) in _product([range(1, 5)] * 2):
                if i >= j:
                    continue
                s = [b] * 4
                s[i - 1] = 0.5 - b; s[j - 1] = 0.5 - b
                pos.append(tuple(s)); w.append(wt)
    return pos, np.array(w)


def init_Domain_Integration_Tetrahedron_Gauss(itg_order):  # :203-206
    pos, w = _init_Integration_Tetrahedron_Gauss(itg_order)
    return [(p[1], p[2], p[3]) for p in pos], w / 6


def init_Boundary_Integration_Tetrahedron_Gauss(itg_order):  # :208-238
    pos, bw = _init_Integration_Triangle_Gauss(itg_order)
    bws = [bw * 0.5 for _ in range(4)]
    bws[2] = bws[2] * math.sqrt(3)
    nq = len(pos)
    bpos = [[None] * nq for _ in range(4)]
    btan = [np.zeros((nq, 3, 2)) for _ in range(4)]
    for i, (a, b, c) in enumerate(pos):
        bpos[0][i] = (b, c, 0.0)
        bpos[1][i] = (b, 0.0, c)
        bpos[2][i] = (b, c, a)
        bpos[3][i] = (0.0, b, c)
    btan[0][:, :, 0] = [-1., 0., 0.]; btan[0][:, :, 1] = [0., 1., 0.]
    btan[1][:, :, 0] = [0., 0., -1.]; btan[1][:, :, 1] = [1., 0., 0.]
    btan[2][:, :, 0] = np.array([-1., 1., 0.]) / math.sqrt(2); btan[2][:, :, 1] = np.array([-1., -1., 2.]) / math.sqrt(6)
    btan[3][:, :, 0] = [0., -1., 0.]; btan[3][:, :, 1] = [0., 0., 1.]
    return bpos, bws, btan


# ---- element topology, 101_Structures.jl ----
class ElementStructure:
    pass


def init_Structure_Cube3D_Serendipity(order):              # :224-247
    s = ElementStructure()
    cps = order - 1
    s.vertex_cp_ids = np.array([[1, 2, 4, 3, 5, 6, 8, 7]])
    seg_orders = [0, 5, 1, 4, 8, 9, 11, 10, 2, 7, 3, 6]
    s.segment_cp_ids = np.array([[seg_orders[j] * cps + 8 + i for j in range(12)] for i in range(1, cps + 1)]).reshape(cps, 12)
    s.segment_cp_pos = np.array([[1. - i / order, i / order] for i in range(1, cps + 1)]).T.reshape(2, cps)
    s.segment_start_vertex = [1, 2, 4, 1, 1, 2, 3, 4, 5, 6, 8, 5]
    s.face_cp_ids = np.zeros((0, 6), int)
    s.block_cp_ids = np.zeros(0, int)
    return s


def init_Structure_Tetrahedron_Lagrange(order):             # :129-196, order 2 (no face/block nodes)
    assert order == 2, "oracle restates tet10 (order 2) only"
    s = ElementStructure()
    cpd = order + 1
    s.vertex_cp_ids = np.array([[1, cpd, cpd * (cpd + 1) // 2, cpd * (cpd + 1) * (cpd + 2) // 6]])
    cps = cpd - 2
    seg = np.zeros((cps, 6), int)
    seg[:, 0] = [i + 1 for i in range(1, cps + 1)]
    seg[:, 1] = [(2 * cpd - i) * (i + 1) // 2 for i in range(1, cps + 1)]
    seg[:, 2] = [(2 * cpd - i + 1) * i // 2 + 1 for i in range(1, cps + 1)]
    last_final = cpd * (cpd + 1) // 2
    for k in range(1, cps + 1):
        seg[k - 1, 3] = last_final + 1
        cur = cpd - k
        seg[k - 1, 4] = last_final + cur
        for j in range(1, cps - k + 1):
            cur += cpd - k - j
        last_final += cur + 1
        seg[k - 1, 5] = last_final
    s.segment_cp_ids = seg
    s.segment_cp_pos = np.array([[1. - i / order, i / order] for i in range(1, cps + 1)]).T.reshape(2, cps)
    s.segment_start_vertex = [1, 2, 1, 1, 2, 3]
    s.face_cp_ids = np.zeros((0, 4), int)
    s.block_cp_ids = np.zeros(0, int)
    return s


class ClassicalDiscretization:
    pass


def initialize_Classical_Element(dim, shape, itp_order, max_sd_order, itg_order, itp_type="Lagrange"):
    """01_Classical_DIscretization.jl:37-81 for the two 3-D second-order elements in scope."""
    assert dim == 3 and itp_order == 2
    sp = ClassicalDiscretization()
    if shape == "CUBE":
        assert itp_type == "Serendipity"
        sp.element_structure = init_Structure_Cube3D_Serendipity(itp_order)
        funcs = init_Interpolation_Cube_Serendipity(itp_order, dim)
        itg_pos, itg_w = init_Domain_Integration_Cube_Gauss(itg_order, dim)
        bpos, bw, btan = init_Boundary_Integration_Cube_Gauss(itg_order, dim)
    else:
        sp.element_structure = init_Structure_Tetrahedron_Lagrange(itp_order)
        funcs = init_Interpolation_Simplex_Lagrange(itp_order, dim)
        itg_pos, itg_w = init_Domain_Integration_Tetrahedron_Gauss(itg_order)
        bpos, bw, btan = init_Boundary_Integration_Tetrahedron_Gauss(itg_order)
    sp.shape, sp.dim, sp.max_sd_order = shape, dim, max_sd_order
    sp.itp_funcs = funcs
    sp.itp_func_num, sp.itg_func_num, sp.bdy_itg_func_num = len(funcs), len(itg_w), len(bw[0])
    sp.itg_pos, sp.itg_weight = itg_pos, np.asarray(itg_w, dtype=np.float64)
    sp.bdy_itg_pos, sp.bdy_itg_weights, sp.bdy_tangent_directions = bpos, bw, btan
    sp.ref_itp_vals = evaluate_Itp_Funcs(funcs, max_sd_order, itg_pos)
    sp.bdy_ref_itp_vals = [evaluate_Itp_Funcs(funcs, max_sd_order, p) for p in bpos]
    return sp
