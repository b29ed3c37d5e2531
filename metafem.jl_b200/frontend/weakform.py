"""Python stand-in for the part of MetaFEM's symbolic front end that feeds the hot path.

Julia is not installed in this environment, so the (dual word, base term) lists that
``build_WeakForm`` / ``construct_AssembleWeakform`` would hand to the code generator
(reference src/symbolics/10_WeakForm.jl:72-124, src/solver/02_LocalAssembly.jl:30-111) are
re-derived here with sympy for the five BASELINE configs and serialised into the *kernel spec*:
the JSON-able dict that ``compile_Updater_GPU`` sends through ``mfb_kernel_compile``.

Spec layout (all positions 0-based, ``sd`` = list of 1-based directions like the reference's sd_ids):
  basic_vars        sorted base symbols (02_LocalAssembly.jl:93-94)
  max_time_level    highest td_order of any unknown word
  sparse_mapping    list of [dual_pos, base_pos]; index = block number (:70-74,104-105)
  globals / cp_vars names of GLOBAL_VAR scalars / CONTROLPOINT_VAR nodal arrays
  qp_vars           names of INTEGRATION_POINT_VAR arrays ([n_q, n_el], one value per quadrature point)
  blocks            [domain, boundary groups...] each with innervars, extervars, temps,
                    residues, linear_gradients, nonlinear_gradients; expressions are C
                    expressions over the word symbols (also valid Python under numpy).
                    qp_calls: user callbacks at quadrature points (J2Plasticity.jl:55: ``ep{i,j} =
                    strain_updater(e{1,1}, ...)``; 08_Tensor.jl:175-183,210): func name, C expressions of
                    the arguments, names of the argument arrays the two-phase update fills, names of the
                    INTEGRATION_POINT_VAR outputs the residual then reads as external words (kind "qp").
"""
import sympy as sp
from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e == 2:
            s = self._print(b)
            return f"(({s})*({s}))"
        if e == -1:
            return f"(1.0/({self._print(b)}))"
        if e == -2:
            s = self._print(b)
            return f"(1.0/(({s})*({s})))"
        return super()._print_Pow(expr)


_printer = _Printer()


def cexpr(e):
    return _printer.doprint(e)


def word_sym(base, td=0, sd=()):
    """word_To_TotalSym analogue: d1, d1_2 (d/dx2), T_t1, d3_t2_1 ..."""
    s = base
    if td:
        s += f"_t{td}"
    if sd:
        s += "_" + "".join(str(i) for i in sd)
    return s


class Form:
    """One weak-form block: list of (dual word, base expression)."""

    def __init__(self, kind, bg_ID=0):
        self.kind, self.bg_ID = kind, bg_ID
        self.terms = []          # (base, td, sd, expr)

    def add(self, dual_base, dual_sd, expr):
        self.terms.append((dual_base, 0, tuple(dual_sd), sp.sympify(expr)))


class Physics:
    """Collects words and forms; ``spec()`` performs what initialize_LocalAssembly! does."""

    def __init__(self, dim=3):
        self.dim = dim
        self.inner = {}      # Symbol -> (base, td, sd)
        self.cp = {}         # Symbol -> (local, sd)
        self.normal = {}     # Symbol -> c
        self.glob = {}       # Symbol -> name
        self.qp = {}         # Symbol -> name (INTEGRATION_POINT_VAR outputs of a callback)
        self.qp_calls = []   # dict(func, args [sympy], outs [Symbol])
        self.forms = []

    # -- word constructors ---------------------------------------------------------------
    def u(self, base, td=0, sd=()):
        s = sp.Symbol(word_sym(base, td, sd), real=True)
        self.inner[s] = (base, td, tuple(sd))
        return s

    def cpvar(self, local, sd=()):
        s = sp.Symbol(word_sym(local, 0, sd), real=True)
        self.cp[s] = (local, tuple(sd))
        return s

    def n(self, c):
        s = sp.Symbol(f"n{c}", real=True)
        self.normal[s] = c
        return s

    def g(self, name):
        s = sp.Symbol(name, real=True)
        self.glob[s] = name
        return s

    def qpcall(self, func, args, out_names, builtin=None, prefix=None):
        """``outs = Main.func(args...)`` evaluated on whole [n_q, n_el] arrays (08_Tensor.jl:175-183,210); the outputs
        are external INTEGRATION_POINT_VAR words: zero variation (09_Differentiation.jl:68-69)."""
        outs = [sp.Symbol(n, real=True) for n in out_names]
        for o in outs:
            self.qp[o] = o.name
        # builtin = "j2_return_map": the library's radial-return update is inlined into the residual kernel instead of
        # calling back (state arrays "<prefix>.*", parameters as GLOBAL_VARs "<prefix>_lam, _mu, _Eb, _Ep, _fres")
        self.qp_calls.append(dict(func=func, args=[sp.sympify(a) for a in args], outs=outs, builtin=builtin, prefix=prefix))
        if builtin:
            for g in ("lam", "mu", "Eb", "Ep", "fres"):
                self.g(f"{prefix}_{g}")
        return outs

    def form(self, kind, bg_ID=0):
        f = Form(kind, bg_ID)
        self.forms.append(f)
        return f

    # -- assembly planning ---------------------------------------------------------------
    def spec(self, cse=True):
        basic_vars = sorted({b for (b, _, _) in self.inner.values()})
        pos = {b: i for i, b in enumerate(basic_vars)}
        max_td = max(td for (_, td, _) in self.inner.values())
        blocks, sparse = [], set()
        for f in self.forms:
            res, lin, nonlin = [], [], []
            for (db, dtd, dsd, expr) in f.terms:
                # regulate: merge terms sharing a dual word (simplification of the reference's LHS regulation)
                res.append(dict(dual_pos=pos[db], dual_sd=list(dsd), expr=expr))
                for w in sorted((s for s in expr.free_symbols if s in self.inner), key=lambda s: s.name):
                    de = sp.diff(expr, w)
                    if de == 0:
                        continue
                    wb, wtd, wsd = self.inner[w]
                    ent = dict(dual_pos=pos[db], dual_sd=list(dsd), deriv_pos=pos[wb], deriv_td=wtd,
                               deriv_sd=list(wsd), expr=de)
                    # linear only without inner words and without integration-point externals (02_LocalAssembly.jl:49)
                    (nonlin if any(s in self.inner or s in self.qp for s in de.free_symbols) else lin).append(ent)
                    sparse.add((pos[db], pos[wb]))
            res = _merge(res, ("dual_pos", "dual_sd"))
            lin = _merge(lin, ("dual_pos", "dual_sd", "deriv_pos", "deriv_td", "deriv_sd"))
            nonlin = _merge(nonlin, ("dual_pos", "dual_sd", "deriv_pos", "deriv_td", "deriv_sd"))
            exprs = [t["expr"] for t in res + lin + nonlin]
            temps = []
            if cse and exprs:
                rep, red = sp.cse(exprs, symbols=sp.numbered_symbols("tmp"), optimizations="basic")
                temps = [dict(sym=str(s), expr=cexpr(e)) for s, e in rep]
                for t, e in zip(res + lin + nonlin, red):
                    t["expr"] = e
            used = set()
            for t in res + lin + nonlin:
                used |= t["expr"].free_symbols
            for s_, e_ in (rep if (cse and exprs) else []):
                used |= e_.free_symbols
            for t in res + lin + nonlin:
                t["expr"] = cexpr(t["expr"])
            calls = []
            for c in self.qp_calls:
                if any(o in used for o in c["outs"]):
                    for a in c["args"]:
                        used |= a.free_symbols
                    calls.append(dict(func=c["func"], args=[cexpr(a) for a in c["args"]],
                                      arg_names=[f"{c['func']}_arg{k + 1}" for k in range(len(c["args"]))],
                                      outs=[o.name for o in c["outs"]], builtin=c["builtin"], prefix=c["prefix"]))
                    if c["builtin"]:
                        used |= {sym for sym, nm in self.glob.items() if nm.startswith(c["prefix"] + "_")}
            inner = [dict(sym=s.name, pos=pos[self.inner[s][0]], td=self.inner[s][1], sd=list(self.inner[s][2]))
                     for s in sorted((s for s in used if s in self.inner), key=lambda s: s.name)]
            ext = []
            for s in sorted(used, key=lambda s: s.name):
                if s in self.cp:
                    ext.append(dict(sym=s.name, kind="cp", local=self.cp[s][0], sd=list(self.cp[s][1])))
                elif s in self.normal:
                    ext.append(dict(sym=s.name, kind="normal", c=self.normal[s]))
                elif s in self.glob:
                    ext.append(dict(sym=s.name, kind="global"))
                elif s in self.qp:
                    ext.append(dict(sym=s.name, kind="qp"))
            blocks.append(dict(kind=f.kind, bg_ID=f.bg_ID, innervars=inner, extervars=ext, temps=temps,
                               residues=res, linear_gradients=lin, nonlinear_gradients=nonlin, qp_calls=calls))
        cp_vars = sorted({v[0] for v in self.cp.values()})
        return dict(dim=self.dim, basic_vars=basic_vars, max_time_level=max_td,
                    sparse_mapping=[list(p) for p in sorted(sparse)],
                    globals=sorted(self.glob.values()), cp_vars=cp_vars, qp_vars=sorted(self.qp.values()),
                    blocks=blocks)


def _merge(terms, keys):
    out, idx = [], {}
    for t in terms:
        k = tuple(tuple(t[x]) if isinstance(t[x], list) else t[x] for x in keys)
        if k in idx:
            out[idx[k]]["expr"] = out[idx[k]]["expr"] + t["expr"]
        else:
            idx[k] = len(out)
            out.append(dict(t))
    return [t for t in out if (sp.expand(t["expr"]) if sp.count_ops(t["expr"]) < 30 else t["expr"]) != 0]


# ------------------------------------------------------------------------------------------
# The BASELINE configs
# ------------------------------------------------------------------------------------------
def thermal_conduction(k=0.6, h=25.0, T_env=293.15, alpha=0.0, bgs=(1,)):
    """examples/thermal_conduction/3D_Script.jl:21-35."""
    P = Physics(3)
    T = P.u("T")
    s = P.cpvar("s")
    dom = P.form("domain")
    dom.add("T", (), s + alpha * (T_env - T))
    for i in (1, 2, 3):
        dom.add("T", (i,), -k * P.u("T", 0, (i,)))
    for bg in bgs:
        P.form("boundary", bg).add("T", (), h * (T_env - T))
    return P.spec()


def _eps_sigma(P, lam, mu):
    gd = [[P.u(f"d{i}", 0, (j,)) for j in (1, 2, 3)] for i in (1, 2, 3)]
    eps = [[(gd[i][j] + gd[j][i]) / 2 for j in range(3)] for i in range(3)]
    tr = eps[0][0] + eps[1][1] + eps[2][2]
    sig = [[lam * (1 if i == j else 0) * tr + 2 * mu * eps[i][j] for j in range(3)] for i in range(3)]
    return gd, eps, sig


_VOIGT = {(1, 1): 1, (2, 2): 2, (3, 3): 3, (2, 3): 4, (3, 2): 4, (1, 3): 5, (3, 1): 5, (1, 2): 6, (2, 1): 6}


def linear_elasticity(lam, mu, tau_b, fixed_bg=1, fixed_components=(1, 2, 3), traction_bgs=((2, "sl"),)):
    """examples/linear_elasticity/cantilever/3D_Script.jl:45-63 (and stress_concentration/3D_Script.jl:37-55).

    traction_bgs: (bg_ID, nodal symmetric tensor name) -> Bilinear(d{i}, s{i,j} n{j}), Voigt-named nodal arrays.
    fixed_bg may be an int (all ``fixed_components`` on one group) or a dict {bg_ID: component}.
    """
    P = Physics(3)
    gd, eps, sig = _eps_sigma(P, lam, mu)
    dom = P.form("domain")
    # -(eps_ij, sigma_ij): dual eps_ij = (d_i;j + d_j;i)/2  ->  dual word d_a;b with base -(sigma_ab + sigma_ba)/2
    for a in range(3):
        for b in range(3):
            dom.add(f"d{a+1}", (b + 1,), -(sig[a][b] + sig[b][a]) / 2)
    fixed = fixed_bg if isinstance(fixed_bg, dict) else {fixed_bg: tuple(fixed_components)}
    for bg, comps in fixed.items():
        f = P.form("boundary", bg)
        for c in (comps if isinstance(comps, (tuple, list)) else (comps,)):
            f.add(f"d{c}", (), tau_b * (P.cpvar(f"dw{c}") - P.u(f"d{c}")))
    for bg, name in traction_bgs:
        f = P.form("boundary", bg)
        comps = name[1] if isinstance(name, tuple) else None
        nm = name[0] if isinstance(name, tuple) else name
        for i in (1, 2, 3):
            js = [j for j in (1, 2, 3) if comps is None or (i, j) in comps]
            if js:
                f.add(f"d{i}", (), sum(P.cpvar(f"{nm}{_VOIGT[(i, j)]}") * P.n(j) for j in js))
    return P.spec()


def neo_hookean(fixed_bg=1, traction_bg=3):
    """examples/hyper_elasticity/static_Neo_Hookean.jl:38-57; mu, lambda, tau_b are GLOBAL_VARs."""
    P = Physics(3)
    mu, lam, tau = P.g("mu"), P.g("lam"), P.g("tau_b")
    gd = [[P.u(f"d{i}", 0, (j,)) for j in (1, 2, 3)] for i in (1, 2, 3)]
    F = sp.Matrix(3, 3, lambda i, j: (1 if i == j else 0) + gd[i][j])
    J = F.det(method="berkowitz")
    C = F.T * F
    W = sp.Rational(1, 2) * mu * (C.trace() - 3 - 2 * sp.log(J)) + sp.Rational(1, 2) * lam * (J - 1) ** 2
    dom = P.form("domain")
    for i in range(3):
        for j in range(3):
            dom.add(f"d{i+1}", (j + 1,), -sp.diff(W, gd[i][j]))
    f = P.form("boundary", fixed_bg)
    for c in (1, 2, 3):
        f.add(f"d{c}", (), tau * (P.cpvar(f"dw{c}") - P.u(f"d{c}")))
    f = P.form("boundary", traction_bg)
    for i in (1, 2, 3):
        # non-symmetric second-order tensor naming: 1 + (i-1) + 3 (j-1)  (03_Word.jl:66)
        f.add(f"d{i}", (), sum(P.cpvar(f"Pl{1 + (i - 1) + 3 * (j - 1)}") * P.n(j) for j in (1, 2, 3)))
    return P.spec()


def thermo_elasticity(E=210e3, nu=0.0, tau_b=None, rho=1e3, c=0.01, h=100.0, C=1000.0, k=100.0, alpha=0.05e-3,
                      L_box=1.0, fixed_bg=1, thermal_bg=3):
    """examples/thermal_elasticity/themal_hypo_elasticity.jl:45-71: fields T, d with first time derivatives
    (max_time_level = 1). No leading minus in this script's forms (only consistency matters)."""
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    tau_b = 1000 * E / L_box if tau_b is None else tau_b
    P = Physics(3)
    T = P.u("T")
    gd = [[P.u(f"d{i}", 0, (j,)) for j in (1, 2, 3)] for i in (1, 2, 3)]
    eps = [[(gd[i][j] + gd[j][i]) / 2 - (alpha * T if i == j else 0) for j in range(3)] for i in range(3)]
    tr = eps[0][0] + eps[1][1] + eps[2][2]
    sig = [[lam * (1 if i == j else 0) * tr + 2 * mu * eps[i][j] for j in range(3)] for i in range(3)]
    dom = P.form("domain")
    dom.add("T", (), C * P.u("T", 1))                                   # C (T, T_t)
    for i in (1, 2, 3):
        dom.add("T", (i,), k * P.u("T", 0, (i,)))                       # k (T_i, T_i)
    # (eps_ij, sigma_ij): the variation of eps_ij = sym grad d - alpha T delta_ij gives a d_a;b test term and a T test term
    for a in range(3):
        for b in range(3):
            dom.add(f"d{a+1}", (b + 1,), (sig[a][b] + sig[b][a]) / 2)
    dom.add("T", (), -alpha * (sig[0][0] + sig[1][1] + sig[2][2]))
    for i in (1, 2, 3):
        dom.add(f"d{i}", (), rho * c * P.u(f"d{i}", 1))                 # (d_i, rho c d_i,t)
    f = P.form("boundary", fixed_bg)
    for i in (1, 2, 3):
        f.add(f"d{i}", (), tau_b * P.u(f"d{i}"))
    P.form("boundary", thermal_bg).add("T", (), h * (T - P.cpvar("Te")))
    return P.spec()


def j2_plasticity(E=100e3, nu=0.0, rho=1e3, c=2.0, tau_b=None, L_box=1.0, fixed_bg=1, traction_bg=2, fused=False):
    """examples/hypo_elastic_plasticity/J2Plasticity.jl:44-63: small-strain J2 flow with the plastic strain ``ep`` an
    INTEGRATION_POINT_VAR produced by the user callback ``strain_updater`` (the return map, :118-198); second time
    derivatives (max_time_level = 2). ``ep`` has zero variation, so the tangent is the constant elastic one and lands
    in K_linear together with the inertia/damping terms."""
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    tau_b = 1000 * E / L_box ** 2 if tau_b is None else tau_b
    P = Physics(3)
    gd = [[P.u(f"d{i}", 0, (j,)) for j in (1, 2, 3)] for i in (1, 2, 3)]
    e = [[(gd[i][j] + gd[j][i]) / 2 for j in range(3)] for i in range(3)]
    # ep{i,j} = strain_updater(e{1,1}, e{1,2}, e{1,3}, e{2,2}, e{2,3}, e{3,3}); six outputs in Voigt order (08_Tensor.jl:160)
    # fused = True: the callback is the library's built-in return map, inlined into the residual kernel (no two-phase update)
    epv = P.qpcall("strain_updater", [e[0][0], e[0][1], e[0][2], e[1][1], e[1][2], e[2][2]],
                   [f"ep{k}" for k in range(1, 7)], builtin="j2_return_map" if fused else None, prefix="j2" if fused else None)
    ep = [[epv[_VOIGT[(i + 1, j + 1)] - 1] for j in range(3)] for i in range(3)]
    ee = [[e[i][j] - ep[i][j] for j in range(3)] for i in range(3)]
    tr = ee[0][0] + ee[1][1] + ee[2][2]
    sig = [[2 * mu * ee[i][j] + (lam * tr if i == j else 0) for j in range(3)] for i in range(3)]
    dom = P.form("domain")
    for i in range(3):
        for j in range(3):
            dom.add(f"d{i+1}", (j + 1,), sig[i][j])                                    # Bilinear(d{i;j}, sigma{i,j})
    for i in (1, 2, 3):
        dom.add(f"d{i}", (), rho * (c * P.u(f"d{i}", 1) + P.u(f"d{i}", 2)))           # rho (c d_i,t + d_i,tt)
    f = P.form("boundary", fixed_bg)
    for i in (1, 2, 3):
        f.add(f"d{i}", (), tau_b * (P.u(f"d{i}") - P.cpvar(f"dw{i}")))
    f = P.form("boundary", traction_bg)
    for i in (1, 2, 3):
        f.add(f"d{i}", (), -sum(P.cpvar(f"sl{_VOIGT[(i, j)]}") * P.n(j) for j in (1, 2, 3)))
    return P.spec()
