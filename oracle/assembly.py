"""ORACLE: global numbering, sparsity pattern and term-by-term assembly (numpy restatement).

reference: src/solver/03_GlobalAssembly.jl:6-168      (DOF numbering, assemble_X!, assemble_SparseID!, K_I/K_J)
           src/misc/04_GPU_Utils.jl:87-118            (COO sort -> K_val_ids, row pointer by atomic_min)
           src/solver/05_CodeGenerator.jl:1-258       (what the generated updaters do, term by term)
           src/solver/06_FEM_Kernel.jl:1-13,28-45,65-79 (_Var_Basic, _Kval_Basic, _Res_Basic)
The kernel spec (see metafem.jl_b200/frontend/weakform.py docstring) is input data; its
expressions are evaluated with numpy here. All stored IDs are 1-based like the reference.
"""
import numpy as np

from .femdict import FemDict, I32I32_To_UI64, UI64_To_UpperHalf, UI64_To_LowerHalf
from . import cpath

_NS = dict(pow=np.power, log=np.log, exp=np.exp, sqrt=np.sqrt, fabs=np.abs)


def local_sym(base, td):
    return base + (f"_t{td}" if td else "")


class GlobalField:
    """GlobalField, 01_Types.jl:110-132."""
    def __init__(self):
        self.max_time_level = 0
        self.basicfield_size = 0
        self.converge_tol = 0.0
        self.t = 0.0
        self.dt = 1.0


class Domain:
    """One-workpiece FEM_Domain as far as the hot path needs it."""
    def __init__(self, mesh, spec, dissipative=True):
        self.mesh, self.spec = mesh, spec
        self.globalfield = GlobalField()
        self.global_vars = {g: 0.0 for g in spec["globals"]}
        N = mesh.variable_size
        self.cp = {}
        for b in spec["basic_vars"]:
            for td in range(spec["max_time_level"] + 1):
                self.cp[local_sym(b, td)] = np.zeros(N)
        for v in spec["cp_vars"]:
            self.cp[v] = np.zeros(N)
        # GeneralAlpha(dissipative) 04_Time_Domain.jl:1-8
        self.alpha_params = (1.0, 1.0, 1.0)
        self.gamma_params = (1.0, 1.0) if dissipative else (0.5, 0.5)
        self.beta_params = None
        self.K_params = None
        self.linear_solver = None
        self.callbacks = {}      # quadrature-point callbacks (Main.<func> in the generated code, 08_Tensor.jl:210)


def sd_slot(sd):
    """sd_ids -> slot of integral_vals: () -> 0 (N), (i,) -> i (d/dx_i). Only max_sd_order = 1 is in scope."""
    assert len(sd) <= 1
    return 0 if len(sd) == 0 else int(sd[0])


def assemble_Global_Variables(dom):
    """03_GlobalAssembly.jl:6-37 (single workpiece, no holes: global_cpID == table slot)."""
    mesh, spec, gf = dom.mesh, dom.spec, dom.globalfield
    N = mesh.variable_size
    mesh.global_cpID = np.arange(1, N + 1, dtype=np.int32)
    mesh.global_cpIDs = mesh.global_cpID[mesh.controlpoint_IDs - 1]
    gf.basicfield_size = len(spec["basic_vars"]) * N
    gf.max_time_level = spec["max_time_level"]
    size = (gf.max_time_level + 1) * gf.basicfield_size
    gf.x, gf.dx, gf.x_star = np.zeros(size), np.zeros(size), np.zeros(size)
    gf.residue = np.zeros(gf.basicfield_size)
    assemble_X(dom)
    assemble_SparseID(dom)


def _x_ranges(dom):
    spec, gf, N = dom.spec, dom.globalfield, dom.mesh.variable_size
    for pos, b in enumerate(spec["basic_vars"]):
        for td in range(spec["max_time_level"] + 1):
            s = pos * N + td * gf.basicfield_size
            yield local_sym(b, td), slice(s, s + N)


def assemble_X(dom):
    """:44-56."""
    for sym, sl in _x_ranges(dom):
        dom.globalfield.x[sl] = dom.cp[sym]


def dessemble_X(dom):
    """:63-75."""
    for sym, sl in _x_ranges(dom):
        dom.cp[sym][:] = dom.globalfield.x[sl]


def assemble_SparseID(dom):
    """:77-140 + assemble_KIJ! :142-168 + sort_CUSPARSE_COO!/generate_J_ptr (04_GPU_Utils.jl:87-118)."""
    mesh, spec, gf = dom.mesh, dom.spec, dom.globalfield
    cpi = mesh.controlpoint_IDs
    n_a, n_el = cpi.shape
    N = mesh.variable_size
    d = FemDict()
    keys = np.empty(n_a * n_a * n_el, np.uint64)
    c = 0
    for i in range(n_a):
        for j in range(n_a):
            keys[c * n_el:(c + 1) * n_el] = I32I32_To_UI64(cpi[i], cpi[j])
            c += 1
    d.set_ids(keys)
    tot = d.total_ids()
    unit = len(tot)
    d.vals[tot - 1] = np.arange(1, unit + 1, dtype=np.int32)        # local sparse IDs = rank of occupied slot
    sid = np.zeros((n_a, n_a, n_el), np.int32)
    for i in range(n_a):
        for j in range(n_a):
            loc = d.get_ids(I32I32_To_UI64(cpi[i], cpi[j]))
            assert np.all(loc > 0)
            sid[i, j] = d.vals[loc - 1]
    mesh.sparse_IDs_by_el = sid
    mesh.sparse_unitsize = unit
    nblk = len(spec["sparse_mapping"])
    nnz = nblk * unit
    K_I = np.zeros(nnz, np.int32)
    K_J = np.zeros(nnz, np.int32)
    k_keys = d.keys[tot - 1]
    rows, cols = UI64_To_UpperHalf(k_keys), UI64_To_LowerHalf(k_keys)
    for m, (dual_pos, base_pos) in enumerate(spec["sparse_mapping"]):
        K_I[m * unit:(m + 1) * unit] = rows + dual_pos * N
        K_J[m * unit:(m + 1) * unit] = cols + base_pos * N
    # canonical (row, col) sort; K_val_ids = permutation (1-based), SURVEY §8c
    perm = np.lexsort((K_J, K_I))
    gf.K_val_ids = (perm + 1).astype(np.int32)
    gf.K_I, gf.K_J = K_I[perm], K_J[perm]
    n = gf.basicfield_size
    J_ptr = np.full(n + 1, nnz + 1, np.int32)                        # generate_J_ptr: atomic_min of position
    np.minimum.at(J_ptr, gf.K_I - 1, np.arange(1, nnz + 1, dtype=np.int32))
    gf.K_J_ptr = J_ptr
    gf.K_linear = np.zeros(nnz)
    gf.K_total = np.zeros(nnz)


# ---------------------------------------------------------------------------------------------
# generated updaters
# ---------------------------------------------------------------------------------------------
def _block_context(dom, blk):
    """elIDs / local_itg_hostIDs / tables of one generated block (05_CodeGenerator.jl:162-190)."""
    mesh = dom.mesh
    if blk["kind"] == "domain":
        n_el = mesh.controlpoint_IDs.shape[1]
        el = np.arange(n_el)
        return Ctx(el, el, mesh.integral_vals, mesh.integral_weights, None)
    f = mesh.bg_fIDs[blk["bg_ID"]] - 1
    el = mesh.facet_element_ID[f] - 1
    return Ctx(el, f, mesh.facet_integral_vals, mesh.facet_integral_weights[f], mesh.facet_normal_directions[f])


class Ctx:
    """elIDs, local_itg_hostIDs, local_integral_vals, weights[:, hostIDs], normals of one block."""
    def __init__(self, el, host, itp, w, normals):
        self.el, self.host, self.itp, self.w, self.normals = el, host, np.ascontiguousarray(itp), w, normals


def _declare_vars(dom, blk, cx, which):
    """declare_Innervar_GPU / declare_Extervar_GPU (:1-50): _Var_Basic per word."""
    mesh, gf = dom.mesh, dom.globalfield
    N = mesh.variable_size
    env = dict(_NS)
    if which != "linear":
        for w in blk["innervars"]:
            shift = w["td"] * gf.basicfield_size + w["pos"] * N
            env[w["sym"]] = cpath.var_basic(cx.itp, sd_slot(w["sd"]), shift, mesh.global_cpIDs, cx.el, cx.host, gf.x_star)
    for w in blk["extervars"]:
        if w["kind"] == "global":
            env[w["sym"]] = gf.t if w["sym"] == "t" else gf.dt if w["sym"] == "dt" else dom.global_vars[w["sym"]]
        elif w["kind"] == "cp":
            env[w["sym"]] = cpath.var_basic(cx.itp, sd_slot(w["sd"]), 0, mesh.controlpoint_IDs, cx.el, cx.host,
                                            np.ascontiguousarray(dom.cp[w["local"]]))
        elif w["kind"] == "normal":
            env[w["sym"]] = cx.normals[:, w["c"] - 1, :]
    if which != "linear":
        # INTEGRATION_POINT_VAR outputs: `(ep1, ..., ep6) = Main.strain_updater(e1_1, ...)` on whole arrays
        # (08_Tensor.jl:175-183); oracle arrays are [n_sel, n_q] = the reference's column-major [n_q, n_sel]
        for c in blk.get("qp_calls", []):
            args = [np.ascontiguousarray(_eval(a, env, cx.w.shape)) for a in c["args"]]
            outs = dom.callbacks[c["func"]](*args)
            for n, v in zip(c["outs"], outs):
                env[n] = v
    return env


def _eval(expr, env, shape):
    v = eval(expr, {"__builtins__": {}}, env)
    return np.broadcast_to(np.asarray(v, dtype=np.float64), shape)


def _temps(blk, env, needed_inner):
    for t in blk["temps"]:
        try:
            env[t["sym"]] = eval(t["expr"], {"__builtins__": {}}, env)
        except NameError:
            if needed_inner:
                raise


def _kval(dom, K, term, vals, cx):
    """_Kval_Basic (06_FEM_Kernel.jl:28-45); loop nest in c/refpath.c."""
    mesh = dom.mesh
    m = dom.spec["sparse_mapping"].index([term["dual_pos"], term["deriv_pos"]])
    shift = m * mesh.sparse_unitsize
    cpath.kval_basic(cx.itp, sd_slot(term["dual_sd"]), sd_slot(term["deriv_sd"]), vals, mesh.sparse_IDs_by_el, shift,
                     cx.el, cx.host, K)


def _kval_numpy(dom, K, term, vals, cx):
    """Same contraction with numpy (cross-check of the C loop nest in tests)."""
    mesh = dom.mesh
    m = dom.spec["sparse_mapping"].index([term["dual_pos"], term["deriv_pos"]])
    shift = m * mesh.sparse_unitsize
    iv = cx.itp[cx.host]
    Ke = np.einsum("eaq,ebq,eq->abe", iv[:, sd_slot(term["dual_sd"])], iv[:, sd_slot(term["deriv_sd"])], vals)
    np.add.at(K, mesh.sparse_IDs_by_el[:, :, cx.el] - 1 + shift, Ke)


def K_linear_func(dom):
    """update_K_Linear_N (:265-276, body gen_K_Linear_GPU :52-91)."""
    gf = dom.globalfield
    gf.K_linear[:] = 0.0
    for blk in dom.spec["blocks"]:
        if not blk["linear_gradients"]:
            continue
        cx = _block_context(dom, blk)
        env = _declare_vars(dom, blk, cx, "linear")
        _temps(blk, env, False)
        for term in blk["linear_gradients"]:
            vals = _eval(term["expr"], env, cx.w.shape) * dom.K_params[term["deriv_td"]] * cx.w
            _kval(dom, gf.K_linear, term, vals, cx)


def K_nonlinear_func(dom):
    """update_K_NonLinear_N (:278-288, body gen_Res_K_NonLinear_GPU :93-154)."""
    gf, mesh = dom.globalfield, dom.mesh
    N = mesh.variable_size
    gf.residue[:] = 0.0
    gf.K_total[:] = gf.K_linear
    for blk in dom.spec["blocks"]:
        cx = _block_context(dom, blk)
        env = _declare_vars(dom, blk, cx, "nonlinear")
        _temps(blk, env, True)
        for term in blk["residues"]:
            vals = _eval(term["expr"], env, cx.w.shape) * cx.w
            cpath.res_basic(cx.itp, sd_slot(term["dual_sd"]), vals, term["dual_pos"] * N, mesh.global_cpIDs,
                            cx.el, cx.host, gf.residue)                                  # _Res_Basic :65-79
        for term in blk["nonlinear_gradients"]:
            vals = _eval(term["expr"], env, cx.w.shape) * dom.K_params[term["deriv_td"]] * cx.w
            _kval(dom, gf.K_total, term, vals, cx)


def csr_operator(gf, K=None):
    """K[K_val_ids] as an OpenMP CSR operator (CPU baseline)."""
    K = gf.K_total if K is None else K
    return cpath.CsrOperator(gf.K_J_ptr, gf.K_J, K[gf.K_val_ids - 1], gf.basicfield_size)


def csr_from_globalfield(gf, K=None):
    """K_total[K_val_ids] as scipy CSR (02_Preconditioner.jl:35-36)."""
    import scipy.sparse as sps
    K = gf.K_total if K is None else K
    n = gf.basicfield_size
    return sps.csr_matrix((K[gf.K_val_ids - 1], gf.K_J - 1, gf.K_J_ptr - 1), shape=(n, n))
