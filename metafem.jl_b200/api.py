"""Host-side mirror of MetaFEM's interface for the hot path (same names, same argument meaning).

In production these bodies are the Julia overrides shown in INTEGRATION.md; Julia is absent in
this environment, so the identical call sequence is issued from Python through the C ABI.

  FEM_Domain / GlobalField / GeneralAlpha      reference src/solver/01_Types.jl:110-168, 04_Time_Domain.jl:1-8
  assemble_Global_Variables / assemble_X / dessemble_X       src/solver/03_GlobalAssembly.jl:6-75
  compile_Updater_GPU                                        src/solver/05_CodeGenerator.jl:265-291
  update_OneStep                                             src/solver/04_Time_Domain.jl:59-80
  iterative_Solve (Sv_func = "idrs" | "bicgstabl_GS")        src/solver/linear_solver/02_Preconditioner.jl:32-76
"""
import ctypes as C
import math

import numpy as np

from . import emitter
from . import lib as L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class GeneralAlpha:
    """04_Time_Domain.jl:1-8."""

    def __init__(self, dissipative=False):
        self.alpha_params = (1.0, 1.0, 1.0)
        self.gamma_params = (1.0, 1.0) if dissipative else (0.5, 0.5)
        self.beta_params = []
        self.K_params = []


class GlobalField:
    """Scalars of GlobalField (01_Types.jl:110-132); the arrays live on the device behind the context."""

    def __init__(self):
        self.max_time_level = 0
        self.basicfield_size = 0
        self.converge_tol = 0.0
        self.t = 0.0
        self.dt = 1.0
        self.nnz = 0
        self.sparse_unitsize = 0


class MeshTables:
    """The arrays the front end (mesh_Classical + Classical_Discretization) hands to the hot path.

    All in the reference's layout: 1-based int32 IDs, column-major tables (numpy arrays are passed
    Fortran-ordered).
      controlpoint_IDs (n_a, n_el); x (3, N); ref_itp_vals (n_q, n_a, 2,2,2); itg_weight (n_q,)
      bdy_ref_itp_vals (n_qb, n_a, 2,2,2, n_faces); bdy_itg_weights (n_qb, n_faces);
      bdy_tangent_directions (n_qb, 3, 2, n_faces); facet_element_ID, facet_element_eindex (n_facets,)
      bg_fIDs {bg_ID: facet IDs}
    """

    def __init__(self, controlpoint_IDs, x, ref_itp_vals, itg_weight, bdy_ref_itp_vals=None, bdy_itg_weights=None,
                 bdy_tangent_directions=None, facet_element_ID=None, facet_element_eindex=None, bg_fIDs=None):
        self.controlpoint_IDs = np.asfortranarray(controlpoint_IDs, dtype=np.int32)
        self.x = _f64(x)
        self.ref_itp_vals = np.asfortranarray(ref_itp_vals, dtype=np.float64)
        self.itg_weight = _f64(itg_weight)
        self.bdy_ref_itp_vals = None if bdy_ref_itp_vals is None else np.asfortranarray(bdy_ref_itp_vals, dtype=np.float64)
        self.bdy_itg_weights = None if bdy_itg_weights is None else np.asfortranarray(bdy_itg_weights, dtype=np.float64)
        self.bdy_tangent_directions = None if bdy_tangent_directions is None else np.asfortranarray(
            bdy_tangent_directions, dtype=np.float64)
        self.facet_element_ID = None if facet_element_ID is None else _i32(facet_element_ID)
        self.facet_element_eindex = None if facet_element_eindex is None else _i32(facet_element_eindex)
        self.bg_fIDs = dict(bg_fIDs or {})

    @property
    def variable_size(self):
        return self.x.shape[1]


class FEM_Domain:
    """One-workpiece FEM_Domain (01_Types.jl:147-168): mesh tables + kernel spec + device context."""

    def __init__(self, tables, spec, dim=3, dissipative=True, device=0):
        self.dim = dim
        self.tables = tables
        self.spec = spec
        self.time_discretization = GeneralAlpha(dissipative=dissipative)
        self.globalfield = GlobalField()
        self.global_vars = {}
        self.K_linear_func = None
        self.K_nonlinear_func = None
        self.linear_solver = None
        self.last_solve = None
        self.callbacks = {}          # quadrature-point callbacks by function name (Main.<func> in the reference)
        N = tables.variable_size
        # controlpoints.<sym> tables (host copies, like cpts.T / cpts.d1 in the example scripts)
        self.controlpoints = {}
        for b in spec["basic_vars"]:
            for td in range(spec["max_time_level"] + 1):
                self.controlpoints[b + (f"_t{td}" if td else "")] = np.zeros(N)
        for v in spec["cp_vars"]:
            self.controlpoints[v] = np.zeros(N)
        self.ctx = L.Context(device)
        t = tables
        n_a, n_el = t.controlpoint_IDs.shape
        self.ctx.call("mfb_mesh_set", n_a, n_el, N, t.ref_itp_vals.shape[0], L.ptr(t.controlpoint_IDs),
                      L.ptr(_f64(t.x[0])), L.ptr(_f64(t.x[1])), L.ptr(_f64(t.x[2])), L.ptr(t.ref_itp_vals),
                      L.ptr(t.itg_weight))
        if t.bdy_ref_itp_vals is not None and t.facet_element_ID is not None:
            self.ctx.call("mfb_facets_set", t.bdy_ref_itp_vals.shape[-1], t.bdy_ref_itp_vals.shape[0],
                          L.ptr(t.bdy_ref_itp_vals), L.ptr(t.bdy_itg_weights), L.ptr(t.bdy_tangent_directions),
                          len(t.facet_element_ID), L.ptr(t.facet_element_ID), L.ptr(t.facet_element_eindex))
            for bg, ids in t.bg_fIDs.items():
                ids = _i32(ids)
                self.ctx.call("mfb_boundary_group_set", int(bg), len(ids), L.ptr(ids))

    # -- vectors / matrices in reference layout ------------------------------------------------
    def get_vector(self, which):
        gf = self.globalfield
        n = gf.basicfield_size * (1 if which == L.VEC_RESIDUE else gf.max_time_level + 1)
        out = np.empty(n)
        self.ctx.call("mfb_vector_get", which, L.ptr(out), n)
        return out

    def set_vector(self, which, v):
        v = _f64(v)
        self.ctx.call("mfb_vector_set", which, L.ptr(v), len(v))

    def get_matrix(self, which=L.MAT_K_TOTAL):
        out = np.empty(self.globalfield.nnz)
        self.ctx.call("mfb_matrix_get", which, L.ptr(out), len(out))
        return out

    def get_pattern(self):
        gf = self.globalfield
        K_I, K_J = np.empty(gf.nnz, np.int32), np.empty(gf.nnz, np.int32)
        K_J_ptr, K_val_ids = np.empty(gf.basicfield_size + 1, np.int32), np.empty(gf.nnz, np.int32)
        self.ctx.call("mfb_pattern_get", L.ptr(K_I), L.ptr(K_J), L.ptr(K_J_ptr), L.ptr(K_val_ids))
        return K_I, K_J, K_J_ptr, K_val_ids

    def get_sparse_IDs_by_el(self, block=0):
        """elements.sparse_IDs_by_el [n_a, n_a, n_el] of one variable block: 1-based CSR position of every element node pair."""
        n_a, n_el = self.tables.controlpoint_IDs.shape
        out = np.empty((n_a, n_a, n_el), np.int32, order="F")
        self.ctx.call("mfb_sparse_ids_get", int(block), L.ptr(out))
        return out

    def sync_fields(self):
        """Push controlpoints.<CONTROLPOINT_VAR> and physics.global_vars to the device (read at call time in the reference)."""
        for v in self.spec["cp_vars"]:
            self.ctx.call("mfb_field_set", v.encode(), L.ptr(_f64(self.controlpoints[v])))
        for g in self.spec["globals"]:
            self.ctx.call("mfb_global_set", g.encode(), float(self.global_vars[g]))

    # -- INTEGRATION_POINT_VAR arrays [n_q, n_el] (reference element order) ------------------------------------
    @property
    def qp_shape(self):
        return (self.tables.controlpoint_IDs.shape[1], self.tables.ref_itp_vals.shape[0])      # C order of [n_q, n_el]

    def qp_get(self, name):
        out = np.empty(self.qp_shape)
        self.ctx.call("mfb_qp_get", name.encode(), L.ptr(out), out.size)
        return out

    def qp_set(self, name, values):
        v = _f64(np.broadcast_to(values, self.qp_shape))
        self.ctx.call("mfb_qp_set", name.encode(), L.ptr(v), v.size)

    def qp_device_ptr(self, name):
        """Device address of a library-owned integration-point array (what a Julia callback wraps with unsafe_wrap)."""
        p, n = C.c_void_p(), C.c_int64(0)
        self.ctx.call("mfb_qp_array", name.encode(), C.byref(p), C.byref(n))
        return p.value, n.value

    def close(self):
        self.ctx.close()


class HostCallback:
    """Generic quadrature-point callback running on the HOST: pulls the argument arrays, calls ``fn(*args)`` (numpy
    arrays shaped [n_el, n_q]) and pushes the returned arrays. The device-side equivalent in Julia wraps the pointers
    from mfb_qp_array as CuArrays."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, fem_domain, call):
        outs = self.fn(*[fem_domain.qp_get(n) for n in call["arg_names"]])
        for n, v in zip(call["outs"], outs):
            fem_domain.qp_set(n, v)


class J2MaterialState:
    """MaterialState of examples/hypo_elastic_plasticity/J2Plasticity.jl:76-198 on the library's built-in return map:
    calling it is ``strain_updater(e1_1, ..., e3_3)`` (iterate_stress!), ``update_States`` commits ep, b, Y."""

    def __init__(self, fem_domain, Y_initial, lam, mu, Eb, Ep, f_res, prefix="j2", func="strain_updater"):
        self.dom, self.prefix, self.Y_initial = fem_domain, prefix, float(Y_initial)
        self.lam, self.mu, self.Eb, self.Ep, self.f_res = lam, mu, Eb, Ep, f_res
        call = next(c for b in fem_domain.spec["blocks"] for c in b.get("qp_calls", []) if c["func"] == func)
        self.fused = bool(call.get("builtin"))   # the return map is inlined in the residual kernel (weak form built with fused=True)
        if self.fused:
            self.prefix = prefix = call["prefix"]
        self._e = (C.c_char_p * 6)(*[n.encode() for n in (call["outs"] if self.fused else call["arg_names"])])
        self._ep = (C.c_char_p * 6)(*[n.encode() for n in call["outs"]])
        self._n_yielded = 0
        self.reset()
        self.sync_params()

    def sync_params(self):
        if self.fused:
            p = self.prefix
            self.dom.global_vars.update({f"{p}_lam": self.lam, f"{p}_mu": self.mu, f"{p}_Eb": self.Eb, f"{p}_Ep": self.Ep,
                                         f"{p}_fres": self.f_res})

    @property
    def n_yielded(self):
        if self.fused:
            n = C.c_int64(0)
            self.dom.ctx.call("mfb_j2_yield_count", self.prefix.encode(), C.byref(n))
            return n.value
        return self._n_yielded

    @n_yielded.setter
    def n_yielded(self, v):
        self._n_yielded = v

    def reset(self):
        """ep .= 0; b .= 0; Y .= Y_initial (J2Plasticity.jl:256-262)."""
        self.dom.ctx.call("mfb_j2_init", self.prefix.encode(), self.Y_initial, self._e, self._ep)

    def __call__(self, fem_domain=None, call=None):
        if self.fused:
            return
        prm = L.J2Params(self.lam, self.mu, self.Eb, self.Ep, self.f_res)
        n = C.c_int64(0)
        self.dom.ctx.call("mfb_j2_iterate_stress", self.prefix.encode(), C.byref(prm), C.byref(n))
        self.n_yielded = n.value

    def update_States(self):
        self.dom.ctx.call("mfb_j2_update_states", self.prefix.encode())

    def state(self, what):
        """Committed state arrays [n_el, n_q]: what in ep1..6, b1..6, Y."""
        return self.dom.qp_get(f"{self.prefix}.{what}")


def comm_unique_id():
    """128-byte NCCL id (rank 0 creates it and hands it to the other ranks, e.g. with torch.distributed.broadcast)."""
    buf = (C.c_ubyte * 128)()
    rc = L.load().mfb_comm_unique_id(buf)
    if rc != 0:
        raise L.MfbError("mfb_comm_unique_id failed: NCCL could not be loaded (set MFB_NCCL_LIB)")
    return bytes(buf)


def init_distributed(fem_domain, subdomain, rank, n_ranks, id_bytes):
    """Attach a rank's FEM_Domain (built on subdomain.tables) to the NCCL communicator and register its interface."""
    ctx = fem_domain.ctx
    idbuf = (C.c_ubyte * 128).from_buffer_copy(id_bytes)
    ctx.call("mfb_comm_init", int(rank), int(n_ranks), idbuf)
    nb = _i32(subdomain.neighbors)
    counts = [len(subdomain.shared[q]) for q in subdomain.neighbors]
    offsets = np.ascontiguousarray(np.concatenate([[0], np.cumsum(counts)]), dtype=np.int64)
    shared = _i32(np.concatenate([subdomain.shared[q] for q in subdomain.neighbors])) if counts else np.zeros(1, np.int32)
    owned = np.ascontiguousarray(subdomain.owned, dtype=np.uint8)
    gids = np.ascontiguousarray(subdomain.node_l2g, dtype=np.int64)
    ctx.call("mfb_interface_set", len(nb), L.ptr(nb) if len(nb) else L.ptr(np.zeros(1, np.int32)), L.ptr(offsets),
             L.ptr(shared), L.ptr(owned), L.ptr(gids))
    fem_domain.subdomain = subdomain


def _x_slices(dom):
    spec, gf, N = dom.spec, dom.globalfield, dom.tables.variable_size
    for pos, b in enumerate(spec["basic_vars"]):
        for td in range(spec["max_time_level"] + 1):
            s = pos * N + td * gf.basicfield_size
            yield b + (f"_t{td}" if td else ""), slice(s, s + N)


def assemble_Global_Variables(fem_domain):
    """assemble_Global_Variables! (03_GlobalAssembly.jl:6-37): DOF numbering, x/dx/x_star/residue, sparsity pattern."""
    dom, spec, gf = fem_domain, fem_domain.spec, fem_domain.globalfield
    N = dom.tables.variable_size
    gf.basicfield_size = len(spec["basic_vars"]) * N
    gf.max_time_level = spec["max_time_level"]
    mapping = _i32(np.array(spec["sparse_mapping"], dtype=np.int32).reshape(-1, 2))
    nnz, unit = C.c_int64(0), C.c_int64(0)
    dom.ctx.call("mfb_pattern_build", len(spec["basic_vars"]), gf.max_time_level, len(mapping), L.ptr(mapping),
                 C.byref(nnz), C.byref(unit))
    gf.nnz, gf.sparse_unitsize = nnz.value, unit.value
    assemble_X(dom)


def assemble_X(fem_domain):
    """assemble_X! (:44-56): controlpoints.<var> -> x."""
    gf = fem_domain.globalfield
    x = np.zeros((gf.max_time_level + 1) * gf.basicfield_size)
    for sym, sl in _x_slices(fem_domain):
        x[sl] = fem_domain.controlpoints[sym]
    fem_domain.set_vector(L.VEC_X, x)


def dessemble_X(fem_domain):
    """dessemble_X! (:63-75): x -> controlpoints.<var>."""
    x = fem_domain.get_vector(L.VEC_X)
    for sym, sl in _x_slices(fem_domain):
        fem_domain.controlpoints[sym][:] = x[sl]


# (cell_type, el_cp_outer_id) of write_VTK (5_VTK.jl:26-118), keyed by (dim, n_a): hex20 serendipity, hex8, tet10, tet4
_VTK_CELLS = {(3, 20): (25, [1, 2, 4, 3, 5, 6, 8, 7, 9, 14, 10, 13, 11, 16, 12, 15, 17, 18, 20, 19]),
              (3, 8): (12, [1, 2, 4, 3, 5, 6, 8, 7]), (3, 10): (24, [1, 3, 6, 10, 2, 5, 4, 7, 8, 9]), (3, 4): (10, [1, 2, 3, 4])}


def write_VTK(fname, fem_domain, scale=1.0, shift_sym=None):
    """write_VTK(fname, wp; scale, shift_sym) (5_VTK.jl:7-157) straight from the device-resident x (dessemble_X! included):
    every inner variable at every time level, in the order of the reference's local_innervar_infos naming
    (d1, d1_t1 -> "d1_t", ...)."""
    dom = fem_domain
    spec, n_a = dom.spec, dom.tables.controlpoint_IDs.shape[0]
    cell_type, outer = _VTK_CELLS[(dom.dim, n_a)]
    names, var, lev = [], [], []
    for pos, b in enumerate(spec["basic_vars"]):
        for td in range(spec["max_time_level"] + 1):
            names.append(b + ("_" + "t" * td if td else ""))
            var.append(pos)
            lev.append(td)
    shift = -1 if shift_sym is None else spec["basic_vars"].index(f"{shift_sym}1")
    arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
    dom.ctx.call("mfb_write_vtk", str(fname).encode(), cell_type, len(outer), L.ptr(_i32(outer)), len(names), arr,
                 L.ptr(_i32(var)), L.ptr(_i32(lev)), float(scale), shift)


def compile_Updater_GPU(domain_ID, fem_domain, tpb=128):
    """compile_Updater_GPU (05_CodeGenerator.jl:265-291): emit CUDA C for every block, compile with NVRTC,
    install K_linear_func / K_nonlinear_func. Returns the generated source (the reference returns the Exprs)."""
    dom, t = fem_domain, fem_domain.tables
    n_a, n_q = t.controlpoint_IDs.shape[0], t.ref_itp_vals.shape[0]
    n_qb = t.bdy_ref_itp_vals.shape[0] if t.bdy_ref_itp_vals is not None else 0
    src, descs = emitter.emit(dom.spec, n_a, n_q, n_qb, tpb=tpb)
    arr = (L.BlockDesc * len(descs))()
    keep = []
    for i, d in enumerate(descs):
        cps = (C.c_char_p * max(len(d["cp_var_names"]), 1))(*[s.encode() for s in d["cp_var_names"]])
        gls = (C.c_char_p * max(len(d["global_names"]), 1))(*[s.encode() for s in d["global_names"]])
        keep += [cps, gls]
        arr[i].kind, arr[i].bg_ID = d["kind"], d["bg_ID"]
        arr[i].linear_kernel = d["linear_kernel"].encode() if d["linear_kernel"] else None
        arr[i].nonlinear_kernel = d["nonlinear_kernel"].encode() if d["nonlinear_kernel"] else None
        arr[i].n_cp_vars, arr[i].cp_var_names = len(d["cp_var_names"]), cps
        arr[i].n_globals, arr[i].global_names = len(d["global_names"]), gls
        arr[i].threads_per_block, arr[i].smem_bytes = d["threads_per_block"], d["smem_bytes"]
        arr[i].has_nonlinear_K = d["has_nonlinear_K"]
        qin = (C.c_char_p * max(len(d["qp_in_names"]), 1))(*[s.encode() for s in d["qp_in_names"]])
        qout = (C.c_char_p * max(len(d["qp_out_names"]), 1))(*[s.encode() for s in d["qp_out_names"]])
        keep += [qin, qout]
        arr[i].eval_kernel = d["eval_kernel"].encode() if d["eval_kernel"] else None
        arr[i].n_qp_in, arr[i].qp_in_names = len(d["qp_in_names"]), qin
        arr[i].n_qp_out, arr[i].qp_out_names = len(d["qp_out_names"]), qout
    dom.ctx.call("mfb_kernel_compile", src.encode(), len(descs), arr)

    def update_K_Linear(time_discretization, fem_domain=dom):
        kp = _f64(time_discretization.K_params)
        fem_domain.sync_fields()
        fem_domain.ctx.call("mfb_assemble_linear", L.ptr(kp), len(kp))

    def update_K_NonLinear(time_discretization, fem_domain=dom):
        kp = _f64(time_discretization.K_params)
        gf = fem_domain.globalfield
        fem_domain.sync_fields()
        all_calls = [c for b in fem_domain.spec["blocks"] for c in b.get("qp_calls", [])]
        for c in all_calls:
            if c.get("builtin"):                 # fused built-in callback: only its parameters travel (as GLOBAL_VARs)
                fem_domain.callbacks[c["func"]].sync_params()
        if any(c.get("builtin") for c in all_calls):
            fem_domain.sync_fields()
        calls = [c for c in all_calls if not c.get("builtin")]
        if calls:
            # two-phase update: argument arrays -> Main.<func> on whole arrays -> residual (08_Tensor.jl:175-183,210)
            fem_domain.ctx.call("mfb_eval_qp_args", gf.t, gf.dt)
            for c in calls:
                if c["func"] not in fem_domain.callbacks:
                    raise KeyError(f"quadrature-point callback {c['func']!r} is not defined (fem_domain.callbacks)")
                fem_domain.callbacks[c["func"]](fem_domain, c)
        fem_domain.ctx.call("mfb_assemble_nonlinear", L.ptr(kp), len(kp), gf.t, gf.dt)

    dom.K_linear_func, dom.K_nonlinear_func = update_K_Linear, update_K_NonLinear
    dom.generated_source = src
    return src


def update_Time(globalfield, td):
    """update_Time! (04_Time_Domain.jl:10-18)."""
    globalfield.t += globalfield.dt
    Lv = globalfield.max_time_level
    prod_gamma = [float(np.prod(td.gamma_params[:i])) for i in range(Lv + 1)]
    dt_params = [globalfield.dt ** i for i in range(Lv + 1)]
    td.beta_params = [1.0 / (g * d) for g, d in zip(prod_gamma, dt_params)]
    td.K_params = [a * b for a, b in zip(td.alpha_params[:Lv + 1], td.beta_params)]


def update_OneStep(time_discretization, max_iter=4, fem_domain=None, log=None):
    """update_OneStep! (04_Time_Domain.jl:59-80); vectors stay on the device."""
    dom, gf, td = fem_domain, fem_domain.globalfield, time_discretization
    update_Time(gf, td)
    gam = _f64(td.gamma_params)
    dom.ctx.call("mfb_initialize_dx", gf.dt, L.ptr(gam), len(gam))
    dom.K_linear_func(td, fem_domain=dom)
    alpha, beta = _f64(td.alpha_params), _f64(td.beta_params)
    counter = -1
    history = []
    while True:
        dom.ctx.call("mfb_update_x_star", L.ptr(alpha), len(alpha))
        dom.K_nonlinear_func(td, fem_domain=dom)
        res = C.c_double(0.0)
        dom.ctx.call("mfb_residue_norm", C.byref(res))
        counter += 1
        history.append(res.value)
        if log:
            log(f"step {counter} residue = {res.value}")
        if res.value < gf.converge_tol or counter > max_iter:
            break
        dom.linear_solver(dom)
        dom.ctx.call("mfb_update_dx", L.ptr(beta), len(beta), -1.0)      # update_dx!(globalfield, .- delta_x, ...)
    dom.ctx.call("mfb_commit_step")
    return history


_METHODS = {"idrs": L.MFB_IDRS, "bicgstabl_GS": L.MFB_BICGSTABL_GS, "bicgstabl": L.MFB_BICGSTABL, "gmres": L.MFB_GMRES,
            "cgs": L.MFB_CGS, "cgs2": L.MFB_CGS2, "tfqmr": L.MFB_TFQMR, "lsqr": L.MFB_LSQR, "idrs_original": L.MFB_IDRS_ORIGINAL}
_PR = {"Pr_Jacobi": L.PR_JACOBI, "Pr_Jacobi_column": L.PR_JACOBI_COLUMN, "Identity": L.PR_IDENTITY}
_PL = {"Identity": L.PL_IDENTITY, "Pl_Jacobi": L.PL_JACOBI, "Pl_Jacobi_row": L.PL_JACOBI_ROW, "Pl_ILU": L.PL_ILU}


def iterative_Solve(fem_domain, Sv_func="idrs", Pr_func="Pr_Jacobi", Pl_func="Identity", max_pass=4, maxiter=2000, s=4,
                    seed=1234, checkiter=200, want_delta=False, log=None):
    """iterative_Solve!(globalfield; Sv_func!, Pr_func!, Pl_func, max_pass, maxiter, s) (02_Preconditioner.jl:32-76).
    Sv_func: idrs, bicgstabl_GS, bicgstabl, gmres, cgs, cgs2, tfqmr, lsqr (a trailing "!" is accepted);
    Pr_func: Pr_Jacobi (default) | Pr_Jacobi_column (normalized_by_column = true) | Identity;
    Pl_func: Identity (default) | Pl_Jacobi | Pl_Jacobi_row (normalized_by_row = true) | Pl_ILU."""
    name = Sv_func.rstrip("!")
    if name not in _METHODS:
        raise ValueError(f"Sv_func {Sv_func!r} is not provided (supported: {', '.join(_METHODS)})")
    if Pr_func not in _PR or Pl_func not in _PL:
        raise ValueError(f"unsupported preconditioner {Pr_func!r} / {Pl_func!r}")
    dom, gf = fem_domain, fem_domain.globalfield
    info = L.SolveInfo()
    delta = np.empty(gf.basicfield_size) if want_delta else None
    rc = dom.ctx.call("mfb_krylov_solve_ex", _METHODS[name], int(s), int(maxiter), int(max_pass),
                      float(gf.converge_tol), int(seed), _PR[Pr_func], _PL[Pl_func], int(checkiter), L.ptr(delta),
                      C.byref(info))
    dom.last_solve = dict(passes=info.passes, iterations=info.iterations, spmv=info.spmv_count,
                          converged=bool(info.converged), residual=info.residual,
                          initial_residual=info.initial_residual)
    if log:
        log(f"solver {Sv_func}: initial res = {info.initial_residual}, passes = {info.passes}, "
            f"iter = {info.iterations}, res = {info.residual}" + ("" if rc == 0 else "  (not converged)"))
    return delta
