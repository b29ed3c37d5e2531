"""Shared case builders for the parity tests: the ORACLE builds the mesh tables (reference numbering,
hash-order mid-edge nodes) and evaluates the reference algorithm; the same arrays are handed to the
product through the C ABI."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refgeom as rg, femmesh as fm, assembly as oasm, solver as osv  # noqa: E402


def box_faces(m, size):
    """Boundary face IDs of a box by side, the way the example scripts select them (static_Neo_Hookean.jl:19-34)."""
    fids = rg.get_BoundaryMesh(m)
    cen = rg.face_centroids(m, fids)
    eps = 1e-6 * max(size)
    out = {}
    for d, name_lo, name_hi in ((0, "left", "right"), (1, "front", "back"), (2, "bottom", "top")):
        out[name_lo] = fids[np.abs(cen[d]) < eps]
        out[name_hi] = fids[np.abs(cen[d] - size[d]) < eps]
    out["all"] = fids
    return out


def tables_from_oracle_mesh(mesh):
    import metafem_b200 as m
    sp = mesh.space
    return m.MeshTables(
        controlpoint_IDs=mesh.controlpoint_IDs, x=mesh.x, ref_itp_vals=sp.ref_itp_vals, itg_weight=sp.itg_weight,
        bdy_ref_itp_vals=np.stack(sp.bdy_ref_itp_vals, axis=-1), bdy_itg_weights=np.stack(sp.bdy_itg_weights, axis=-1),
        bdy_tangent_directions=np.stack(sp.bdy_tangent_directions, axis=-1),
        facet_element_ID=mesh.facet_element_ID, facet_element_eindex=mesh.facet_element_eindex, bg_fIDs=mesh.bg_fIDs)


@functools.lru_cache(maxsize=None)
def spec_for(name):
    import metafem_b200  # noqa: F401  (registers metafem_jl_b200)
    from metafem_jl_b200.frontend import weakform as wf
    if name == "thermal":
        return wf.thermal_conduction()
    if name == "linear_elasticity":
        E, nu = 1.0, 0.3
        lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
        return wf.linear_elasticity(lam, mu, 1000.0 * E, fixed_bg=1, traction_bgs=((2, "sl"),))
    if name == "neo_hookean":
        return wf.neo_hookean(fixed_bg=1, traction_bg=2)
    if name == "thermo_elasticity":
        return wf.thermo_elasticity(fixed_bg=1, thermal_bg=3)
    if name == "j2":
        return wf.j2_plasticity(fixed_bg=1, traction_bg=2)
    if name == "j2_fused":
        return wf.j2_plasticity(fixed_bg=1, traction_bg=2, fused=True)
    raise KeyError(name)


def build_case(name, n=(3, 2, 2), size=(1.5, 1.0, 1.0), seed=0):
    """Returns (oracle Domain, spec, oracle mesh) with fields and a perturbed state set, BEFORE assemble_Global_Variables."""
    rng = np.random.default_rng(seed)
    shape = "SIMPLEX" if name == "thermal" else "CUBE"
    c, conn = rg.make_Brick(size, n, shape)
    m = rg.construct_TotalMesh_3D(c, conn)
    faces = box_faces(m, size)
    spec = spec_for(name)
    if name == "thermal":
        bgs = [faces["all"]]
    elif name == "thermo_elasticity":
        bgs = [faces["left"], np.concatenate([faces["bottom"], faces["top"], faces["right"]]),
               np.concatenate([faces["front"], faces["back"]])]
    else:
        bgs = [faces["left"], faces["right"]]
    mesh = fm.mesh_Classical(m, bgs, shape)
    fm.update_Mesh(mesh)
    dom = oasm.Domain(mesh, spec)
    N = mesh.variable_size
    h = min(s / k for s, k in zip(size, n))
    if name == "thermal":
        dom.cp["T"][:] = 293.15 + 5.0 * np.sin(mesh.x[0]) + rng.uniform(-1e-3, 1e-3, N)
        dom.cp["s"][:] = 1600.0 + 10 * mesh.x[1]
        dom.globalfield.converge_tol = 1e-8
    elif name == "thermo_elasticity":
        dom.cp["T"][:] = 20.0 * np.cos(mesh.x[1]) + rng.uniform(-1e-3, 1e-3, N)
        dom.cp["T_t1"][:] = 0.3 * mesh.x[0]
        dom.cp["Te"][:] = 300.0 * (mesh.x[1] < 1e-9)
        for i, b in enumerate(("d1", "d2", "d3")):
            dom.cp[b][:] = 1e-3 * np.sin(1.3 * mesh.x[(i + 1) % 3] + 0.2 * i) * mesh.x[0]
            dom.cp[b + "_t1"][:] = 1e-4 * mesh.x[(i + 2) % 3]
        dom.globalfield.dt = 1.0
        dom.globalfield.converge_tol = 1e-6
    elif name in ("j2", "j2_fused"):
        # strains around 1e-3: part of the quadrature points are beyond the yield surface (Y = 100, E = 1e5)
        for i, b in enumerate(("d1", "d2", "d3")):
            dom.cp[b][:] = 4e-4 * np.sin(1.3 * mesh.x[(i + 1) % 3] + 0.2 * i) * mesh.x[0] + rng.uniform(-1e-5, 1e-5, N) * h
            dom.cp[b + "_t1"][:] = 1e-4 * mesh.x[(i + 2) % 3]
            dom.cp[b + "_t2"][:] = 1e-5 * mesh.x[i]
        dom.cp["sl1"][:] = 120.0
        dom.globalfield.dt = 1.0
        dom.globalfield.converge_tol = 1e-3
    else:
        for i, b in enumerate(("d1", "d2", "d3")):
            dom.cp[b][:] = 0.02 * np.sin(1.3 * mesh.x[(i + 1) % 3] + 0.2 * i) * mesh.x[0] + rng.uniform(-1e-3, 1e-3, N) * h
        if name == "linear_elasticity":
            dom.cp["sl1"][:] = 0.01
            dom.cp["sl6"][:] = 0.003
            dom.globalfield.converge_tol = 1e-10
        else:
            dom.global_vars.update(mu=1.0, lam=10.0, tau_b=1000.0 * 10.0)
            dom.cp["Pl1"][:] = 0.05
            dom.cp["Pl2"][:] = 0.01
            dom.globalfield.converge_tol = 1e-9
    return dom, spec, mesh


def product_from_oracle(dom, device=0):
    """FEM_Domain of the product fed with the oracle's mesh tables and the same state."""
    import metafem_b200 as m
    fd = m.FEM_Domain(tables_from_oracle_mesh(dom.mesh), dom.spec, device=device)
    for k, v in dom.cp.items():
        fd.controlpoints[k][:] = v
    fd.global_vars.update(dom.global_vars)
    fd.globalfield.converge_tol = dom.globalfield.converge_tol
    return fd


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


J2_PARAMS = dict(Y_initial=100.0, lam=0.0, mu=50e3, Eb=12.5e3, Ep=25e3, f_res=1.0)   # J2Plasticity.jl:46-50,199-203 (group 2)


def j2_states(dom, fd=None):
    """The example's strain_updater on both sides: oracle MaterialState (numpy) and the library's built-in return map."""
    from oracle import j2 as oj2
    n_el, n_q = dom.mesh.controlpoint_IDs.shape[1], dom.mesh.space.ref_itp_vals.shape[0]
    ost = oj2.MaterialState((n_el, n_q), **J2_PARAMS)
    dom.callbacks["strain_updater"] = ost
    pst = None
    if fd is not None:
        import metafem_b200 as m
        fd.global_vars.update({g: 0.0 for g in fd.spec["globals"]})     # the material state fills its own parameters
        pst = m.api.J2MaterialState(fd, **J2_PARAMS)
        fd.callbacks["strain_updater"] = pst
    return ost, pst
