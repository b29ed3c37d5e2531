"""The weak-form planning of frontend/weakform.py (stand-in for build_WeakForm / construct_AssembleWeakform) against the
closed forms SURVEY.md §8(a) and Appendix D derive from the reference's generator: term counts per _Kval_Basic launch,
variable order, block numbering, time levels, linear / nonlinear classification."""
import metafem_b200  # noqa: F401
from metafem_jl_b200.frontend import weakform as wf


def _dom(spec):
    return spec["blocks"][0]


def test_thermal_conduction_terms():
    """3D_Script.jl:27-35 with alpha = 0: three domain launches (T_i, -0.6 T_i), one boundary launch with -25."""
    s = wf.thermal_conduction()
    assert s["basic_vars"] == ["T"] and s["max_time_level"] == 0 and s["sparse_mapping"] == [[0, 0]]
    d, b = s["blocks"]
    assert len(d["linear_gradients"]) == 3 and not d["nonlinear_gradients"] and len(d["residues"]) == 4
    assert sorted((t["dual_sd"], t["deriv_sd"], float(t["expr"])) for t in d["linear_gradients"]) == \
        [([i], [i], -0.6) for i in (1, 2, 3)]
    assert len(b["linear_gradients"]) == 1 and float(b["linear_gradients"][0]["expr"]) == -25.0
    assert [w["local"] for w in d["extervars"] if w["kind"] == "cp"] == ["s"]


def test_linear_elasticity_term_counts():
    """Isotropic elasticity: 21 K terms, 15 if nu = 0; all linear; 9 blocks in lexicographic order."""
    E = 1.0
    for nu, n_terms in ((0.3, 21), (0.0, 15)):
        lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
        s = wf.linear_elasticity(lam, mu, 1000.0, fixed_bg=1, traction_bgs=((2, "sl"),))
        d = _dom(s)
        assert len(d["linear_gradients"]) == n_terms and not d["nonlinear_gradients"]
        assert s["basic_vars"] == ["d1", "d2", "d3"]
        assert s["sparse_mapping"] == [[i, j] for i in range(3) for j in range(3)]


def test_neo_hookean_term_counts():
    """static_Neo_Hookean.jl:42-57: 9 residual terms and 81 nonlinear K terms in the domain, mu / lam / tau_b GLOBAL_VARs read at
    call time, penalty face linear, traction face residual only with the non-symmetric tensor naming Pl1..Pl9."""
    s = wf.neo_hookean(fixed_bg=1, traction_bg=2)
    d, fixed, trac = s["blocks"]
    assert len(d["residues"]) == 9 and len(d["nonlinear_gradients"]) == 81 and not d["linear_gradients"]
    assert sorted(s["globals"]) == ["lam", "mu", "tau_b"]
    assert len(fixed["linear_gradients"]) == 3 and not fixed["nonlinear_gradients"]
    assert not trac["linear_gradients"] and not trac["nonlinear_gradients"] and len(trac["residues"]) == 3
    assert {w["local"] for w in trac["extervars"] if w["kind"] == "cp"} == {f"Pl{k}" for k in range(1, 10)}
    assert {w["c"] for w in trac["extervars"] if w["kind"] == "normal"} == {1, 2, 3}


def test_thermo_elasticity_blocks_and_levels():
    """themal_hypo_elasticity.jl:58-71: variable order is the string sort [T, d1, d2, d3] (02_LocalAssembly.jl:93); the dual
    eps_ij contains T, so all 16 blocks are populated; first time derivatives -> max_time_level 1."""
    s = wf.thermo_elasticity()
    assert s["basic_vars"] == ["T", "d1", "d2", "d3"] and s["max_time_level"] == 1
    assert s["sparse_mapping"] == [[i, j] for i in range(4) for j in range(4)]
    d = _dom(s)
    assert not d["nonlinear_gradients"]
    assert {t["deriv_td"] for t in d["linear_gradients"]} == {0, 1}          # stiffness/conduction and the d/dt terms


def test_j2_plasticity_callback_and_tangent():
    """J2Plasticity.jl:51-63: ep is an INTEGRATION_POINT_VAR produced by strain_updater(e11, e12, e13, e22, e23, e33) with zero
    variation, so the tangent is the constant elastic one (+ inertia, K_params of levels 0, 1, 2) in K_linear; nu = 0 -> 15 + 3 + 3 terms."""
    s = wf.j2_plasticity()
    assert s["max_time_level"] == 2 and s["qp_vars"] == [f"ep{k}" for k in range(1, 7)]
    d = _dom(s)
    assert not d["nonlinear_gradients"]
    by_td = {td: sum(1 for t in d["linear_gradients"] if t["deriv_td"] == td) for td in (0, 1, 2)}
    assert by_td == {0: 15, 1: 3, 2: 3}
    (call,) = d["qp_calls"]
    assert call["func"] == "strain_updater" and len(call["args"]) == 6 and call["outs"] == [f"ep{k}" for k in range(1, 7)]
    assert call["args"][0] == "d1_1" and "d1_2" in call["args"][1] and "d2_1" in call["args"][1]        # e11, e12
    assert not call["builtin"]
    f = wf.j2_plasticity(fused=True)
    assert _dom(f)["qp_calls"][0]["builtin"] == "j2_return_map"
    assert sorted(f["globals"]) == ["j2_Eb", "j2_Ep", "j2_fres", "j2_lam", "j2_mu"]
