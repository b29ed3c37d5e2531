"""CPU: the oracle pinned on checks that do not go through the shared weak-form derivation (see tests/pins.py)."""
import numpy as np
import pytest

import pins
from helpers import build_case, spec_for


def _fd_backend(name, n=(2, 2, 1)):
    dom, spec, mesh = build_case(name, n, size=(1.0, 0.8, 0.5))
    cp = {k: v.copy() for k, v in dom.cp.items()}
    if name == "j2":                       # J2-elastic: small strains, nothing yields, the tangent is the elastic one
        for b in ("d1", "d2", "d3"):
            cp[b] *= 0.05
    dt = 0.5 if name in ("thermo_elasticity", "j2") else dom.globalfield.dt      # K_params = (1, 2, 4): the levels are weighted differently
    return pins.OracleBackend(mesh, spec, cp, dict(dom.global_vars), dt=dt, j2=(name == "j2"))


@pytest.mark.parametrize("name,tol", [("neo_hookean", 1e-7), ("linear_elasticity", 1e-9), ("thermo_elasticity", 1e-9), ("j2", 1e-8),
                                      ("thermal", 1e-9)])
def test_tangent_is_the_derivative_of_the_residual(name, tol):
    """K_total v = d residue / d x_star . (K_params-weighted v), central differences (Neo-Hookean: truncation O(h^2))."""
    be = _fd_backend(name, (2, 2, 1) if name != "thermal" else (2, 1, 1))
    if name == "j2":
        assert be.dom.callbacks["strain_updater"] is not None
    err = pins.fd_tangent(be)
    assert err < tol, err
    if name == "thermo_elasticity":
        assert be.levels == 2 and be.K_params == [1.0, 2.0]


def test_neo_hookean_P_and_A_against_hand_coded_closed_form():
    er, ek = pins.neo_hookean_closed_form(pins.OracleBackend)
    assert er < 1e-12 and ek < 1e-12, (er, ek)


def test_thermo_elastic_free_expansion_is_stress_free():
    assert pins.thermo_free_expansion(pins.OracleBackend) < 1e-12


@pytest.mark.parametrize("shape", ["CUBE", "SIMPLEX"])
def test_patch_test_linear_field_on_a_distorted_mesh(shape):
    ratio, n_inner = pins.patch_test(pins.OracleBackend, shape)
    assert n_inner > 0 and ratio < 1e-11, (ratio, n_inner)


def test_thermo_elastic_domain_term_count_is_29():
    """1 (T,T_t) + 3 (conduction) + 15 (d-d) + 3 (d_a;a / T) + 4 (T test term of the dual eps) + 3 (damping)."""
    dom_block = spec_for("thermo_elasticity")["blocks"][0]
    assert dom_block["kind"] == "domain"
    assert len(dom_block["linear_gradients"]) + len(dom_block["nonlinear_gradients"]) == 29
