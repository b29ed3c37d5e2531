"""GPU: the same weak-form-independent pins as tests/test_pins_oracle.py, on the CUDA path through the C ABI."""
import pytest

import pins
from test_pins_oracle import _fd_backend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,tol", [("neo_hookean", 1e-7), ("linear_elasticity", 1e-9), ("thermo_elasticity", 1e-9), ("j2", 1e-8),
                                      ("thermal", 1e-9)])
def test_tangent_is_the_derivative_of_the_residual_cuda(built_lib, name, tol):
    ob = _fd_backend(name, (2, 2, 1) if name != "thermal" else (2, 1, 1))
    cp = {k: v.copy() for k, v in ob.dom.cp.items()}
    be = pins.ProductBackend(ob.mesh, ob.spec, cp, dict(ob.dom.global_vars), dt=ob.dom.globalfield.dt, j2=(name == "j2"))
    try:
        err = pins.fd_tangent(be)
    finally:
        be.close()
    assert err < tol, err


def test_neo_hookean_P_and_A_against_hand_coded_closed_form_cuda(built_lib):
    er, ek = pins.neo_hookean_closed_form(pins.ProductBackend)
    assert er < 1e-12 and ek < 1e-12, (er, ek)


def test_thermo_elastic_free_expansion_is_stress_free_cuda(built_lib):
    assert pins.thermo_free_expansion(pins.ProductBackend) < 1e-12


@pytest.mark.parametrize("shape", ["CUBE", "SIMPLEX"])
def test_patch_test_linear_field_on_a_distorted_mesh_cuda(built_lib, shape):
    ratio, n_inner = pins.patch_test(pins.ProductBackend, shape)
    assert n_inner > 0 and ratio < 1e-11, (ratio, n_inner)
