"""Pins the J2 oracle (quadrature-point callback + history variables + generalized-alpha dynamics with second time
derivatives) against the reference's own known answers: the analytical load-displacement table of
examples/hypo_elastic_plasticity/J2Plasticity.jl:223-228 for a uniaxially loaded bar, reproduced with the script's
loading loop (:264-283)."""
import numpy as np
import pytest
from threadpoolctl import threadpool_limits

from helpers import build_case
from oracle import assembly as oasm, solver as osv, j2 as oj2

EY = 100e3
# (s_tests, d1_analytical, Eb, Ep): prefixes of groups 1 and 3 of the script (isotropic / kinematic hardening)
GROUPS = [([40, 80, 100, 120, 140, 180, 200, 180, 100], [4, 8, 10, 16, 22, 34, 40, 38, 30], 0.0, EY / 2),
          ([40, 100, 140, 200, 100, 0, -40, -100], [4, 10, 22, 40, 30, 20, 8, -10], EY / 2, 0.0)]


@pytest.mark.parametrize("group", [0, 1])
def test_uniaxial_j2_matches_analytical_table(group):
    s_tests, d1_ana, Eb, Ep = GROUPS[group]
    dom, spec, mesh = build_case("j2", (5, 2, 2), size=(10.0, 1.0, 1.0))
    for k in dom.cp:
        dom.cp[k][:] = 0.0
    oasm.assemble_Global_Variables(dom)
    n_el, n_q = mesh.controlpoint_IDs.shape[1], mesh.space.ref_itp_vals.shape[0]
    st = oj2.MaterialState((n_el, n_q), 100.0, 0.0, EY / 2, Eb, Ep, 1.0)          # nu = 0: lambda = 0, mu = E/2
    dom.callbacks["strain_updater"] = st
    dom.linear_solver = lambda d: osv.iterative_Solve(d, osv.bicgstabl_GS, maxiter=2000, max_pass=20, s=8)
    dom.globalfield.converge_tol, dom.globalfield.dt = 1e-3, 1.0
    right = np.abs(mesh.x[0] - 10.0) < 1e-6
    with threadpool_limits(limits=1, user_api="blas"):
        for s, ana in zip(s_tests, d1_ana):
            dom.cp["sl1"][:] = s
            for counter in range(300):
                osv.update_OneStep(dom, max_iter=3)
                oasm.dessemble_X(dom)
                st.update_States()
                if np.abs(dom.cp["d1_t1"]).max() < 1e-4:
                    break
            else:
                pytest.fail(f"load {s}: the viscous relaxation loop did not settle")
            d1 = dom.cp["d1"][right].mean()
            assert abs(d1 - ana * 1e-3) < 0.03 * abs(ana * 1e-3) + 4e-4, (s, d1, ana * 1e-3)
