import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import metafem_b200 as m
from helpers import build_case, product_from_oracle, rel, j2_states
from oracle import assembly as oasm, solver as osv
for name, n in (("neo_hookean", (2, 2, 1)), ("thermal", (2, 1, 1)), ("thermo_elasticity", (2, 1, 1)), ("j2", (2, 1, 1))):
    dom, spec, mesh = build_case(name, n)
    oasm.assemble_Global_Variables(dom)
    fd = product_from_oracle(dom)
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
    if name == "j2":
        j2_states(dom, fd)
    osv.update_Time(dom); osv.initialize_dx(dom); oasm.K_linear_func(dom); osv.update_x_star(dom); oasm.K_nonlinear_func(dom)
    td = fd.time_discretization
    m.api.update_Time(fd.globalfield, td)
    gam, al = np.array(td.gamma_params), np.array(td.alpha_params)
    fd.ctx.call("mfb_initialize_dx", fd.globalfield.dt, m.lib.ptr(gam), len(gam))
    fd.K_linear_func(td, fem_domain=fd)
    fd.ctx.call("mfb_update_x_star", m.lib.ptr(al), len(al))
    fd.K_nonlinear_func(td, fem_domain=fd)
    gf = dom.globalfield
    print(name, rel(fd.get_vector(m.lib.VEC_RESIDUE), gf.residue), rel(fd.get_matrix(m.lib.MAT_K_TOTAL), gf.K_total[gf.K_val_ids - 1]))
    m.iterative_Solve(fd, Sv_func="bicgstabl_GS", maxiter=50, max_pass=1, s=2)
    m.iterative_Solve(fd, Sv_func="lsqr", maxiter=5, max_pass=1)
    fd.close()
print("SAN_CASE_DONE")
