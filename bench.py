#!/usr/bin/env python
"""bench.py -- Newton-step benchmark of the MetaFEM hot path on B200 (see DESIGN.md "Measurement").

One *step* = one Newton iteration of the reference's update_OneStep! loop (src/solver/04_Time_Domain.jl:66-77)
on the hyper-elastic box of BASELINE.json configs[2] (Neo-Hookean, hex20, 27 Gauss points):
    update_x_star! -> K_nonlinear_func (residual + tangent assembly) -> normalized_norm(residue)
    -> linear_solver (bicgstabl_GS!, right Jacobi, to the script's tolerance) -> update_dx!
Every step starts from the same seeded state, so all K timed steps do identical work.
`value`   : DOF/s with the state resident in HBM.
`e2e`     : same metric through the public API with HOST buffers: x (pinned) H2D and dx D2H inside the timed region.
`roofline`: the dominant kernel of the step (block SpMV of the Krylov loop), timed live with CUDA events.
`extra`   : at N = 1 the other BASELINE configs (linear elasticity 43^3 / idrs!, thermo-elasticity 62^3, J2 plasticity 88^3)
            each with its own Newton-step time, assembly DOF/s, SpMV GB/s and clocks; at N > 1 the J2 step on the
            partitioned mesh and `dist_check` (the distributed residual / SpMV / Newton update against the same step on ONE GPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--box", dest="n", type=int, default=None, help="elements per box side (default: the BASELINE size of the workload)")
    p.add_argument("--cpu-box", dest="cpu_n", type=int, default=24, help="elements per side of the bounded CPU-baseline sample")
    p.add_argument("--numbering", default="scattered", choices=["scattered", "sorted"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs / the partitioned J2 step")
    p.add_argument("--no-dist-check", action="store_true", help="N > 1: skip the comparison with the single-GPU step")
    p.add_argument("--assembly-only", action="store_true", help="development aid: skip the Krylov solve in every step "
                                                                "(prints timings of the assembly only; never a bench value)")
    p.add_argument("--workload", default="neo_hookean", choices=list(WORKLOADS),
                   help="development aid: run another BASELINE config as the main workload (the driver's line is configs[2])")
    p.add_argument("--ilu-only", action="store_true", help="development aid: one step, then the same system solved with Pl_ILU")
    p.add_argument("--spmv-sweep", action="store_true", help="development aid: time the SpMV tuning variants on the assembled "
                                                             "matrix and exit")
    p.add_argument("--detail-timers", action="store_true", help="time every Krylov reduction group and interface exchange inside the "
                                                                "timed region too (perturbs the stream; diagnosis only)")
    p.add_argument("--ncu", action="store_true", help="profiler capture run: exactly W warm-up steps, no e2e pass; "
                                                      "numbers printed by such a run are never bench values")
    return p.parse_args()


# ---- workload definitions: material, load, solver and tolerance of the example scripts -----------------------------------
MU, LAM, LOAD, TOL = 1e6, 1e6, 4e5, 1e-5       # examples/hyper_elasticity/static_Neo_Hookean.jl, setups[1]
SOLVER = dict(Sv_func="bicgstabl_GS", maxiter=3000, max_pass=10, s=4)
WORKLOADS = {
    # name: (BASELINE config, box side, solver of the script, converge_tol, label)
    "neo_hookean": dict(cfg=2, box=88, solver=SOLVER, tol=TOL,
                        label="examples/hyper_elasticity: Neo-Hookean Newton step, hex20 x 27 Gauss points"),
    "thermal_conduction": dict(cfg=0, box=100, solver=dict(Sv_func="idrs", maxiter=2000, max_pass=20, s=8), tol=1e-6,
                               label="examples/thermal_conduction: steady heat conduction with convection boundaries (1 field; the script's "
                                     "weak form and solver on a synthetic box instead of the bundled tet mesh), hex20"),
    "linear_elasticity": dict(cfg=1, box=43, solver=dict(Sv_func="idrs", maxiter=2000, max_pass=20, s=8), tol=1e-5,
                              label="examples/linear_elasticity: 3-D linear elasticity Newton step (cantilever script's solver), hex20"),
    "thermo_elasticity": dict(cfg=3, box=62, solver=dict(Sv_func="bicgstabl_GS", maxiter=2000, max_pass=20, s=8), tol=1e-6,
                              label="examples/thermal_elasticity: coupled thermo-mechanical Newton step (4 fields, 2 time levels), hex20"),
    "j2": dict(cfg=4, box=88, solver=dict(Sv_func="bicgstabl_GS", maxiter=2000, max_pass=20, s=8), tol=1e-3,
               label="examples/hypo_elastic_plasticity: J2 Newton step, two-phase quadrature-point callback (built-in return map), hex20"),
    "j2_fused": dict(cfg=4, box=88, solver=dict(Sv_func="bicgstabl_GS", maxiter=2000, max_pass=20, s=8), tol=1e-3,
                     label="examples/hypo_elastic_plasticity: J2 Newton step, return map fused into the residual kernel, hex20"),
}


def initial_state(x, h, seed=1234):
    """Smooth displacement + uniform noise of +-1e-3 h (keeps det F > 0); SURVEY.md §8(d)."""
    rng = np.random.default_rng(seed)
    N = x.shape[1]
    d = [0.02 * np.sin(1.3 * x[(i + 1) % 3] + 0.2 * i) * x[0] + rng.uniform(-1e-3, 1e-3, N) * h for i in range(3)]
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=int(rows[0][1]), reasons=reasons, samples=len(rows))


def ncu_traffic(prefix):
    """DRAM bytes per launch (read + write) of the kernel from the newest committed ncu --set full summary under profiles/
    (profiles/summarize.py writes them from the .ncu-rep brought back by gpurun); None if there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"{prefix}_*_summary.txt")))   # round tags sort by name
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(files[-1]):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+(\w+)\s+([0-9.]+)", line)
        if m:
            tot += float(m.group(3)) * unit.get(m.group(2), 1.0)
    return (tot or None), os.path.relpath(files[-1], ROOT)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def spmv_bytes_csr(nnz, ndof):
    """scalar CSR (the reference's cuSPARSE operand): 8 B value + 4 B column per entry, row pointer, x once, y once."""
    return 12.0 * nnz + 4.0 * (ndof + 1) + 16.0 * ndof


def spmv_bytes_block(unit, n_nodes, nv):
    """What the block-CSR kernel has to move at minimum: nv^2 values + ONE column index per node pair, node-row pointer, x, y."""
    return (8.0 * nv * nv + 4.0) * unit + 4.0 * (n_nodes + 1) + 16.0 * n_nodes * nv


# ---- CPU reference arm --------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's algorithm (oracle port: term-by-term _Var/_Res/_Kval loops over pre-tabulated integral_vals,
    CSR SpMV, bicgstabl_GS!) on the host cores, on a bounded sample of the same workload."""

    def __init__(self, n_side, threads=None):
        from oracle import refgeom as rg, femmesh as fm, assembly as oasm, solver as osv, cpath
        import metafem_b200  # noqa: F401
        from metafem_jl_b200.frontend import weakform as wf
        self.threads = threads or os.cpu_count()
        cpath.set_threads(self.threads)
        self.n_side = n_side
        size, n = (1.0, 1.0, 1.0), (n_side,) * 3
        c, conn = rg.make_Brick(size, n, "CUBE")
        m = rg.construct_TotalMesh_3D(c, conn)
        fids = rg.get_BoundaryMesh(m)
        cen = rg.face_centroids(m, fids)
        mesh = fm.mesh_Classical(m, [fids[np.abs(cen[0]) < 1e-9], fids[np.abs(cen[0] - 1.0) < 1e-9]], "CUBE")
        fm.update_Mesh(mesh)
        dom = oasm.Domain(mesh, wf.neo_hookean(fixed_bg=1, traction_bg=2))
        dom.global_vars.update(mu=MU, lam=LAM, tau_b=1000 * max(MU, LAM))
        for b, v in zip(("d1", "d2", "d3"), initial_state(mesh.x, 1.0 / n_side)):
            dom.cp[b][:] = v
        dom.cp["Pl1"][:] = LOAD
        dom.globalfield.converge_tol = TOL
        oasm.assemble_Global_Variables(dom)
        osv.update_Time(dom)
        oasm.K_linear_func(dom)
        self.dom, self.oasm, self.osv, self.cpath = dom, oasm, osv, cpath
        self.ndof = dom.globalfield.basicfield_size
        self.nnz = len(dom.globalfield.K_I)

    def step(self):
        from threadpoolctl import threadpool_limits
        dom, oasm, osv, Op = self.dom, self.oasm, self.osv, self.cpath.CsrOperator
        with threadpool_limits(limits=1, user_api="blas"):   # OpenBLAS's spinning threads fight the OpenMP team otherwise
            Op.calls, Op.seconds = 0, 0.0
            t0 = time.perf_counter()
            osv.initialize_dx(dom)
            osv.update_x_star(dom)
            oasm.K_nonlinear_func(dom)
            t_asm = time.perf_counter() - t0
            res = osv.normalized_norm(dom.globalfield.residue)
            t1 = time.perf_counter()
            delta = osv.iterative_Solve(dom, osv.bicgstabl_GS, max_pass=SOLVER["max_pass"], maxiter=SOLVER["maxiter"],
                                        s=SOLVER["s"])
            t_solve = time.perf_counter() - t1
            osv.update_dx(dom, -delta)
            t = time.perf_counter() - t0
        n_spmv, t_spmv = Op.calls, Op.seconds
        return dict(value=self.ndof / t, unit="DOF/s", cores=self.threads, kind="port",
                    sample=f"one Newton step of the same Neo-Hookean hex20 box at {self.n_side}^3 elements ({self.ndof} DOF): "
                           f"{t:.2f} s total, assembly {t_asm:.2f} s, {sum(dom.last_solve['iters'])} Krylov iterations, "
                           f"initial residual {res:.3e}",
                    sample_box=self.n_side, sample_dof=self.ndof, same_config=False,
                    # size-invariant sub-metrics (same definitions as the GPU line)
                    assembly_dof_per_s=self.ndof / t_asm, krylov_iterations=int(sum(dom.last_solve["iters"])),
                    spmv_count=int(n_spmv), spmv_gbs_csr_model=spmv_bytes_csr(self.nnz, self.ndof) * n_spmv / max(t_spmv, 1e-12) / 1e9,
                    ms_per_spmv_equivalent=t_solve * 1e3 / max(n_spmv, 1)), t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    ref = CpuReference(args.cpu_n)
    ts, cb = [], None
    for i in range(W + K):
        cb, t = ref.step()
        if i >= W:
            ts.append(t)
    ms = float(np.mean(ts)) * 1e3
    v = ref.ndof / (ms * 1e-3)
    cb["value"] = v
    print(json.dumps({"impl": "reference", "metric": "newton_step_dof_per_s", "value": v, "unit": "DOF/s",
                      "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": workload_config("neo_hookean", args.cpu_n, args.numbering,
                                                extra={"dof": ref.ndof, "same_config_as_gpu_arm": False,
                                                       "note": "bounded sample: the CPU port of the reference path cannot hold the 88^3 "
                                                               "box in a few minutes; the GPU line carries cpu_baseline.gpu_same_sample "
                                                               "(the CUDA path on THIS box size) and size-invariant sub-metrics"}),
                      "cpu_baseline": cb, "assembly_dof_per_s": cb["assembly_dof_per_s"],
                      "ms_per_spmv_equivalent": cb["ms_per_spmv_equivalent"], "spmv_gbs_csr_model": cb["spmv_gbs_csr_model"],
                      "e2e": {"value": v, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(name, n, numbering, extra=None):
    w = WORKLOADS[name]
    s = w["solver"]
    cfg = {"workload": f"{w['label']}, unit box {n}^3 elements (BASELINE configs[{w['cfg']}]; full size {w['box']}^3)",
           "converge_tol": w["tol"],
           "solver": f"{s['Sv_func']} s={s['s']} maxiter={s['maxiter']} max_pass={s['max_pass']}, right Jacobi",
           "midedge_numbering": numbering,
           "cache": "inputs larger than L2 (matrix values alone exceed 126 MB); no explicit flush"}
    if name == "neo_hookean":
        cfg.update({"mu": MU, "lambda": LAM, "traction": LOAD})
    if extra:
        cfg.update(extra)
    return cfg


# ---- the CUDA path ------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide handles of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import metafem_b200 as m
        self.torch, self.dist, self.m, self.L = torch, dist, m, m.lib
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.stream = torch.cuda.Stream()

    def comm_id(self):
        torch, dist = self.torch, self.dist
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            idt.copy_(torch.frombuffer(bytearray(self.m.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        torch.cuda.synchronize()
        dist.barrier()            # torch's own communicator is fully up before the library creates its own
        return bytes(idt.cpu().numpy().tobytes())


def make_spec(name):
    from metafem_jl_b200.frontend import weakform as wf
    return {"neo_hookean": lambda: wf.neo_hookean(fixed_bg=1, traction_bg=2),
            "thermal_conduction": lambda: wf.thermal_conduction(bgs=(1, 2)),          # 3D_Script.jl:21-35, convection on both ends
            # E = 1e6, nu = 0.3 (lambda, mu = 0.5769e6, 0.3846e6), penalty 1000 E as in the cantilever script (3D_Script.jl:45-63)
            "linear_elasticity": lambda: wf.linear_elasticity(0.5769e6, 0.3846e6, 1e9, fixed_bg=1, traction_bgs=((2, "sl"),)),
            "thermo_elasticity": lambda: wf.thermo_elasticity(fixed_bg=1, thermal_bg=2),
            "j2": lambda: wf.j2_plasticity(fixed_bg=1, traction_bg=2),
            "j2_fused": lambda: wf.j2_plasticity(fixed_bg=1, traction_bg=2, fused=True)}[name]()


class Case:
    """One workload set up on the device (all ranks when partitioned=True, else this process alone on the full mesh)."""

    def __init__(self, env, name, n, numbering, partitioned=True):
        m, L = env.m, env.L
        from metafem_jl_b200.frontend import mesh as fmesh
        t0 = time.perf_counter()
        self.env, self.name, self.n = env, name, n
        self.world = env.world if partitioned else 1
        gt = fmesh.box_tables((1.0, 1.0, 1.0), (n, n, n), "CUBE", groups=("left", "right"), numbering=numbering)
        spec = make_spec(name)
        self.ndof_global = len(spec["basic_vars"]) * gt.variable_size
        self.n_nodes_global = gt.variable_size
        self.n_el_global = gt.controlpoint_IDs.shape[1]
        gx = gt.x
        scale = 1.0 if name in ("neo_hookean", "linear_elasticity") else (4e-4 / 0.02 * 4.0 if name.startswith("j2") else 0.05)
        gstate = [v * scale for v in initial_state(gx, 1.0 / n)]
        if self.world > 1:
            from metafem_jl_b200.frontend import partition as pt
            self.pt = pt
            self.sub = pt.make_subdomains(gt, pt.split_elements(gt, self.world), ranks=[env.rank])[env.rank]
            tables = self.sub.tables
            pick = lambda v: pt.scatter_field(self.sub, v)
        else:
            self.sub, tables, pick = None, gt, (lambda v: v)
        self.gx = gx if not partitioned or env.world == 1 or env.rank == 0 else None
        del gt
        self.tables = tables
        fd = self.fd = m.FEM_Domain(tables, spec, device=env.local)
        fd.ctx.call("mfb_set_stream", L.ptr(env.stream.cuda_stream))
        if self.world > 1:
            cid = env.comm_id()
            with env.torch.cuda.stream(env.stream):
                m.init_distributed(fd, self.sub, env.rank, env.world, cid)
        if name != "thermal_conduction":
            for b, v in zip(("d1", "d2", "d3"), gstate):
                fd.controlpoints[b][:] = pick(v)
        lx = tables.x
        if name == "thermal_conduction":
            fd.controlpoints["T"][:] = 293.15 + 20.0 * np.cos(3.0 * lx[1]) * lx[0]
            fd.controlpoints["s"][:] = 1000.0
        elif name == "neo_hookean":
            fd.controlpoints["Pl1"][:] = LOAD
            fd.global_vars.update(mu=MU, lam=LAM, tau_b=1000 * max(MU, LAM))
        elif name == "thermo_elasticity":
            fd.controlpoints["T"][:] = 20.0 * np.cos(lx[1])
            fd.controlpoints["Te"][:] = 300.0
            fd.globalfield.dt = 1.0
        elif name in ("j2", "j2_fused"):
            fd.controlpoints["sl1"][:] = 120.0
            fd.globalfield.dt = 1.0
        else:
            fd.controlpoints["sl1"][:] = 1e4
        fd.globalfield.converge_tol = WORKLOADS[name]["tol"]
        m.assemble_Global_Variables(fd)
        m.compile_Updater_GPU(1, fd)
        self.j2 = None
        if name in ("j2", "j2_fused"):
            fd.global_vars.update({g: 0.0 for g in spec["globals"]})
            self.j2 = m.api.J2MaterialState(fd, Y_initial=100.0, lam=0.0, mu=50e3, Eb=12.5e3, Ep=25e3, f_res=1.0)
            fd.callbacks["strain_updater"] = self.j2
            fd.sync_fields()
        gf, td = fd.globalfield, fd.time_discretization
        m.api.update_Time(gf, td)
        self.gf, self.td = gf, td
        self.kp, self.alpha = np.array(td.K_params), np.array(td.alpha_params)
        self.beta, self.gam = np.array(td.beta_params), np.array(td.gamma_params)
        fd.sync_fields()
        self.nv = len(spec["basic_vars"])
        self.ndof, self.nnz = gf.basicfield_size, gf.nnz
        self.nx = (gf.max_time_level + 1) * self.ndof
        self.x_host = env.torch.empty(self.nx, dtype=env.torch.float64).pin_memory()
        self.x_host.numpy()[:] = fd.get_vector(L.VEC_X)
        self.dx_host = env.torch.empty(self.nx, dtype=env.torch.float64).pin_memory()
        self.info, self.res = L.SolveInfo(), C.c_double(0.0)
        self.history = []
        self.setup_s = time.perf_counter() - t0
        # K_linear_func: once per time step, outside the Newton loop (04_Time_Domain.jl:64)
        self.k_linear_ms = self.time_call(lambda: fd.ctx.call("mfb_assemble_linear", L.ptr(self.kp), len(self.kp)))

    def time_call(self, fn):
        torch, s = self.env.torch, self.env.stream
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def step(self, e2e=False, solve=True):
        fd, L, gf = self.fd, self.env.L, self.gf
        call = fd.ctx.call
        if e2e:
            call("mfb_vector_set", L.VEC_X, L.ptr(self.x_host), self.nx)
        call("mfb_initialize_dx", gf.dt, L.ptr(self.gam), len(self.gam))
        call("mfb_update_x_star", L.ptr(self.alpha), len(self.alpha))
        if self.name == "j2":
            call("mfb_eval_qp_args", gf.t, gf.dt)
            self.j2()
        call("mfb_assemble_nonlinear", L.ptr(self.kp), len(self.kp), gf.t, gf.dt)
        call("mfb_residue_norm", C.byref(self.res))
        if not solve:
            return
        s = WORKLOADS[self.name]["solver"]
        call("mfb_krylov_solve", {"bicgstabl_GS": L.MFB_BICGSTABL_GS, "idrs": L.MFB_IDRS}[s["Sv_func"]], s["s"], s["maxiter"],
             s["max_pass"], WORKLOADS[self.name]["tol"], 1234, None, C.byref(self.info))
        i = self.info
        self.history.append((i.iterations, i.spmv_count, i.passes, bool(i.converged), i.residual))
        call("mfb_update_dx", L.ptr(self.beta), len(self.beta), -1.0)
        if e2e:
            call("mfb_vector_get", L.VEC_DX, L.ptr(self.dx_host), self.nx)

    def timed(self, steps, e2e=False, solve=True):
        env = self.env
        torch, dist = env.torch, env.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(env.stream):
            e0.record()
            for _ in range(steps):
                self.step(e2e, solve)
            e1.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def ilu_solve(self):
        """The same linear system (matrix and residual of the last step) solved with the script's bicgstabl_GS! + Pl_ILU: what the
        incomplete factorisation does to the iteration count and to the solve time (extra, not the headline)."""
        L, s = self.env.L, WORKLOADS[self.name]["solver"]
        defect, levels = C.c_double(0.0), C.c_int32(0)
        self.fd.ctx.call("mfb_ilu_selftest", C.byref(defect), C.byref(levels), None, 0)
        info, info1 = L.SolveInfo(), L.SolveInfo()
        # factorisation + one outer iteration (2 s products): what a solve costs before it iterates
        ms1 = self.time_call(lambda: self.fd.ctx.call("mfb_krylov_solve_ex", L.MFB_BICGSTABL_GS, s["s"], 1, 1,
                                                      WORKLOADS[self.name]["tol"], 1234, L.PR_JACOBI, L.PL_ILU, 200, None, C.byref(info1)))
        self.profile(1)
        self.profile_get()
        ms = self.time_call(lambda: self.fd.ctx.call("mfb_krylov_solve_ex", L.MFB_BICGSTABL_GS, s["s"], s["maxiter"], s["max_pass"],
                                                     WORKLOADS[self.name]["tol"], 1234, L.PR_JACOBI, L.PL_ILU, 200, None, C.byref(info)))
        pms, pcnt = self.profile_get()
        return {"factorisation_plus_first_iteration_ms": ms1, "first_iteration_products": int(info1.spmv_count),
                "sweeps_ms_per_product": pms[7] / max(pcnt[7], 1), "spmv_ms": pms[0] / max(pcnt[0], 1), "applications": int(pcnt[7]),"solver": f"bicgstabl_GS s={s['s']}, right Jacobi + left Pl_ILU (block ILU(0), elimination by colour classes)",
                "solve_ms": ms, "krylov_iterations": int(info.iterations), "spmv_count": int(info.spmv_count), "passes": int(info.passes),
                "converged": bool(info.converged), "final_residual": info.residual, "dependency_levels": int(levels.value),
                "factorisation_defect_rel": defect.value,
                "note": "solve_ms includes the factorisation; every product is followed by 2 x dependency_levels - 1 sweep launches"}

    def profile(self, level):
        self.fd.ctx.call("mfb_profile_enable", level)

    def profile_get(self):
        pms, pcnt = (C.c_double * 8)(), (C.c_int64 * 8)()
        self.fd.ctx.call("mfb_profile_get", pms, pcnt)
        return list(pms), list(pcnt)

    def launches(self):
        return self.fd.ctx.lib.mfb_launch_count(self.fd.ctx.h)

    def close(self):
        self.fd.close()


def iteration_stats(hist):
    it = np.array([h[0] for h in hist], dtype=float)
    sp = np.array([h[1] for h in hist], dtype=float)
    return {"krylov_iterations": {"mean": float(it.mean()), "min": int(it.min()), "max": int(it.max())},
            "spmv": {"mean": float(sp.mean()), "min": int(sp.min()), "max": int(sp.max()), "total": int(sp.sum())},
            "passes_max": int(max(h[2] for h in hist)), "all_converged": bool(all(h[3] for h in hist)),
            "final_residual_max": float(max(h[4] for h in hist))}


def measure(case, K, W, e2e=True, clocks=True, solve=True, level=1):
    """W warm-up steps, K timed steps resident (+ K end to end); returns the sub-metrics every workload reports."""
    env = case.env
    with env.torch.cuda.stream(env.stream):
        for _ in range(W):
            case.step(False, solve)
    env.torch.cuda.synchronize()
    cs = ClockSampler(env.local).start() if clocks else None
    case.history.clear()
    l0 = case.launches()
    case.profile(level)
    ms_total = case.timed(K, False, solve)
    launches = case.launches() - l0
    pms, pcnt = case.profile_get()
    case.profile(0)
    hist = list(case.history)
    ms_e2e = case.timed(K, True, solve) if e2e else None
    clk = cs.stop() if cs else None
    ms_step = ms_total / K
    spmv_ms = pms[0] / max(pcnt[0], 1)
    asm_ms = pms[1] / max(pcnt[1], 1)
    solve_ms = pms[3] / max(pcnt[3], 1)
    out = {"dof": case.ndof_global, "ms_per_step": ms_step, "value": case.ndof_global / (ms_step * 1e-3), "unit": "DOF/s",
           "assembly_ms": asm_ms, "assembly_dof_per_s": case.ndof_global / (asm_ms * 1e-3) if asm_ms else None,
           "element_kernel_ms": pms[4] / max(pcnt[4], 1), "K_linear_ms": case.k_linear_ms, "gpu_launches": int(launches),
           "setup_s": case.setup_s}
    if solve and hist:
        st = iteration_stats(hist)
        n_sp = st["spmv"]["total"]
        b_csr = spmv_bytes_csr(case.nnz, case.ndof)
        b_blk = spmv_bytes_block(case.gf.sparse_unitsize, case.tables.variable_size, case.nv)
        out.update(st)
        out.update({"solve_ms": solve_ms, "spmv_ms": spmv_ms, "spmv_share_of_step": pms[0] / ms_total,
                    "ms_per_spmv_equivalent": pms[3] / max(n_sp, 1), "non_spmv_ms_per_spmv": (pms[3] - pms[0]) / max(n_sp, 1),
                    "launches_per_spmv": launches / max(n_sp, 1),
                    "spmv_gbs_csr_model": b_csr / (spmv_ms * 1e-3) / 1e9, "spmv_gbs_block_format": b_blk / (spmv_ms * 1e-3) / 1e9,
                    "spmv_launches_timed": int(pcnt[0]), "spmv_bytes_csr_model": b_csr, "spmv_bytes_block_format": b_blk})
    if ms_e2e is not None:
        out["e2e"] = {"value": case.ndof_global / (ms_e2e / K * 1e-3), "unit": "DOF/s", "h2d_bytes_per_step": 8 * case.nx,
                      "d2h_bytes_per_step": 8 * case.nx + 8, "ms_per_step": ms_e2e / K}
    if clk is not None:
        out["clocks"] = clk
    if case.j2 is not None:
        ny = env.torch.tensor([float(case.j2.n_yielded)], device="cuda")
        if case.world > 1:
            env.dist.all_reduce(ny)
        out["yielded_points"] = int(ny.item())
    out["initial_residual"] = case.res.value
    return out


def gather_owned(case, local_vec, nv):
    """Reference-layout local vector (variable-major [nv, N_local]) -> global [nv, N] on every rank (owned entries only)."""
    env, sub = case.env, case.sub
    torch = env.torch
    g = np.zeros((nv, case.n_nodes_global))
    own = sub.owned.astype(bool)
    lv = np.asarray(local_vec).reshape(nv, -1)
    g[:, sub.node_l2g[own] - 1] = lv[:, own]
    t = torch.from_numpy(g).cuda()
    env.dist.all_reduce(t)
    out = t.cpu().numpy()
    # interface copies: every rank's value minus the owner's
    mism = float(np.abs(lv - out[:, sub.node_l2g - 1]).max()) if lv.size else 0.0
    tm = torch.tensor([mism], device="cuda")
    env.dist.all_reduce(tm, op=env.dist.ReduceOp.MAX)
    return out, float(tm.item())


def dist_check(env, case, args):
    """The partitioned step against the SAME step on one GPU (rank 0 holds the undivided mesh): assembled residual, one SpMV with
    interface completion, and the Newton update dx after the full Krylov solve. Interface copies must be bit-identical."""
    L, fd = env.L, case.fd
    nv, N = case.nv, case.n_nodes_global
    with env.torch.cuda.stream(env.stream):
        case.step(False, True)
    env.torch.cuda.synchronize()
    dx_d, mis_dx = gather_owned(case, fd.get_vector(L.VEC_DX)[:case.ndof], nv)
    res_d, mis_res = gather_owned(case, fd.get_vector(L.VEC_RESIDUE), nv)
    # deterministic global test vector from the global node id (no coordinates needed on the other ranks)
    ids = np.arange(1, N + 1, dtype=np.float64)
    gv = np.stack([np.sin(0.37 * ids + 0.9 * k) for k in range(nv)])
    lv = np.ascontiguousarray(gv[:, case.sub.node_l2g - 1].ravel())
    ly = np.empty_like(lv)
    fd.ctx.call("mfb_spmv", L.MAT_K_TOTAL, L.ptr(lv), L.ptr(ly), len(lv))
    y_d, mis_y = gather_owned(case, ly, nv)
    iters_d = case.info.iterations
    out = None
    if env.rank == 0:
        single = Case(env, case.name, case.n, args.numbering, partitioned=False)
        with env.torch.cuda.stream(env.stream):
            single.step(False, True)
        env.torch.cuda.synchronize()
        sf = single.fd
        dx_s = sf.get_vector(L.VEC_DX)[:single.ndof].reshape(nv, N)
        res_s = sf.get_vector(L.VEC_RESIDUE).reshape(nv, N)
        ys = np.empty(nv * N)
        gvf = np.ascontiguousarray(gv.ravel())
        sf.ctx.call("mfb_spmv", L.MAT_K_TOTAL, L.ptr(gvf), L.ptr(ys), len(gvf))
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        # both solves stop at ||residue - K delta|| / sqrt(n) <= tol (ABSOLUTE, 02_Preconditioner.jl:45-48), so the two Newton updates
        # differ by a vector whose image under K is at most 2 tol: that is the rigorous bound (the relative difference is reported too)
        diff = np.ascontiguousarray((dx_d - dx_s).ravel())
        kd = np.empty_like(diff)
        sf.ctx.call("mfb_spmv", L.MAT_K_TOTAL, L.ptr(diff), L.ptr(kd), len(diff))
        tol = WORKLOADS[case.name]["tol"]
        out = {"rel_residual_vs_single": rel(res_d, res_s), "rel_spmv_vs_single": rel(y_d, ys.reshape(nv, N)),
               "rel_dx_vs_single": rel(dx_d, dx_s), "K_times_dx_difference_normalized": float(np.linalg.norm(kd) / np.sqrt(len(kd))),
               "interface_mismatch": max(mis_dx, mis_res, mis_y),
               "initial_residual": case.res.value, "initial_residual_single": single.res.value,
               "krylov_iterations": int(iters_d), "krylov_iterations_single": int(single.info.iterations),
               "tolerances": {"residual": 1e-11, "spmv": 1e-11, "K_times_dx_difference_normalized": 2.2 * tol, "interface_mismatch": 0.0}}
        out["ok"] = bool(out["rel_residual_vs_single"] <= 1e-11 and out["rel_spmv_vs_single"] <= 1e-11
                         and out["K_times_dx_difference_normalized"] <= 2.2 * tol and out["interface_mismatch"] == 0.0)
        single.close()
    if env.world > 1:
        env.dist.barrier()
    return out


def fp64_peak_with_clocks(env, case):
    """FP64 FMA peak probe (register-resident DFMA chains) with the SM clock sampled WHILE it runs."""
    cs = ClockSampler(env.local).start()
    best, v = 0.0, C.c_double(0.0)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.2:
        case.fd.ctx.call("mfb_measure_fp64_peak", C.byref(v))
        best = max(best, v.value)
    return best, cs.stop()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    env = Env()
    torch, dist, m, L = env.torch, env.dist, env.m, env.L
    rank, world = env.rank, env.world
    K, W = args.steps, max(args.warmup, 0)
    if not args.ncu:
        W = max(W, 3)
    name = args.workload
    n = args.n or WORKLOADS[name]["box"]
    case = Case(env, name, n, args.numbering)
    fd, gf = case.fd, case.gf

    if args.spmv_sweep:
        with torch.cuda.stream(env.stream):
            case.step(False, False)
        out = {}
        for rnd in range(2):
            for v in ([int(x) for x in os.environ['MFB_SWEEP_VARIANTS'].split(',')] if os.environ.get('MFB_SWEEP_VARIANTS') else range(0, 32)):
                ms, diff = C.c_double(0.0), C.c_double(0.0)
                try:
                    fd.ctx.call("mfb_spmv_variant_bench", v, 20, C.byref(ms), C.byref(diff) if rnd == 0 else None)
                except m.lib.MfbError as e:
                    out.setdefault(v, []).append(f"error: {e}"[:120])
                    continue
                out.setdefault(v, []).append(round(ms.value, 4))
                if rnd == 0:
                    out[v].append(f"maxdiff={diff.value:.2e}")
        print(json.dumps({"spmv_sweep_ms": out, "workload": name, "box": n}))
        return
    if args.ilu_only:
        with torch.cuda.stream(env.stream):
            case.step(False, True)
        torch.cuda.synchronize()
        out = [case.ilu_solve() for _ in range(2)]
        print(json.dumps({"ilu_only": out, "order": os.environ.get("MFB_ILU_ORDER", "color"), "sweep": os.environ.get("MFB_ILU_SWEEP", "stream"), "cfg": os.environ.get("MFB_ILU_CFG", "0"), "workload": name, "box": n}))
        return
    if args.assembly_only:
        r = measure(case, K, W, e2e=False, solve=False)
        if rank == 0:
            print(json.dumps({"assembly_only": True, "workload": name, "box": n, **r}))
        return

    r = measure(case, K, W, e2e=not args.ncu, level=2 if args.detail_timers else 1)
    # ---- N > 1: limiter breakdown from ONE extra step with the per-exchange / per-reduction timers on --------------------
    breakdown = None
    if world > 1 and not args.ncu:
        case.profile(2)
        case.history.clear()
        ms1 = case.timed(1)
        pms, pcnt = case.profile_get()
        case.profile(0)
        nsp = max(case.history[-1][1], 1)
        breakdown = {"note": "one extra step with per-exchange / per-reduction CUDA events (they perturb the stream: not part of the timed region)",
                     "step_ms": ms1, "halo_exchange_ms_per_step": pms[5], "halo_exchanges": int(pcnt[5]),
                     "krylov_reductions_ms_per_step": pms[6], "reductions": int(pcnt[6]), "spmv_ms_per_step": pms[0],
                     "halo_us_per_spmv": pms[5] * 1e3 / nsp, "reductions_us_per_spmv": pms[6] * 1e3 / nsp}
    dc = None
    if world > 1 and not args.no_dist_check and not args.ncu and name == "neo_hookean":
        dc = dist_check(env, case, args)
    ilu = None
    if world == 1 and not args.no_extras and not args.ncu:
        try:
            ilu = case.ilu_solve()
            ilu["jacobi_only"] = {"solve_ms": r["solve_ms"], "krylov_iterations": r["krylov_iterations"]["mean"]}
            ilu["newton_step_ms_with_pl_ilu"] = r["ms_per_step"] - r["solve_ms"] + ilu["solve_ms"]   # same assembly, the other solve
            ilu["newton_step_dof_per_s_with_pl_ilu"] = case.ndof_global / (ilu["newton_step_ms_with_pl_ilu"] * 1e-3)
        except Exception as e:
            ilu = {"failed": f"{type(e).__name__}: {e}"}
    peak, peak_kind = hbm_peak()
    out = None
    if rank == 0:
        n_el = case.tables.controlpoint_IDs.shape[1]
        spmv_traffic, spmv_src = ncu_traffic("prof_spmv")
        elem_traffic, elem_src = ncu_traffic("prof_elem")
        plane = "single GPU"
        if world > 1:
            p2p = os.environ.get("MFB_P2P", "1") != "0"
            halo = p2p and os.environ.get("MFB_P2P_HALO", "1") != "0"
            plane = (f"element blocks (x slabs) over {world} GPUs, one process per GPU; Krylov scalar batches: "
                     f"{'peer-memory mailboxes (CUDA IPC over NVLink) inside the reducing kernels' if p2p else 'ncclAllReduce'}; "
                     f"interface exchange-add: {'peer-memory push/signal/merge kernels' if halo else 'grouped ncclSend/ncclRecv + merge kernel'}"
                     + ("; NCCL for the handshake only" if p2p and halo else ""))
        a_blk = r["spmv_gbs_block_format"]
        out = {
            "metric": "newton_step_dof_per_s", "value": r["value"], "unit": "DOF/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(name, n, args.numbering, extra={
                "dof": case.ndof_global, "dof_rank0": case.ndof, "nnz_rank0": case.nnz, "elements": case.n_el_global,
                "parallelism": plane, "krylov_iterations_per_step": r["krylov_iterations"], "spmv_per_step": r["spmv"],
                "passes_max": r["passes_max"], "solver_converged": r["all_converged"], "initial_residual": r["initial_residual"],
                "final_residual_max": r["final_residual_max"], "setup_s": case.setup_s,
                "spmv_kernel": "k_spmv_mr (multi-row streams)" if os.environ.get("MFB_SPMV", "mr")[0] not in "r0" else "k_spmv_bsr (one warp per row)",
                "dist_check": dc}),
            "newton_step_ms": r["ms_per_step"], "assembly_ms": r["assembly_ms"], "assembly_dof_per_s": r["assembly_dof_per_s"],
            "element_kernel_ms": r["element_kernel_ms"], "K_linear_ms": r["K_linear_ms"], "solve_ms": r["solve_ms"], "spmv_ms": r["spmv_ms"],
            "spmv_share_of_step": r["spmv_share_of_step"], "ms_per_spmv_equivalent": r["ms_per_spmv_equivalent"],
            "non_spmv_ms_per_spmv": r["non_spmv_ms_per_spmv"], "launches_per_spmv": r["launches_per_spmv"],
            "spmv_gbs_csr_model": r["spmv_gbs_csr_model"],
            "halo_exchange_ms_per_step": breakdown["halo_exchange_ms_per_step"] if breakdown else 0.0,
            "krylov_reductions_ms_per_step": breakdown["krylov_reductions_ms_per_step"] if breakdown else None,
            "breakdown_pass": breakdown,
            # SpMV roofline on the bytes the block-CSR kernel has to move ((8 nv^2 + 4) B per node pair + vectors); the scalar-CSR
            # model of SURVEY §8(d) (12 B/nnz: what the reference's cuSPARSE operand moves) is kept beside it
            "roofline": {"bound": "hbm", "kernel": "k_spmv_mr<3>" if os.environ.get("MFB_SPMV", "mr")[0] not in "r0" else "k_spmv_bsr<3>",
                         "achieved": a_blk, "peak": peak, "unit": "GB/s", "frac": a_blk / peak,
                         "byte_model": "block format actually stored: (8 nv^2 + 4) U + 4 (N + 1) + 16 n",
                         "algorithmic_bytes": r["spmv_bytes_block_format"],
                         "achieved_csr_model": r["spmv_gbs_csr_model"], "frac_csr_model": r["spmv_gbs_csr_model"] / peak,
                         "algorithmic_bytes_csr_model": r["spmv_bytes_csr_model"],
                         "traffic": spmv_traffic if world == 1 else None, "traffic_source": spmv_src,
                         "peak_kind": peak_kind, "launches_timed": r["spmv_launches_timed"], "avg_ms": r["spmv_ms"]},
            "e2e": r.get("e2e"), "gpu_launches": r["gpu_launches"], "clocks": r.get("clocks"),
        }
        if name == "neo_hookean":
            # element kernel: FP64-pipe bound for hex20 x 27. "achieved" counts the flops the sum-factorised kernel
            # EXECUTES in its tangent/residual contraction (2*NQ*NA*NV*(NSD*NV*KS + NV*NA*KS + 4) = 0.68 Mflop/element); the
            # reference's term-by-term form would need 2.62 Mflop/element for the same matrix ("reference_equivalent_tflops").
            elem_ms = r["element_kernel_ms"]
            flops_exec = 2.0 * 27 * 20 * 3 * (3 * 3 * 3 + 3 * 20 * 3 + 4) * n_el
            fp64_peak, fp64_clk = fp64_peak_with_clocks(env, case)
            asm_bytes = 8.0 * case.nnz + 8.0 * case.ndof + 8.0 * case.ndof + 24.0 * case.tables.variable_size + 4.0 * 20 * n_el + 4.0 * 400 * n_el
            out["roofline_assembly"] = {
                "bound": "fp64", "kernel": "mfb_b0_nl (fused element kernel)", "avg_ms": elem_ms,
                "achieved": flops_exec / (elem_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": flops_exec / (elem_ms * 1e-3) / 1e12 / max(fp64_peak, 1e-9),
                "peak_kind": "measured live (mfb_measure_fp64_peak: register-resident DFMA chains, best of ~1 s of launches)",
                "peak_probe_clocks": fp64_clk,
                "flops_executed": flops_exec, "reference_equivalent_tflops": 2.62e6 * n_el / (elem_ms * 1e-3) / 1e12,
                "hbm_algorithmic_bytes": asm_bytes, "hbm_frac": asm_bytes / (elem_ms * 1e-3) / 1e9 / peak,
                "traffic": elem_traffic if world == 1 else None, "traffic_source": elem_src}
    case.close()
    del case
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (N = 1), the partitioned J2 step (N > 1) ---------------------------------------------
    extras = {}
    if not args.no_extras and not args.ncu and name == "neo_hookean":
        todo = (["thermal_conduction", "linear_elasticity", "thermo_elasticity", "j2_fused", "j2"] if world == 1 else ["j2_fused"])
        for wn in todo:
            try:
                c2 = Case(env, wn, WORKLOADS[wn]["box"], args.numbering)
                r2 = measure(c2, 1 if wn.startswith("j2") else 2, 1, e2e=False)
                r2["config"] = workload_config(wn, WORKLOADS[wn]["box"], args.numbering)
                r2["n_gpus"] = world
                r2["spmv_frac_of_hbm_peak_block_format"] = r2["spmv_gbs_block_format"] / peak
                r2["spmv_frac_of_hbm_peak_csr_model"] = r2["spmv_gbs_csr_model"] / peak
                extras[wn] = r2
                c2.close()
                del c2
            except Exception as e:      # an extra never costs the main line
                extras[wn] = {"failed": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if ilu is not None:
        extras["neo_hookean_pl_ilu"] = ilu
    out["extra"] = extras
    if not args.no_cpu_baseline and world == 1 and name == "neo_hookean":      # the CPU port is timed beside the single-GPU run only
        try:
            cb = CpuReference(args.cpu_n).step()[0]
            # the CUDA path on the SAME sample (same box, same state, same solver): the only same-workload GPU/CPU pair
            cs = Case(env, "neo_hookean", args.cpu_n, args.numbering)
            rs = measure(cs, 5, 3, e2e=True, clocks=False)
            cb["gpu_same_sample"] = {"value": rs["value"], "unit": "DOF/s", "e2e_value": rs["e2e"]["value"], "ms_per_step": rs["ms_per_step"],
                                     "krylov_iterations": rs["krylov_iterations"], "assembly_dof_per_s": rs["assembly_dof_per_s"],
                                     "ms_per_spmv_equivalent": rs["ms_per_spmv_equivalent"], "spmv_gbs_csr_model": rs["spmv_gbs_csr_model"]}
            cs.close()
            out["cpu_baseline"] = cb
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
            out["cpu_baseline"] = {"value": None, "unit": "DOF/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"failed: {type(e).__name__}: {e}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
