"""Multi-rank worker (launched by torchrun from tests/test_multi_gpu.py and usable by hand):
partitions a small Neo-Hookean box over WORLD_SIZE GPUs, runs one full update_OneStep! on the distributed
library and checks it against the same step on a single GPU (rank 0) and against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import metafem_b200 as m
    from metafem_jl_b200.frontend import mesh as fmesh, partition as pt, weakform as wf
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "6,4,3").split(","))
    size = (1.5, 1.0, 1.0)
    tables = fmesh.box_tables(size, n, "CUBE", groups=("left", "right"), numbering="scattered")
    spec = wf.neo_hookean(fixed_bg=1, traction_bg=2)
    N = tables.variable_size
    rng = np.random.default_rng(0)
    state = {b: 0.02 * np.sin(1.3 * tables.x[(i + 1) % 3] + 0.2 * i) * tables.x[0] + rng.uniform(-1e-4, 1e-4, N)
             for i, b in enumerate(("d1", "d2", "d3"))}
    glob = dict(mu=1.0, lam=10.0, tau_b=1e4)

    def setup(fd, scatter):
        for b, v in state.items():
            fd.controlpoints[b][:] = scatter(v)
        fd.controlpoints["Pl1"][:] = 0.05
        fd.global_vars.update(glob)
        fd.globalfield.converge_tol = 1e-9
        fd.linear_solver = lambda d: m.iterative_Solve(d, Sv_func="bicgstabl_GS", maxiter=3000, max_pass=10, s=4)

    part = pt.split_elements(tables, world)
    sub = pt.make_subdomains(tables, part, ranks=[rank])[rank]
    fd = m.FEM_Domain(sub.tables, spec, device=local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(m.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    m.init_distributed(fd, sub, rank, world, bytes(idt.cpu().numpy().tobytes()))
    setup(fd, lambda v: pt.scatter_field(sub, v))
    m.assemble_Global_Variables(fd)
    m.compile_Updater_GPU(1, fd)
    hist = m.update_OneStep(fd.time_discretization, max_iter=7, fem_domain=fd)
    m.dessemble_X(fd)
    # gather the owned part of d1..d3 on rank 0
    mine = np.zeros((3, N))
    own = sub.owned.astype(bool)
    for i, b in enumerate(("d1", "d2", "d3")):
        mine[i, sub.node_l2g[own] - 1] = fd.controlpoints[b][own]
    t = torch.from_numpy(mine).cuda()
    dist.all_reduce(t)
    xd = t.cpu().numpy()
    # interface copies must be bit-identical across ranks
    chk = np.zeros((3, N))
    cnt = np.zeros(N)
    for i, b in enumerate(("d1", "d2", "d3")):
        chk[i, sub.node_l2g - 1] = fd.controlpoints[b]
    cnt[sub.node_l2g - 1] = 1
    tc, tn = torch.from_numpy(chk).cuda(), torch.from_numpy(cnt).cuda()
    dist.all_reduce(tc); dist.all_reduce(tn)
    assert np.array_equal(tc.cpu().numpy() / tn.cpu().numpy() * tn.cpu().numpy(), tc.cpu().numpy())
    consistent = np.abs(tc.cpu().numpy() - xd * tn.cpu().numpy()).max()
    fd.close()
    ok = True
    if rank == 0:
        fs = m.FEM_Domain(tables, spec, device=local)
        setup(fs, lambda v: v)
        m.assemble_Global_Variables(fs)
        m.compile_Updater_GPU(1, fs)
        hs = m.update_OneStep(fs.time_discretization, max_iter=7, fem_domain=fs)
        m.dessemble_X(fs)
        xs = np.stack([fs.controlpoints[b] for b in ("d1", "d2", "d3")])
        fs.close()
        err = np.linalg.norm(xd - xs) / np.linalg.norm(xs)
        print(f"world={world} newton history distributed={hist} single={hs} rel diff x = {err:.3e} interface mismatch = {consistent:.3e}")
        ok = (len(hist) == len(hs) and abs(hist[0] - hs[0]) <= 1e-12 * hs[0] and hist[-1] < 1e-9 and err < 1e-4
              and consistent <= 1e-12 * np.abs(xs).max())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit("distributed Newton step does not match the single-GPU step")
    if rank == 0:
        print("DIST_OK")


if __name__ == "__main__":
    main()
