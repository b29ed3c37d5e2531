"""Multi-GPU tests. GPU part: torchrun with 2 ranks over NCCL (skipped on a one-GPU box).
CPU part: the same partition / interface exchange-add logic with world_size 2 over gloo, checked with the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, env=None, nproc=2, port=29533, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


def _need_gpus(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs (run with gpurun --gpus {n})")


PLANES = {"peer_memory": {}, "nccl": {"MFB_P2P": "0"}, "nccl_halo": {"MFB_P2P_HALO": "0"}}


@pytest.mark.gpu
@pytest.mark.parametrize("plane", list(PLANES))
@pytest.mark.parametrize("case", ["neo_hookean", "thermo_elasticity", "j2", "j2_fused"])
def test_two_rank_newton_step_matches_single_gpu(built_lib, case, plane):
    """One full update_OneStep! on 2 ranks == the same step on one GPU, for 3- and 4-variable systems and for the J2 history
    arrays on a partitioned mesh, over every data plane (peer-memory mailboxes + halo flags, NCCL allreduce + send/recv)."""
    _need_gpus(2)
    out = _torchrun([case, "6,4,3"], env=PLANES[plane], port=29533 + list(PLANES).index(plane))
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    print(out.stdout[-600:])


@pytest.mark.gpu
@pytest.mark.parametrize("plane", ["peer_memory", "nccl"])
def test_two_rank_soak_with_random_stream_delays(built_lib, plane):
    _need_gpus(2)
    out = _torchrun(["neo_hookean", "6,4,3", "soak"], env=PLANES[plane], port=29543)
    assert out.returncode == 0 and "SOAK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    print(out.stdout[-300:])


def test_partition_covers_mesh_and_interfaces_are_consistent():
    import metafem_b200  # noqa: F401
    from metafem_jl_b200.frontend import mesh as fmesh, partition as pt
    t = fmesh.box_tables((1.0, 1.0, 1.0), (5, 3, 2), "CUBE", groups=("left", "right"))
    for P, method in ((2, "slab"), (3, "slab"), (4, "slab"), (4, "box"), (8, "box"), (5, "box")):
        part = pt.split_elements(t, P, method=method)
        assert np.array_equal(np.unique(part), np.arange(P))
        counts = np.bincount(part, minlength=P)
        assert counts.max() - counts.min() <= (1 if method == "slab" else P)               # balanced element counts
        subs = pt.make_subdomains(t, part)
        assert sum(s.tables.controlpoint_IDs.shape[1] for s in subs.values()) == t.controlpoint_IDs.shape[1]
        assert sum(int(s.owned.sum()) for s in subs.values()) == t.variable_size        # every node owned exactly once
        for r, s in subs.items():
            assert np.array_equal(s.tables.x, t.x[:, s.node_l2g - 1])
            assert np.array_equal(s.node_l2g[s.tables.controlpoint_IDs - 1], t.controlpoint_IDs[:, s.elem_l2g - 1])
            for q, l in s.shared.items():
                assert np.array_equal(s.node_l2g[l - 1], subs[q].node_l2g[subs[q].shared[r] - 1])
            nfac = sum(len(v) for v in s.tables.bg_fIDs.values())
            assert nfac == len(s.tables.facet_element_ID)
        for g in t.bg_fIDs:
            assert sum(len(s.tables.bg_fIDs[g]) for s in subs.values()) == len(t.bg_fIDs[g])


def _gloo_worker(rank, world, port, q, method="slab"):
    """Oracle on each rank's subdomain; interface exchange-add through gloo; compare with the undivided oracle."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import metafem_b200  # noqa: F401
    from metafem_jl_b200.frontend import mesh as fmesh, partition as pt
    from helpers import spec_for
    from oracle import assembly as oasm, solver as osv, femmesh as fm, discretization as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = fmesh.box_tables((1.5, 1.0, 1.0), (4, 2, 2), "CUBE", groups=("left", "right"))
    spec = spec_for("neo_hookean")
    sp = D.initialize_Classical_Element(3, "CUBE", 2, 1, 5, "Serendipity")

    def oracle_domain(tab):
        mesh = fm.WPMesh()
        mesh.space, mesh.x, mesh.controlpoint_IDs, mesh.variable_size = sp, tab.x, np.asarray(tab.controlpoint_IDs), tab.x.shape[1]
        mesh.facet_element_ID, mesh.facet_element_eindex = tab.facet_element_ID, tab.facet_element_eindex
        mesh.bg_fIDs = {g: np.asarray(v, dtype=np.int32) for g, v in tab.bg_fIDs.items()}
        fm.update_Mesh(mesh)
        dom = oasm.Domain(mesh, spec)
        dom.global_vars.update(mu=1.0, lam=10.0, tau_b=1e4)
        return dom

    def fill(dom, gx, pick):
        for i, b in enumerate(("d1", "d2", "d3")):
            dom.cp[b][:] = pick(0.02 * np.sin(1.3 * gx[(i + 1) % 3] + 0.2 * i) * gx[0])
        dom.cp["Pl1"][:] = 0.05
        oasm.assemble_Global_Variables(dom)
        osv.update_Time(dom); osv.initialize_dx(dom); oasm.K_linear_func(dom); osv.update_x_star(dom); oasm.K_nonlinear_func(dom)

    part = pt.split_elements(t, world, method=method)
    sub = pt.make_subdomains(t, part, ranks=[rank])[rank]
    dom = oracle_domain(sub.tables)
    fill(dom, t.x, lambda v: pt.scatter_field(sub, v))
    Nl, N = sub.tables.variable_size, t.variable_size

    def halo_add(vec):                                   # vec: (3, Nl) variable-major local vector
        out = vec.copy()
        for q in sub.neighbors:
            ids = sub.shared[q] - 1
            send = torch.from_numpy(np.ascontiguousarray(vec[:, ids]))
            recv = torch.empty_like(send)
            reqs = [dist.isend(send, q), dist.irecv(recv, q)]
            for r_ in reqs:
                r_.wait()
            out[:, ids] += recv.numpy()
        return out

    res = halo_add(dom.globalfield.residue.reshape(3, Nl))
    A = oasm.csr_from_globalfield(dom.globalfield)
    rng = np.random.default_rng(7)
    xg = rng.standard_normal((3, N))
    y = halo_add((A @ xg[:, sub.node_l2g - 1].ravel()).reshape(3, Nl))
    own = sub.owned.astype(bool)
    dot_local = float((xg[:, sub.node_l2g - 1][:, own] * y[:, own]).sum())
    td = torch.tensor([dot_local], dtype=torch.float64)
    dist.all_reduce(td)
    ok = True
    if rank == 0:
        gdom = oracle_domain(t)
        fill(gdom, t.x, lambda v: v)
        gres = gdom.globalfield.residue.reshape(3, N)
        gy = (oasm.csr_from_globalfield(gdom.globalfield) @ xg.ravel()).reshape(3, N)
        ok &= np.abs(res - gres[:, sub.node_l2g - 1]).max() <= 1e-12 * np.abs(gres).max()
        ok &= np.abs(y - gy[:, sub.node_l2g - 1]).max() <= 1e-12 * np.abs(gy).max()
        ok &= abs(td.item() - float((xg * gy).sum())) <= 1e-11 * abs(float((xg * gy).sum()))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_subdomain_assembly_equals_global():
    """N>1 host logic on CPU: partition + interface exchange-add + owner-masked dot reproduce the undivided result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_gloo_world4_box_partition_equals_global():
    """The same check on a 2 x 2 box partition (recursive coordinate bisection): nodes on the centre line are shared by four
    ranks, every rank has three neighbours."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 4, 29613, q, "box")) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res
