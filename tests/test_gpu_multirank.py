"""Two-GPU check of the halo sum (the Tpetra Export(overlapped -> owned, ADD) replacement, linearAlgebraInterface_matrix.hpp:233-237):
two ranks assemble their z-slabs on their own GPUs, exchange ghost rows -- through the p2p transport (peer stores over NVLink, the
default on one node) and through the NCCL transport -- and the owned rows must equal BOTH the single-GPU assembly of the whole mesh
and the CPU oracle's assembly of the whole mesh to 1e-12.  Skipped on boxes with one GPU (the driver's gpu tier); run with
`gpurun --gpus 2` (log of the last run: profiles/r02_multirank_2gpu.log)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = (12, 10, 8)


def _make(kind, n, rank, world):
    from mrhyde_b200.problems import SystemBrick, ThermalBrick
    if kind.startswith("thermal"):
        options = {"column elements": 16, "min segment levels": 2}
        if kind == "thermal_overlap" and world > 1:   # ghost-row chains first, exchange beside the rest of the assembly
            options["overlap halo"] = "true"
        if kind == "thermal_nccl":
            options["halo transport"] = "nccl"
        if kind == "thermal_nopush":   # p2p transport with the copy phase in the halo kernel instead of the assembly kernel
            options["halo push"] = "false"
        return ThermalBrick(3, n, device=rank, rank=rank, nranks=world, options=options)
    if kind == "maxwell":      # edge / face lattices cut into z-slabs (problems.slab_partition); n[0] == n[1]
        from mrhyde_b200.problems import MaxwellBrick
        return MaxwellBrick(n[0], device=rank, rank=rank, nranks=world, nz=n[2])
    if kind == "leq2":         # hex-Q2 node lattice cut into z-slabs
        from mrhyde_b200.problems import ElasticityQ2Brick
        return ElasticityQ2Brick(n[0], device=rank, rank=rank, nranks=world, nz=n[2])
    return SystemBrick({"le": "linearelasticity", "ns": "navier stokes"}[kind], 3, n, device=rank, rank=rank, nranks=world, options={"batch elems": 300})


def _size(kind):
    return {"maxwell": (6, 6, 4), "leq2": (4, 4, 3)}.get(kind, N)


def _worker(rank, world, port, q, kind="thermal"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mrhyde_b200.problems import ThermalBrick
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        _worker_body(rank, world, q, kind, dev)
    except Exception as e:   # report instead of leaving the parent to wait for the queue time-out
        q.put(("error", rank, repr(e)))
        os._exit(1)
    finally:
        dist.destroy_process_group()


def _worker_body(rank, world, q, kind, dev):
    import torch
    import torch.distributed as dist
    prob = _make(kind, _size(kind), rank, world)
    uid = torch.from_numpy(prob.plan.comm_unique_id()).to(dev) if rank == 0 else torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    prob.plan.comm_init(uid.cpu().numpy(), rank, world)
    prob.plan.set_halo(prob.col_gids)
    u = torch.from_numpy(prob.state()).to(dev)
    res = torch.zeros(prob.n_rows, dtype=torch.float64, device=dev)
    jac = torch.zeros(prob.nnz, dtype=torch.float64, device=dev)
    assert prob.plan.stat("halo_p2p") == (0 if kind in ("thermal_nccl", "thermal_overlap") else 1)
    for _ in range(4 if kind in ("thermal_overlap", "thermal", "thermal_nopush") else 1):   # repeated: side stream / events, both slab parities of the p2p transport
        res.zero_()
        jac.zero_()
        prob.plan.assemble_jacres(u, res, jac)
        prob.plan.halo_sum(res, jac)
    torch.cuda.synchronize()
    if kind in ("thermal", "thermal_nopush"):   # in-kernel halo push: the assembly kernel of a rank with ghost rows stores them into the owner's slab
        assert prob.plan.stat("pushed_assembles") == (4 if kind == "thermal" and prob.n_owned < prob.n_rows else 0)
    if kind == "thermal_overlap" and prob.n_owned < prob.n_rows:   # only ranks that hold ghost rows start an exchange early
        assert prob.plan.stat("overlapped_assembles") == 4 and 0 < prob.plan.stat("n_early_chains") < prob.plan.stat("n_chains")
    no = prob.n_owned
    assert prob.plan.owned_extent() == (no, int(prob.rowptr[no]))   # the owned matrix is the prefix of the local CSR arrays (no compaction pass)
    q.put((rank, prob.col_gids.copy(), res[:no].cpu().numpy(), prob.rowptr[: no + 1].copy(), prob.colind[: prob.rowptr[no]].copy(),
           jac[: prob.rowptr[no]].cpu().numpy(), prob.state()[:no].copy()))
    dist.barrier()


def _oracle_global(oracle_lib, kind, n, ug):
    """The CPU oracle on the whole mesh with the same state (the builders number dofs like the oracle: tests/test_builders.py)."""
    import configs
    mesh = {"Mesh/NX": n[0], "Mesh/NY": n[1], "Mesh/NZ": n[2], "Mesh/perturb": 0.0}
    if kind.startswith("thermal"):
        cfg = configs.variant(configs.THERMAL_3D, **mesh)
    elif kind == "maxwell":
        cfg = configs.variant(configs.MAXWELL_3D, **dict(mesh, **{"Physics/Dirichlet conditions": {}, "Functions": {"current x": "sin(2*pi*z)"}}))
    elif kind == "leq2":
        cfg = configs.variant(configs.LE_3D, **dict(mesh, **{"Discretization/order": {"dx": 2, "dy": 2, "dz": 2}, "Discretization/quadrature": 4,
                                                             "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)",
                                                                           "source dy": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)", "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"}}))
    elif kind == "le":
        cfg = configs.variant(configs.LE_3D, **dict(mesh, Functions={"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)",
                                                                       "source dy": "sin(2*pi*x)*sin(2*pi*y)", "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"}))
    else:
        cfg = configs.variant(configs.NS_3D, **dict(mesh, Functions={"source ux": "1.0", "viscosity": "1.0", "density": "1.0"}))
    op = oracle_lib.OracleProblem(cfg)
    return op.assemble_jacres(ug)


@pytest.mark.parametrize("kind", ["thermal", "thermal_nopush", "thermal_nccl", "thermal_overlap", "le", "ns", "maxwell", "leq2"])
def test_two_gpu_halo_sum_equals_single_gpu(oracle_lib, product_lib, kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from mrhyde_b200.problems import ThermalBrick
    # MRHYDE_B200_TEST_WORLD=4 (or 8) on a larger box: interior ranks hold ghost rows AND receive them (push to one side, add from the other)
    world = max(2, min(torch.cuda.device_count(), int(os.environ.get("MRHYDE_B200_TEST_WORLD", "2"))))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + {"thermal": 0, "thermal_overlap": 7, "le": 3, "ns": 5, "thermal_nccl": 9, "maxwell": 11, "leq2": 13, "thermal_nopush": 15}[kind]
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    outs = []
    try:
        for _ in range(world):
            item = q.get(timeout=180)
            if item[0] == "error":
                pytest.fail("rank %d: %s" % (item[1], item[2]))
            outs.append(item)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:   # never leave a rank behind (a rank stuck in a flag wait would keep its GPU busy after the test)
        for p in procs:
            if p.is_alive():
                p.kill()
    outs.sort(key=lambda t: t[0])
    n = _size(kind)
    glob = _make(kind, (n[0], n[1], world * n[2]), 0, 1)
    dev = torch.device("cuda:0")
    # the same state on the global mesh: value of every global row from the rank that owns it
    ug = np.zeros(glob.n_rows)
    for rank, gids, res, rp, ci, jac, st in outs:
        ug[gids[: len(st)]] = st
    d_u = torch.from_numpy(ug).to(dev)
    d_res = torch.zeros(glob.n_rows, dtype=torch.float64, device=dev)
    d_jac = torch.zeros(glob.nnz, dtype=torch.float64, device=dev)
    glob.plan.assemble_jacres(d_u, d_res, d_jac)
    torch.cuda.synchronize()
    res_g, jac_g = d_res.cpu().numpy(), d_jac.cpu().numpy()
    scale_r, scale_j = np.abs(res_g).max(), np.abs(jac_g).max()
    # oracle leg: the single-GPU assembly of the whole mesh is itself the oracle's (same rows, same graph)
    res_o, jac_o = _oracle_global(oracle_lib, kind, (n[0], n[1], world * n[2]), ug)
    assert np.max(np.abs(res_g - res_o)) <= 1e-12 * scale_r and np.max(np.abs(jac_g - jac_o)) <= 1e-12 * scale_j
    for rank, gids, res, rp, ci, jac, st in outs:
        own = gids[: len(st)]
        assert np.max(np.abs(res - res_g[own])) <= 1e-12 * scale_r
        for i, g in enumerate(own):
            a, b = glob.rowptr[g], glob.rowptr[g + 1]
            cg = gids[ci[rp[i]:rp[i + 1]]]
            o = np.argsort(cg)
            assert np.array_equal(cg[o], glob.colind[a:b])
            assert np.max(np.abs(jac[rp[i]:rp[i + 1]][o] - jac_g[a:b])) <= 1e-12 * scale_j
