"""world_size-2 checks of the multi-GPU host logic on CPU (gloo): element partition into z-slabs, owned-then-ghost
row ordering, global ids, the fixed-dof mask on the global boundary only, and the halo sum semantics (ghost rows added
into the owner's rows = Tpetra Export(overlapped -> owned, ADD), linearAlgebraInterface_matrix.hpp:233-237) -- the
summed result must equal the single-rank assembly of the same global mesh.  The per-rank plans are host-only
(device = -1); staged element matrices are a function of the GLOBAL element id so every rank agrees on them."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = (6, 5, 4)   # elements per rank slab


def _stage(gelem, stage_len):
    k = np.arange(stage_len)[None, :]
    return np.sin(0.37 * gelem[:, None] + 1.3 * k) + 0.01 * k


def _local_assembly(kind, rank, world):
    """Rank-local overlapped residual / Jacobian values on a host-only plan.  thermal: the sweep plan's scatter programs on
    staged element vectors that are a function of the global element id; le: the general path's kernel stages replayed on
    the host (debug aid, see test_general_emulation.py) with a state that is a function of the global dof id."""
    from mrhyde_b200.problems import SystemBrick, ThermalBrick
    n = N if world > 1 else (N[0], N[1], 2 * N[2])
    if kind == "thermal":
        prob = ThermalBrick(3, n, device=-1, rank=rank, nranks=world, options={"column elements": 4, "min segment levels": 2})
        ne = prob.n_elem
        stage = _stage(np.arange(ne) + rank * ne, prob.plan.stat("stage_len"))   # 44 doubles, or 16 with the class ring
        res, jac = np.zeros(prob.n_rows), np.zeros(prob.nnz)
        prob.plan.debug_scatter_host(stage, 1, res, jac)
    else:
        prob = SystemBrick("linearelasticity", 3, n, device=-1, rank=rank, nranks=world, options={"batch elems": 40})
        u = np.sin(0.01 * prob.row_gids) + 0.1
        res, jac = np.zeros(prob.n_rows), np.zeros(prob.nnz)
        prob.plan.debug_emulate(u, res, jac)
    return prob, res, jac


def _worker(rank, world, port, q, kind="thermal"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mrhyde_b200.problems import SystemBrick, ThermalBrick
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res_jac = _local_assembly(kind, rank, world)
        prob, res, jac = res_jac
        # ---- halo sum over gloo: ghost rows (local ids >= n_owned) go to the rank that owns their gid
        gids = prob.col_gids   # rows first, then column-only ghosts
        ghost = np.arange(prob.n_owned, prob.n_rows)
        payload = []
        for r in ghost:
            cols = prob.colind[prob.rowptr[r]:prob.rowptr[r + 1]]
            payload.append((int(gids[r]), res[r], gids[cols].tolist(), jac[prob.rowptr[r]:prob.rowptr[r + 1]].tolist()))
        gathered = [None] * world
        dist.all_gather_object(gathered, (gids[: prob.n_owned].tolist(), payload))
        mine = {int(g): i for i, g in enumerate(gids[: prob.n_owned])}
        for src in range(world):          # ascending source rank = the fixed order of HaloExchange::sum
            if src == rank:
                continue
            for gid, rv, cg, jv in gathered[src][1]:
                if gid not in mine:
                    continue
                r = mine[gid]
                res[r] += rv
                lc = {int(gids[c]): p for p, c in zip(range(prob.rowptr[r], prob.rowptr[r + 1]), prob.colind[prob.rowptr[r]:prob.rowptr[r + 1]])}
                for g, v in zip(cg, jv):
                    jac[lc[g]] += v
        owned = slice(0, prob.n_owned)
        q.put((rank, gids[owned].copy(), res[owned].copy(), prob.rowptr[: prob.n_owned + 1].copy(), gids[prob.colind[: prob.rowptr[prob.n_owned]]].copy(),
               jac[: prob.rowptr[prob.n_owned]].copy(), prob.is_fixed[owned].copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["thermal", "le"])
def test_two_rank_partition_and_halo_sum_equal_single_rank(product_lib, kind):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if kind == "le" else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference: the same global mesh (N[0] x N[1] x 2 N[2]) as ONE slab
    glob, res, jac = _local_assembly(kind, 0, 1)
    seen = np.zeros(glob.n_rows, dtype=bool)
    for rank, gids, r_res, rp, cg, r_jac, fixed in outs:
        assert not seen[gids].any(), "a global row is owned by two ranks"
        seen[gids] = True
        assert np.array_equal(fixed, glob.is_fixed[gids]), "Dirichlet mask must follow the GLOBAL boundary"
        assert np.allclose(r_res, res[gids], rtol=0, atol=1e-13)
        for i, g in enumerate(gids):
            a, b = glob.rowptr[g], glob.rowptr[g + 1]
            o = np.argsort(cg[rp[i]:rp[i + 1]])   # local column order is by LOCAL id (column-only ghosts last)
            assert np.array_equal(cg[rp[i]:rp[i + 1]][o], glob.colind[a:b]), "owned row pattern differs from the global graph"
            assert np.allclose(r_jac[rp[i]:rp[i + 1]][o], jac[a:b], rtol=0, atol=1e-13)
    assert seen.all(), "some global row is owned by no rank"
