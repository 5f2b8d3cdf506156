"""The synthetic-brick builders of mrhyde_b200/problems.py (numpy; they play the reference host's role for bench.py and the
full-size GPU tests) against the oracle's mesh / DOF-map / graph / tabulation arrays, and end to end: a host-only plan built
from a builder, replayed on the host, must reproduce the oracle's assembly of the same deck."""
import numpy as np
import pytest

import configs
import helpers

TOL = 1e-12


def _le_q2_cfg(n):
    return configs.variant(configs.LE_3D, **{"Discretization/order": {"dx": 2, "dy": 2, "dz": 2}, "Discretization/quadrature": 4, "Mesh/NX": n, "Mesh/NY": n,
                                              "Mesh/NZ": n, "Mesh/perturb": 0.0, "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)",
                                                                                                "source dy": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)",
                                                                                                "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"}})


def _same_arrays(pb, op):
    assert np.array_equal(pb.lids, op.lids) and np.array_equal(pb.rowptr, op.rowptr) and np.array_equal(pb.colind, op.colind)
    assert np.array_equal(pb.is_fixed, op.is_fixed)


def _same_assembly(pb, op, time=None, oracle_kw=None):
    u = pb.state()
    res_ref, jac_ref = op.assemble_jacres(u, **(oracle_kw or {}))
    res, jac = np.zeros(pb.n_rows), np.zeros(pb.nnz)
    pb.plan.debug_emulate(u, res, jac, time=time)
    assert helpers.rel_err_vec(res, res_ref) < TOL and helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL


def test_system_brick_matches_oracle(oracle_lib, product_lib):
    from mrhyde_b200.problems import SystemBrick
    n = [4, 3, 5]
    le = configs.variant(configs.LE_3D, **{"Mesh/perturb": 0.0, "Mesh/NX": 4, "Mesh/NY": 3, "Mesh/NZ": 5, "Functions": {"lambda": "1.0", "mu": "1.0",
                         "source dx": "sin(pi*x)*sin(pi*y)", "source dy": "sin(2*pi*x)*sin(2*pi*y)", "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"}})
    ns = configs.variant(configs.NS_3D, **{"Mesh/perturb": 0.0, "Mesh/NX": 4, "Mesh/NY": 3, "Mesh/NZ": 5, "Functions": {"source ux": "1.0", "viscosity": "1.0", "density": "1.0"}})
    for phys, cfg in (("linearelasticity", le), ("navier stokes", ns)):
        op = oracle_lib.OracleProblem(cfg)
        pb = SystemBrick(phys, 3, n, device=-1)
        _same_arrays(pb, op)
        _same_assembly(pb, op)


@pytest.mark.parametrize("physics", ["linearelasticity", "thermal"])
def test_q2_brick_matches_oracle(oracle_lib, product_lib, physics):
    from mrhyde_b200.problems import ElasticityQ2Brick, q2_reference_3d
    n = 3
    cfg = _le_q2_cfg(n) if physics == "linearelasticity" else configs.variant(configs.THERMAL_3D, **{"Discretization/order/T": 2, "Discretization/quadrature": 4,
                                                                                                     "Mesh/NX": n, "Mesh/NY": n, "Mesh/NZ": n})
    op = oracle_lib.OracleProblem(cfg)
    pb = ElasticityQ2Brick(n, device=-1, options={"kernel": "general"}, physics=physics)
    _same_arrays(pb, op)
    pts, wts, val, grad = q2_reference_3d()
    rb = op.ref_basis(0)
    assert np.abs(pts - op.qpts).max() < 1e-15 and np.abs(wts - op.qwts).max() < 1e-15
    assert np.abs(val - rb["val"]).max() < 1e-15 and np.abs(grad - rb["grad"]).max() < 1e-15
    _same_assembly(pb, op)


def test_maxwell_brick_matches_oracle(oracle_lib, product_lib):
    from mrhyde_b200.problems import MaxwellBrick, hcurl_hdiv_reference_3d
    n = 4
    cfg = configs.variant(configs.MAXWELL_3D, **{"Mesh/NX": n, "Mesh/NY": n, "Mesh/NZ": n, "Mesh/perturb": 0.0, "Physics/Dirichlet conditions": {},
                                                  "Functions": {"current x": "sin(2*pi*z)"}})
    op = oracle_lib.OracleProblem(cfg)
    pb = MaxwellBrick(n, device=-1)
    _same_arrays(pb, op)
    assert pb.nnz == 252 * n ** 3 + 69 * n ** 2 + 3 * n           # SURVEY 8(d) closed form
    pts, wts, (ev, ec), (fv, fd) = hcurl_hdiv_reference_3d()
    r0, r1 = op.ref_basis(0), op.ref_basis(1)
    assert max(np.abs(ev - r0["val"]).max(), np.abs(ec - r0["curl"]).max(), np.abs(fv - r1["val"]).max(), np.abs(fd - r1["div"]).max()) == 0.0
    rng = np.random.default_rng(5)
    up = 0.1 * rng.standard_normal(op.num_dofs)
    us = 0.1 * rng.standard_normal(op.num_dofs)
    op.set_time(True, time=0.3, dt=0.01, stage=0, A=[[0.5]], b=[1.0], c=[0.5], bdf=(1.0, -1.0))
    ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=[[0.5]], b=[1.0], c=[0.5], bdf=(1.0, -1.0), sol_prev=[up], sol_stage=[us])
    _same_assembly(pb, op, time=ts, oracle_kw=dict(sol_prev=[up], sol_stage=[us]))
    op.set_time(False)


@pytest.mark.parametrize("kind", ["maxwell", "leq2"])
def test_slab_partitions_reassemble_the_global_graph(product_lib, kind):
    """Element-wise z-slab partitions of the edge / face lattices (Maxwell) and of the hex-Q2 node lattice: every global dof is owned
    by exactly one rank, owned rows carry the global CSR row (through col_gids, column-only ghosts included), ghost rows come last,
    and the strong-Dirichlet mask / state are functions of the global id (meshInterface_construct.hpp:143-160 element partition,
    discretizationInterface_dof.hpp:129-137 owned-then-ghost numbering)."""
    from mrhyde_b200.problems import ElasticityQ2Brick, MaxwellBrick
    n, nz, world = (4, 2, 3) if kind == "maxwell" else (3, 2, 3)
    make = (lambda **kw: MaxwellBrick(n, device=-1, **kw)) if kind == "maxwell" else (lambda **kw: ElasticityQ2Brick(n, device=-1, **kw))
    g = make(nz=nz * world)
    seen = np.zeros(g.n_rows, dtype=int)
    for r in range(world):
        p = make(rank=r, nranks=world, nz=nz)
        assert p.n_owned == p.n_rows if r == world - 1 else p.n_owned < p.n_rows
        seen[p.row_gids[: p.n_owned]] += 1
        assert np.array_equal(p.is_fixed, g.is_fixed[p.row_gids])
        for i in range(0, p.n_owned, 5):
            gi = p.row_gids[i]
            cols = np.sort(p.col_gids[p.colind[p.rowptr[i]:p.rowptr[i + 1]]])
            assert np.array_equal(cols, g.colind[g.rowptr[gi]:g.rowptr[gi + 1]])
        for i in range(p.n_owned, p.n_rows):      # ghost rows hold the couplings of this rank's elements only: a subset of the global row
            gi = p.row_gids[i]
            cols = p.col_gids[p.colind[p.rowptr[i]:p.rowptr[i + 1]]]
            assert np.isin(cols, g.colind[g.rowptr[gi]:g.rowptr[gi + 1]]).all() and (p.colind[p.rowptr[i]:p.rowptr[i + 1]] < p.n_rows).all()
        # the element dof lists address local rows, and map to the global lists of the same elements
        e0 = r * nz * n * n
        assert np.array_equal(p.row_gids[p.lids], g.lids[e0:e0 + p.n_elem].astype(np.int64))
    assert (seen == 1).all()
