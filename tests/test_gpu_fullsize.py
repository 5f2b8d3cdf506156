"""Oracle comparison of every BASELINE config AT ITS FULL SIZE ("no BASELINE config is compared with the oracle
at its own size").  The whole mesh is too slow for the CPU oracle, so the oracle assembles a SUB-BRICK of it -- same coordinates,
same element order inside the brick, the global state restricted to its dofs -- and every row whose elements all lie inside the
sub-brick is compared entry by entry (1e-12 relative: residual entries against the vector max-norm, matrix values against the row
max-norm; the columns of such a row, mapped to global ids, must be exactly the global CSR row).  The sub-bricks cut through chain /
segment boundaries of the sweep plan and through the batches of the general path, so the comparison exercises the layouts the
benchmark times (class ring, bulk-copy flush, element batches), not a small-mesh special case.
Follows assemblyManager_jacres.hpp:336-606 (volume loop + scatter) and workset.cpp:600-834 (transient seeding)."""
import numpy as np
import pytest

import configs
import helpers

TOL = 1e-12
gpu = pytest.mark.gpu


def _sub_cfg(base, n_glob, origin, size):
    lo = [origin[d] / float(n_glob[d]) for d in range(3)]
    hi = [(origin[d] + size[d]) / float(n_glob[d]) for d in range(3)]
    return configs.variant(base, **{"Mesh/NX": size[0], "Mesh/NY": size[1], "Mesh/NZ": size[2], "Mesh/perturb": 0.0,
                                    "Mesh/xmin": lo[0], "Mesh/xmax": hi[0], "Mesh/ymin": lo[1], "Mesh/ymax": hi[1], "Mesh/zmin": lo[2], "Mesh/zmax": hi[2]})


def compare_subbrick(prob, op, n_glob, origin, size, res_g, jac_g, u_glob, assemble_oracle):
    """prob: global builder object (lids, rowptr, colind on the host), op: OracleProblem of the sub-brick, res_g / jac_g: device
    tensors of the global assembly.  Returns (rows compared, res error, J error)."""
    import torch
    nx, ny, nz = size
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    e_glob = ((origin[0] + i) + n_glob[0] * ((origin[1] + j) + n_glob[1] * (origin[2] + k))).ravel()
    assert len(e_glob) == op.num_elems
    # sub dof -> global dof through the element dof lists (same local ordering: tests/test_builders.py)
    s2g = np.full(op.num_dofs, -1, dtype=np.int64)
    gl = prob.lids[e_glob].astype(np.int64)
    s2g[op.lids.ravel()] = gl.ravel()
    assert (s2g >= 0).all() and np.array_equal(s2g[op.lids], gl)        # the map is a function
    # geometry of the sub-brick is the global one
    assert np.abs(op.nodes[op.conn] - prob.nodes[prob.conn[e_glob]]).max() < 1e-14
    u_sub = u_glob[s2g]
    res_o, jac_o = assemble_oracle(u_sub, s2g)
    # complete rows: every global element of the dof lies in the sub-brick
    cnt_sub = np.bincount(op.lids.ravel(), minlength=op.num_dofs)
    cnt_glob = np.bincount(prob.lids.ravel(), minlength=prob.n_rows)
    rows_s = np.nonzero(cnt_sub == cnt_glob[s2g])[0]
    assert len(rows_s) > 0
    assert np.array_equal(op.is_fixed[rows_s], prob.is_fixed[s2g[rows_s]])
    rows_g = s2g[rows_s]
    # entries of those rows, ordered by (global row, global column)
    len_s = (op.rowptr[rows_s + 1] - op.rowptr[rows_s]).astype(np.int64)
    assert np.array_equal(len_s, prob.rowptr[rows_g + 1] - prob.rowptr[rows_g])   # same row lengths: the pattern of a complete row is local
    starts = np.repeat(op.rowptr[rows_s], len_s)
    within = np.arange(len_s.sum()) - np.repeat(np.cumsum(len_s) - len_s, len_s)
    idx_s = starts + within
    rg = np.repeat(rows_g, len_s)
    cg = s2g[op.colind[idx_s]]
    order = np.lexsort((cg, rg))
    ro = np.argsort(rows_g, kind="stable")
    len_g = len_s[ro]
    idx_g = np.repeat(prob.rowptr[rows_g[ro]], len_g) + (np.arange(len_g.sum()) - np.repeat(np.cumsum(len_g) - len_g, len_g))
    assert np.array_equal(cg[order], prob.colind[idx_g])                # identical CSR pattern, bit-exact
    def take(a, idx):   # device tensor or host array
        return a[torch.from_numpy(idx).to(a.device)].cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)[idx]
    jv = take(jac_g, idx_g)
    jo = jac_o[idx_s][order]
    rowmax = np.maximum.reduceat(np.abs(jo), np.cumsum(len_g) - len_g)
    rowmax[rowmax == 0.0] = 1.0
    e_jac = float(np.max(np.abs(jv - jo) / np.repeat(rowmax, len_g)))
    rv = take(res_g, rows_g)
    scale = float(abs(res_g).max())
    e_res = float(np.max(np.abs(rv - res_o[rows_s]))) / (scale if scale > 0 else 1.0)
    return len(rows_s), e_res, e_jac


def _device_assemble(prob, u, time=None):
    import torch
    dev = torch.device("cuda:0")
    d_u = torch.from_numpy(u).to(dev)
    res = torch.full((prob.n_rows,), 3.0, dtype=torch.float64, device=dev)
    jac = torch.full((prob.nnz,), 3.0, dtype=torch.float64, device=dev)
    prob.plan.assemble_jacres(d_u, res, jac, time=time)
    torch.cuda.synchronize()
    return res, jac


@gpu
@pytest.mark.parametrize("ring", ["auto", "full"])
def test_thermal_128_subbricks_match_oracle(oracle_lib, product_lib, ring):
    """BASELINE configs[1]: 128^3 hex-Q1 thermal through the sweep kernel (class ring + bulk flush = what bench.py times; ring=full =
    the layout of general meshes) against the oracle on 128 x 128 x 4 slabs at the bottom, across a segment cut and at the top."""
    from mrhyde_b200.problems import ThermalBrick
    n = 128
    prob = ThermalBrick(3, [n, n, n], device=0, options={"accumulate": "false", "ring": ring})
    u = prob.state()
    res, jac = _device_assemble(prob, u)
    total = 0
    for k0 in (0, 62, n - 4):
        op = oracle_lib.OracleProblem(_sub_cfg(configs.THERMAL_3D, (n, n, n), (0, 0, k0), (n, n, 4)))
        nrows, e_res, e_jac = compare_subbrick(prob, op, (n, n, n), (0, 0, k0), (n, n, 4), res, jac, u, lambda us, m: op.assemble_jacres(us))
        assert e_res < TOL and e_jac < TOL, (ring, k0, e_res, e_jac)
        total += nrows
    assert total >= 3 * 3 * 129 * 129


@gpu
def test_thermal_128_transient_subbrick_matches_oracle(oracle_lib, product_lib):
    """The transient build of the class ring (J = alpha_u K + alpha_t M, BDF1 / DIRK-1,2 stage) at full size."""
    import torch
    from mrhyde_b200.problems import ThermalBrick
    n = 128
    fns = {"density": "2.0", "specific heat": "1.5"}
    prob = ThermalBrick(3, [n, n, n], device=0, functions=fns, options={"accumulate": "false"})
    rng = np.random.default_rng(11)
    u, up = prob.state(), rng.standard_normal(prob.n_rows)
    d_up = torch.from_numpy(up).to(torch.device("cuda:0"))
    A, b, c = [[0.5]], [1.0], [0.5]
    ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[d_up], sol_stage=[d_up])
    res, jac = _device_assemble(prob, u, time=ts)
    k0 = 30
    cfg = _sub_cfg(configs.variant(configs.THERMAL_3D, **{"Functions/density": "2.0", "Functions/specific heat": "1.5"}), (n, n, n), (0, 0, k0), (n, n, 3))
    op = oracle_lib.OracleProblem(cfg)
    op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
    nrows, e_res, e_jac = compare_subbrick(prob, op, (n, n, n), (0, 0, k0), (n, n, 3), res, jac, u,
                                           lambda us, m: op.assemble_jacres(us, sol_prev=[up[m]], sol_stage=[us]))
    op.set_time(False)
    assert e_res < TOL and e_jac < TOL, (e_res, e_jac)


@gpu
def test_navier_stokes_96_subbrick_matches_oracle(oracle_lib, product_lib):
    """BASELINE configs[3] per-GPU size: 96^3 hex-Q1 Navier-Stokes (SUPG + PSPG) against the oracle on 96 x 96 x 2 slabs."""
    from mrhyde_b200.problems import SystemBrick
    n = 96
    prob = SystemBrick("navier stokes", 3, [n, n, n], device=0, options={"accumulate": "false"})
    u = prob.state()
    res, jac = _device_assemble(prob, u)
    base = configs.variant(configs.NS_3D, **{"Functions": {"source ux": "1.0", "viscosity": "1.0", "density": "1.0"}})
    for k0 in (0, 47):
        op = oracle_lib.OracleProblem(_sub_cfg(base, (n, n, n), (0, 0, k0), (n, n, 2)))
        nrows, e_res, e_jac = compare_subbrick(prob, op, (n, n, n), (0, 0, k0), (n, n, 2), res, jac, u, lambda us, m: op.assemble_jacres(us))
        assert nrows >= 4 * 97 * 97 and e_res < TOL and e_jac < TOL, (k0, e_res, e_jac)


@gpu
def test_elasticity_q2_64_subbrick_matches_oracle(oracle_lib, product_lib):
    """BASELINE configs[2]: 64^3 hex-Q2 linear elasticity (6.44 M dofs, 1.2 G non-zeros) against the oracle on 12 x 12 x 2 bricks in a corner
    and in the interior (the oracle's 81-wide AD makes larger bricks slow)."""
    from mrhyde_b200.problems import ElasticityQ2Brick
    n = 64
    prob = ElasticityQ2Brick(n, device=0, options={"accumulate": "false"})
    u = prob.state()
    res, jac = _device_assemble(prob, u)
    base = configs.variant(configs.LE_3D, **{"Discretization/order": {"dx": 2, "dy": 2, "dz": 2}, "Discretization/quadrature": 4,
                                              "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)",
                                                            "source dy": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)", "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"}})
    for origin in ((0, 0, 0), (29, 40, 31), (52, 52, 62)):
        op = oracle_lib.OracleProblem(_sub_cfg(base, (n, n, n), origin, (12, 12, 2)))
        nrows, e_res, e_jac = compare_subbrick(prob, op, (n, n, n), origin, (12, 12, 2), res, jac, u, lambda us, m: op.assemble_jacres(us))
        assert nrows > 3 * 23 * 23 and e_res < TOL and e_jac < TOL, (origin, e_res, e_jac)


@gpu
def test_maxwell_64_subbrick_matches_oracle(oracle_lib, product_lib):
    """BASELINE configs[4]: 64^3 hex Maxwell (HCURL E + HDIV B), one DIRK-1,2 stage, against the oracle on 64 x 64 x 2 slabs."""
    import torch
    from mrhyde_b200.problems import MaxwellBrick
    n = 64
    prob = MaxwellBrick(n, device=0, options={"accumulate": "false"})
    rng = np.random.default_rng(5)
    u, up = prob.state(), 0.1 * rng.standard_normal(prob.n_rows)
    d_up = torch.from_numpy(up).to(torch.device("cuda:0"))
    A, b, c = [[0.5]], [1.0], [0.5]
    ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[d_up], sol_stage=[d_up])
    res, jac = _device_assemble(prob, u, time=ts)
    base = configs.variant(configs.MAXWELL_3D, **{"Physics/Dirichlet conditions": {}, "Functions": {"current x": "sin(2*pi*z)"}})
    for k0 in (0, 31):
        op = oracle_lib.OracleProblem(_sub_cfg(base, (n, n, n), (0, 0, k0), (n, n, 2)))
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
        nrows, e_res, e_jac = compare_subbrick(prob, op, (n, n, n), (0, 0, k0), (n, n, 2), res, jac, u,
                                               lambda us, m: op.assemble_jacres(us, sol_prev=[up[m]], sol_stage=[us]))
        op.set_time(False)
        assert nrows > 64 * 64 * 4 and e_res < TOL and e_jac < TOL, (k0, e_res, e_jac)


def test_subbrick_comparison_is_sound_on_the_host(oracle_lib, product_lib):
    """`compare_subbrick` itself, without a GPU: the oracle's assembly of a whole 7 x 6 x 5 thermal brick against the oracle on sub-bricks of
    it (complete rows agree to round-off; a perturbed global value is detected)."""
    from mrhyde_b200.problems import ThermalBrick
    n = (7, 6, 5)
    prob = ThermalBrick(3, n, device=-1)
    opg = oracle_lib.OracleProblem(configs.variant(configs.THERMAL_3D, **{"Mesh/NX": n[0], "Mesh/NY": n[1], "Mesh/NZ": n[2]}))
    u = prob.state()
    res_g, jac_g = opg.assemble_jacres(u)
    for origin, size in (((0, 0, 0), (7, 6, 2)), ((2, 1, 1), (4, 4, 3)), ((0, 0, 3), (7, 6, 2))):
        op = oracle_lib.OracleProblem(_sub_cfg(configs.THERMAL_3D, n, origin, size))
        nrows, e_res, e_jac = compare_subbrick(prob, op, n, origin, size, res_g, jac_g, u, lambda us, m: op.assemble_jacres(us))
        assert nrows > 0 and e_res < TOL and e_jac < TOL
    bad = jac_g.copy()
    row = 3 + 8 * (3 + 7 * 2)
    bad[prob.rowptr[row] + 5] *= 1.0 + 1e-9
    op = oracle_lib.OracleProblem(_sub_cfg(configs.THERMAL_3D, n, (2, 1, 1), (4, 4, 3)))
    assert compare_subbrick(prob, op, n, (2, 1, 1), (4, 4, 3), res_g, bad, u, lambda us, m: op.assemble_jacres(us))[2] > 10 * TOL
