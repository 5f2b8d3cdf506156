"""CPU checks of the general path's kernel logic: mrhyde_b200_plan_debug_emulate replays the stage functions of
general_kernel.cuh (the source the CUDA build compiles) and the pull on a host-only plan, and the result is compared
with the oracle to the north_star tolerance.  This is a debugging aid for machines without a GPU -- the assemble entry
points refuse host-only plans (asserted below); the parity tests proper are tests/test_gpu_general.py."""
import numpy as np
import pytest

import configs
import helpers

TOL = 1e-12


def _setup_time(op, tableau):
    if tableau is None:
        return None, {}
    A, b, c, stage = tableau
    rng = np.random.default_rng(5)
    up, us = 0.1 * rng.standard_normal(op.num_dofs), 0.1 * rng.standard_normal(op.num_dofs)
    op.set_time(True, time=0.3, dt=0.01, stage=stage, A=A, b=b, c=c, bdf=(1.0, -1.0))
    ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=stage, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[up], sol_stage=[us] * len(b))
    return ts, dict(sol_prev=[up], sol_stage=[us] * len(b))


@pytest.mark.parametrize("name,cfg,opts,tableau,zero", configs.general_cases(), ids=[c[0] for c in configs.general_cases()])
def test_kernel_stages_match_oracle(oracle_lib, product_lib, name, cfg, opts, tableau, zero):
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options=dict({"kernel": "general"}, **opts))
    assert plan.stat("general") == 1
    u = np.zeros(op.num_dofs) if zero else helpers.manufactured_state(op)
    ts, kw = _setup_time(op, tableau)
    res_ref, jac_ref = op.assemble_jacres(u, **kw)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, res, jac, time=ts)
    op.set_time(False)
    assert helpers.rel_err_vec(res, res_ref) < TOL
    assert helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL
    if "batched" in name:
        assert plan.stat("general_batches") > 1


@pytest.mark.parametrize("name", ["thermal3d-state-dirk", "le3d-state-mu", "ns2d-state-viscosity", "thermal2d-weak-state", "le3d", "le3d-q2", "le2d-weak-neumann",
                                  "ns3d-neumann", "ns2d-bwe", "thermal3d-q2", "thermal3d-advection"])
def test_tensor_core_build_stages_match_oracle(oracle_lib, product_lib, name):
    """option jacobian=tensor (single-basis HGRAD modules): field-direction derivatives (S4d) + the contraction the device runs on the
    FP64 tensor cores (S4m), including coefficients that read solution fields, against the oracle.  The default replay above runs
    the stages of the derivative-lane build (S4b)."""
    _, cfg, opts, tableau, zero = next(c for c in configs.general_cases() if c[0] == name)
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options=dict({"kernel": "general", "jacobian": "tensor"}, **opts))
    u = np.zeros(op.num_dofs) if zero else helpers.manufactured_state(op)
    ts, kw = _setup_time(op, tableau)
    res_ref, jac_ref = op.assemble_jacres(u, **kw)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, res, jac, time=ts)
    op.set_time(False)
    assert helpers.rel_err_vec(res, res_ref) < TOL and helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL


def test_state_dependent_thermal_coefficient_leaves_the_sweep_kernel(oracle_lib, product_lib):
    """thermal diffusion: 1.0+T*T on HGRAD-1: the sweep kernel's collapsed (linear) Jacobian does not apply, the plan takes the general path."""
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 4, "Mesh/NY": 3, "Mesh/NZ": 3, "Functions/thermal diffusion": "1.0+T*T"})
    op = oracle_lib.OracleProblem(cfg)
    assert helpers.plan_from_oracle(op, cfg, device=-1).stat("general") == 1
    assert helpers.plan_from_oracle(op, configs.variant(cfg, **{"Functions/thermal diffusion": "1.0+x*x"}), device=-1).stat("general") == 0
    # element reductions are evaluated over all points of an element: general path too
    assert helpers.plan_from_oracle(op, configs.variant(cfg, **{"Functions/thermal diffusion": "1.0+emax(x*y)"}), device=-1).stat("general") == 1
    # Solver: lump mass redirects the scatter's columns: general path's pull
    lump = configs.variant(cfg, **{"Functions/thermal diffusion": "1.0", "Solver/lump mass": True})
    assert helpers.plan_from_oracle(op, lump, device=-1).stat("general") == 1


def test_host_only_plan_cannot_assemble(oracle_lib, product_lib):
    from mrhyde_b200.capi import MrhydeB200Error
    op = oracle_lib.OracleProblem(configs.LE_2D)
    plan = helpers.plan_from_oracle(op, configs.LE_2D, device=-1)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    with pytest.raises(MrhydeB200Error):
        plan.assemble_jacres_host(np.zeros(op.num_dofs), res, jac)


def test_residual_only_and_overwrite_modes(oracle_lib, product_lib):
    cfg = configs.NS_2D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    u = helpers.manufactured_state(op)
    res_ref = op.assemble_res(u)                      # the ScalarT workset path
    res = np.zeros(op.num_dofs)
    plan.debug_emulate(u, res, None, compute_jacobian=False)
    assert helpers.rel_err_vec(res, res_ref) < TOL
    r1, j1 = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, r1, j1)
    plan.set_option("accumulate", "false")
    r2, j2 = np.full(op.num_dofs, 7.0), np.full(op.nnz, -3.0)
    plan.debug_emulate(u, r2, j2)
    assert np.array_equal(r1, r2) and np.array_equal(j1, j2)


def test_unsupported_configuration_is_an_error(oracle_lib, product_lib):
    from mrhyde_b200.capi import MrhydeB200Error
    cfg = configs.variant(configs.LE_2D, **{"Discretization/quadrature": 6})   # 4x4 Gauss points: no kernel built
    op = oracle_lib.OracleProblem(cfg)
    with pytest.raises(MrhydeB200Error) as e:
        helpers.plan_from_oracle(op, cfg, device=-1)
    assert "no kernel" in str(e.value)


MASS_CASES = [("le3d", configs.LE_3D, [1.0, 2.0, 0.5]), ("maxwell", configs.MAXWELL_3D, [1.5, 0.7]), ("ns2d", configs.NS_2D, [1.0, 0.0, 1.0]),
              ("thermal3d-q2", configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 3, "Mesh/NY": 3, "Mesh/NZ": 2, "Mesh/perturb": 0.02,
                                                                    "Discretization/order/T": 2, "Discretization/quadrature": 4}), [2.0])]


@pytest.mark.parametrize("name,cfg,wts", MASS_CASES, ids=[c[0] for c in MASS_CASES])
@pytest.mark.parametrize("lump", [False, True], ids=["jacobi", "lumped"])
def test_weighted_mass_stages_match_oracle(oracle_lib, product_lib, name, cfg, wts, lump):
    """getWeightedMass: weighted mass values and the Jacobi / lumped diagonal vector."""
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    Mref, dref = op.weighted_mass(wts, lump)
    M, d = np.zeros(op.nnz), np.zeros(op.num_dofs)
    plan.debug_emulate_mass(wts, M, d, lump=lump)
    assert helpers.rel_err_rows(M, Mref, op.rowptr) < TOL and helpers.rel_err_vec(d, dref) < TOL


def test_general_path_error_reporting(oracle_lib, product_lib):
    """Set-up errors of the general path surface as status codes + messages (reference style: set-up errors are loud)."""
    from mrhyde_b200.capi import AssemblyPlan, MrhydeB200Error
    op = oracle_lib.OracleProblem(configs.LE_2D)
    rb = op.ref_basis(0)
    bases = [dict(type="HGRAD", order=1, card=rb["card"], val=rb["val"], grad=rb["grad"])]

    def make(physics, names):
        plan = AssemblyPlan(physics, 2, names, [0, 0], bases, op.ndof_elem, op.offsets, op.qpts, op.qwts, device=-1)
        plan.set_mesh(op.elem_nodes, op.lids)
        plan.set_graph(op.rowptr, op.colind, op.is_fixed)
        return plan

    with pytest.raises(MrhydeB200Error, match="has no device kernel"):
        make("porous", ["dx", "dy"]).finalize()
    with pytest.raises(MrhydeB200Error, match="variable order"):
        make("linearelasticity", ["dy", "dx"]).finalize()
    good = make("linear elasticity", ["dx", "dy"])       # the importer's alias (physicsImporter.cpp:170)
    good.finalize()
    assert good.stat("general") == 1
    # the mass entry point needs the general path; a sweep-kernel plan says so
    opt = oracle_lib.OracleProblem(configs.variant(configs.THERMAL_2D, **{"Mesh/NX": 4, "Mesh/NY": 4}))
    sweep = helpers.plan_from_oracle(opt, configs.THERMAL_2D, device=-1)
    with pytest.raises(MrhydeB200Error, match="general path"):
        sweep.debug_emulate_mass([1.0], np.zeros(opt.nnz), np.zeros(opt.num_dofs))
    # a graph that lacks an element coupling is rejected at plan time, not at run time
    bad = make("linearelasticity", ["dx", "dy"])
    rp = op.rowptr.copy()
    keep = np.ones(op.nnz, dtype=bool)
    keep[rp[5]] = False                                   # drop the first entry of row 5
    rp[6:] -= 1
    bad.set_graph(rp, op.colind[keep], op.is_fixed)
    with pytest.raises(MrhydeB200Error, match="lacks an entry"):
        bad.finalize()


@pytest.mark.parametrize("name,cfg,wts", MASS_CASES, ids=[c[0] for c in MASS_CASES])
def test_mass_apply_stages_match_oracle(oracle_lib, product_lib, name, cfg, wts):
    """applyMassMatrixFree: y = M x without forming M, against the oracle and against the assembled mass matrix."""
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    x = np.random.default_rng(11).standard_normal(op.num_dofs)
    yref = op.apply_mass(wts, x)
    y = np.zeros(op.num_dofs)
    plan.debug_emulate_apply_mass(wts, x, y)
    assert helpers.rel_err_vec(y, yref) < TOL
    M, _ = op.weighted_mass(wts)
    assert helpers.rel_err_vec(op.csr(M) @ x, yref) < 1e-12


INITIAL_CASES = [
    ("le3d", configs.variant(configs.LE_3D, **{"Physics/Initial conditions": {"dx": "sin(pi*x)*y", "dy": "1.0+z"}}), 0.0),
    ("maxwell", configs.variant(configs.MAXWELL_3D, **{"Physics/Initial conditions": {"E[x]": "sin(pi*z)", "E[y]": "x*y", "B[z]": "cos(pi*x)+y"}}), 0.0),
    ("ns2d", configs.variant(configs.NS_2D, **{"Physics/Initial conditions": {"ux": "1.0-y*y", "pr": "x"}}), 0.0),
    ("thermal3d-q2", configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 3, "Mesh/NY": 3, "Mesh/NZ": 2, "Mesh/perturb": 0.02, "Discretization/order/T": 2,
                                                           "Discretization/quadrature": 4, "Physics/Initial conditions": {"T": "exp(-x)*y+t"}}), 0.3),
]


@pytest.mark.parametrize("name,cfg,time", INITIAL_CASES, ids=[c[0] for c in INITIAL_CASES])
def test_initial_projection_stages_match_oracle(oracle_lib, product_lib, name, cfg, time):
    """setInitial: right-hand side of the L2 projection of the `Initial conditions`, scalar and vector-valued
    bases; the mass matrix of the projection is getMass = the weighted mass with unit weights (tested above)."""
    op = oracle_lib.OracleProblem(cfg)
    op.set_time(False, time=time)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    ref = op.project_initial()
    op.set_time(False)
    assert np.abs(ref).max() > 0
    rhs = np.zeros(op.num_dofs)
    plan.debug_emulate_initial(rhs, time=time)
    assert helpers.rel_err_vec(rhs, ref) < TOL
    if name == "ns2d":   # solving M u = rhs with the consistent mass matrix reproduces a function of the discrete space: pr = x
        import scipy.sparse.linalg as spla
        M, _ = op.weighted_mass([1.0, 1.0, 1.0])
        u = spla.spsolve(op.csr(M).tocsc(), rhs)
        pr = op.offsets[1]
        for e in range(op.num_elems):
            assert np.allclose(u[op.lids[e][pr]], op.elem_nodes[e][:, 0], atol=1e-12)


def test_initial_projection_reproduces_reference_gold(oracle_lib, product_lib):
    """regression/maxwell/NonzeroIC through the kernel stages: projection right-hand side and unit-weight mass matrix from the
    host replay of the general path, solved on the host, give the L2 errors mrhyde.gold prints at time 0."""
    import scipy.sparse.linalg as spla
    deck, errs = helpers.gold_errors("maxwell/NonzeroIC")
    cfg = helpers.deck_to_cfg(deck)
    gold = {e["field"]: e["value"] for e in errs if e["time"] == 0.0}
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    rhs, M, d = np.zeros(op.num_dofs), np.zeros(op.nnz), np.zeros(op.num_dofs)
    plan.debug_emulate_initial(rhs)
    plan.debug_emulate_mass([1.0, 1.0], M, d)
    u = spla.spsolve(op.csr(M).tocsc(), rhs)
    for f in ("E", "B"):
        got = op.l2_error([f + "[x]", f + "[y]", f + "[z]"], u)
        assert abs(got - gold[f]) <= 0.5e-5 * gold[f], (f, got, gold[f])


def test_point_constraints_replace_rows(oracle_lib, product_lib):
    """disc->point_dofs: dofConstraints turns the whole Jacobian row into the identity row and leaves the residual alone
    (assemblyManager_constraints.hpp:97-116, 261-266)."""
    from mrhyde_b200.capi import MrhydeB200Error
    cfg = configs.NS_2D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    pts = np.array([5, 17, op.num_dofs - 3], dtype=np.int32)
    op.set_point_dofs(pts)
    plan.set_point_dofs(pts)
    u = helpers.manufactured_state(op)
    res_ref, jac_ref = op.assemble_jacres(u)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, res, jac)
    assert helpers.rel_err_vec(res, res_ref) < TOL and helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL
    for d in pts:
        row = jac[op.rowptr[d]:op.rowptr[d + 1]]
        assert row.sum() == 1.0 and np.count_nonzero(row) == 1 and row[np.where(op.colind[op.rowptr[d]:op.rowptr[d + 1]] == d)[0][0]] == 1.0
    with pytest.raises(MrhydeB200Error):
        plan.set_point_dofs(np.array([op.num_dofs], dtype=np.int32))
    plan.set_point_dofs(np.zeros(0, dtype=np.int32))   # clears the list
    op.set_point_dofs(np.zeros(0, dtype=np.int32))
    res_ref, jac_ref = op.assemble_jacres(u)
    res[:], jac[:] = 0.0, 0.0   # the plan accumulates
    plan.debug_emulate(u, res, jac)
    assert helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL


def test_thermoelastic_gold_through_kernel_stages(oracle_lib, product_lib):
    """regression/thermoelastic/2D_transient (block "thermal, linearelasticity"): ten BWE steps assembled by the host replay of the two-module
    kernel's stages reproduce the reference's printed L2 norms of T (0.331419 at t = 0.1 ... 0.498946 at 0.9)."""
    from test_oracle_golden import _thermoelastic_steps
    deck, errs = helpers.gold_errors("thermoelastic/2D_transient")
    cfg = helpers.deck_to_cfg(deck)
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1)
    assert plan.stat("general") == 1
    state = {}

    def set_time(t, dt):
        state["t"], state["dt"] = t, dt

    def assemble(us, u):
        ts = helpers.TimeSpec(time=state["t"], deltat=state["dt"], stage=0, A=((1.0,),), b=(1.0,), c=(1.0,), bdf=(1.0, -1.0), sol_prev=[u], sol_stage=[us])
        res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
        plan.debug_emulate(us, res, jac, time=ts)
        return res, jac

    _thermoelastic_steps(cfg, errs, op.num_dofs, assemble, op.l2_error, set_time, op.csr)


def test_fix_zero_rows_gives_unit_diagonal(oracle_lib, product_lib):
    """Solver: fix zero rows with a vanishing diffusion coefficient: every Jacobian row is empty, so every diagonal entry becomes 1
    (assemblyManager_jacres.hpp:609-626) -- in the oracle and in the product's replay."""
    name, cfg, opts, tableau, zero = [c for c in configs.general_cases() if c[0] == "thermal3d-fix-zero-rows"][0]
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    u = helpers.manufactured_state(op)
    _, jac_ref = op.assemble_jacres(u)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, res, jac)
    import scipy.sparse as sp
    for J in (jac_ref, jac):
        A = op.csr(J)
        assert abs(A - sp.identity(op.num_dofs)).max() == 0.0


@pytest.mark.parametrize("seed_what,seed_index", [(2, 0), (2, 1), (3, 0)], ids=["prev-step-0", "prev-step-1", "prev-stage-0"])
@pytest.mark.parametrize("name", ["thermal3d-advection", "ns2d-bwe", "maxwell-abc-bwe", "thermoelastic2d"])
def test_previous_step_and_stage_jacobians(oracle_lib, product_lib, name, seed_what, seed_index):
    """compute_previous_jac (seedwhat = 2, seedindex = stepindex: assemblyManager_jacres.hpp:176-190) and the previous-stage seeding
    (workset.cpp:727-785): the Jacobian with respect to sol_prev[index] / sol_stage[index], same residual, same evaluation point.
    Second stage of a two-stage DIRK tableau with BDF-2 weights so that every branch of the seeding has something to differentiate."""
    cfg = [c for c in configs.general_cases() if c[0] == name][0][1]
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1, options={"kernel": "general"})
    rng = np.random.default_rng(11)
    u = helpers.manufactured_state(op)
    prev = [0.1 * rng.standard_normal(op.num_dofs) for _ in range(2)]
    stg = [0.1 * rng.standard_normal(op.num_dofs) for _ in range(2)]
    A, b, c, bdf = ((0.25, 0.0), (0.5, 0.25)), (0.5, 0.5), (0.25, 0.75), (1.5, -2.0, 0.5)
    op.set_time(True, time=0.3, dt=0.01, stage=1, A=A, b=b, c=c, bdf=bdf)
    op.set_seeding(seed_what, seed_index)
    res_ref, jac_ref = op.assemble_jacres(u, sol_prev=prev, sol_stage=stg)
    op.set_seeding(1, 0)
    res_1, jac_1 = op.assemble_jacres(u, sol_prev=prev, sol_stage=stg)
    op.set_time(False)
    assert helpers.rel_err_vec(res_ref, res_1) < 1e-14 and not np.allclose(jac_ref, jac_1)   # same residual, another derivative
    ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=1, A=A, b=b, c=c, bdf=bdf, sol_prev=prev, sol_stage=stg, seed_what=seed_what, seed_index=seed_index)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.debug_emulate(u, res, jac, time=ts)
    assert helpers.rel_err_vec(res, res_ref) < TOL
    assert helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL


def test_set_initial_entry_point_and_oracle_composition(oracle_lib, product_lib):
    """setInitial as a whole: the oracle's composition (projection, unit-weight mass, fix_zero_rows) against the stage replays of the two
    kernels it is made of; the device entry point refuses host-only plans like every assemble call (no CPU path)."""
    from mrhyde_b200.capi import MrhydeB200Error
    cfg = configs.variant(configs.LE_3D, **{"Physics/Initial conditions": {"dx": "sin(pi*x)*y", "dy": "1.0+z"}})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, device=-1)
    rhs_ref, M_ref = op.set_initial()
    rhs, M, d = np.zeros(op.num_dofs), np.zeros(op.nnz), np.zeros(op.num_dofs)
    plan.debug_emulate_initial(rhs)
    plan.debug_emulate_mass(np.ones(3), M, d)
    assert helpers.rel_err_vec(rhs, rhs_ref) < TOL and helpers.rel_err_rows(M, M_ref, op.rowptr) < TOL   # a mass matrix has no empty row
    with pytest.raises(MrhydeB200Error):
        plan.set_initial(rhs, M)
