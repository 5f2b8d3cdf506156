"""Pins the CPU oracle (oracle/, the checker) to the reference's own regression answers: the input
decks and the L2-error lines of mrhyde.gold committed under tests/golden/ (generated from
/root/reference by tests/golden/make_golden.py).  The reference prints 6 significant digits."""
import json
import os

import numpy as np
import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))


def _tol(v):
    return 0.5e-5 * abs(v) + 1e-12   # half a unit of the 6th significant digit


def _errs(case):
    deck, errs = helpers.gold_errors(case)
    return helpers.deck_to_cfg(deck), errs


@pytest.mark.parametrize("case", ["thermal/2D_verification", "thermal/2D_verification_mpi", "thermal/3D_verification", "thermal/2D_mixed_bcs"])
def test_steady_thermal_gold(oracle_lib, case):
    cfg, errs = _errs(case)
    op = oracle_lib.OracleProblem(cfg)
    u = helpers.newton_steady(op, cfg, lambda v: op.assemble_jacres(v))
    checked = 0
    for e in errs:
        if e["norm"] != "L2 norm":
            continue   # L2-face norms need the face-term machinery, which is off the hot path
        labels = ["T"] if e["field"] == "T" else ["grad(T)[x]", "grad(T)[y]"]
        got = op.l2_error(labels, u)
        assert abs(got - e["value"]) <= _tol(e["value"]), (case, e, got)
        checked += 1
    assert checked >= 1


def test_transient_thermal_gold(oracle_lib):
    """regression/thermal/2D_verification_transient: BWE + BDF1, 20 steps (computeSolnTransientSeeded path)."""
    import scipy.sparse.linalg as spla
    cfg, errs = _errs("thermal/2D_verification_transient")
    op = oracle_lib.OracleProblem(cfg)
    nsteps = int(cfg["Solver"]["number of steps"])
    dt = float(cfg["Solver"]["final time"]) / nsteps
    u = np.zeros(op.num_dofs)
    gold = {round(e["time"], 6): e["value"] for e in errs}
    for n in range(nsteps):
        t = n * dt
        op.set_time(True, time=t, dt=dt, stage=0, A=((1.0,),), b=(1.0,), c=(1.0,), bdf=(1.0, -1.0))
        us = u.copy()
        for _ in range(2):
            res, jac = op.assemble_jacres(us, sol_prev=[u], sol_stage=[us])
            us = us + spla.spsolve(op.csr(jac).tocsc(), res)
        u = us
        got = op.l2_error(["T"], u)
        want = gold[round(t + dt, 6)]
        assert abs(got - want) <= _tol(want), (n, got, want)


def _thermoelastic_steps(cfg, errs, num_dofs, assemble, l2_error, set_time, csr):
    """regression/thermoelastic/2D_transient: block "thermal, linearelasticity" (T, dx, dy), BWE, 10 steps; the true solutions are 0, so the
    printed errors are the L2 norms of the fields: T is the pin, dx / dy stay at the linear solver's noise (gold 4.8e-8, here ~0)."""
    import scipy.sparse.linalg as spla
    nsteps = int(cfg["Solver"]["number of steps"])
    dt = float(cfg["Solver"]["final time"]) / nsteps
    u = np.zeros(num_dofs)
    gold = {(e["field"], round(e["time"], 6)): e["value"] for e in errs}
    for n in range(nsteps):
        t = n * dt
        set_time(t, dt)
        us = u.copy()
        for _ in range(2):
            res, jac = assemble(us, u)
            us = us + spla.spsolve(csr(jac).tocsc(), res)
        u = us
        want = gold[("T", round(t + dt, 6))]
        got = l2_error(["T"], u)
        assert abs(got - want) <= _tol(want), (n, got, want)
        assert l2_error(["dx"], u) < 1e-6 and l2_error(["dy"], u) < 1e-6
    return u


def test_thermoelastic_two_module_gold(oracle_lib):
    cfg, errs = _errs("thermoelastic/2D_transient")
    op = oracle_lib.OracleProblem(cfg)
    assert op.modules().replace(" ", "") == "thermal,linearelasticity" and op.var_names() == ["T", "dx", "dy"]
    _thermoelastic_steps(cfg, errs, op.num_dofs,
                         lambda us, u: op.assemble_jacres(us, sol_prev=[u], sol_stage=[us]), op.l2_error,
                         lambda t, dt: op.set_time(True, time=t, dt=dt, stage=0, A=((1.0,),), b=(1.0,), c=(1.0,), bdf=(1.0, -1.0)), op.csr)


def test_function_forest_matches_functions_valid_gold(oracle_lib):
    """regression/functions/Valid prints the decomposition forest; the oracle's Interpreter::split restatement
    must produce the same branches in the same order (pins the parse / evaluation order)."""
    fv = json.load(open(os.path.join(HERE, "golden", "functions_valid.json")))
    import configs
    cfg = configs.variant(configs.THERMAL_2D, **{"Mesh/NX": 2, "Mesh/NY": 2})
    # trees that only use fields the thermal block has (T); the others need HDIV/HCURL/HVOL extra variables
    skip = {"f14", "f15", "f16", "f17"}
    cfg["Functions"] = {k: str(v) for k, v in fv["functions"].items() if k not in skip}
    op = oracle_lib.OracleProblem(cfg)
    n = 0
    for name, branches in fv["forests"]["ip"].items():
        if name in skip or name not in cfg["Functions"]:
            continue
        got = [line.split(":", 1)[1].rsplit("|", 1)[0] for line in op.print_tree(name).strip().splitlines()]
        assert got == branches, (name, got, branches)
        n += 1
    assert n == 18


def test_q1_laplace_stencil_and_symmetry(oracle_lib):
    """Analytic check: the interior row of the hex-Q1 Laplace matrix on a uniform grid of size h is
    8h/3 (diagonal), 0 (face neighbours), -h/6 (edge neighbours), -h/12 (corner neighbours)."""
    import configs
    n = 4
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": n, "Mesh/NY": n, "Mesh/NZ": n})
    op = oracle_lib.OracleProblem(cfg)
    res, jac = op.assemble_jacres(np.zeros(op.num_dofs))
    A = op.csr(jac).toarray()
    h = 1.0 / n
    nn = n + 1
    c = 2 + 2 * nn + 2 * nn * nn
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                m = abs(di) + abs(dj) + abs(dk)
                want = {0: 8 * h / 3, 1: 0.0, 2: -h / 6, 3: -h / 12}[m]
                assert abs(A[c, c + di + dj * nn + dk * nn * nn] - want) < 1e-14
    free = ~op.is_fixed.astype(bool)
    Aff = A[np.ix_(free, free)]
    assert np.max(np.abs(Aff - Aff.T)) < 1e-14
    # constant state: K * 1 = 0, so the residual is only the source load (and zero with no source)
    cfg0 = configs.variant(cfg, **{"Functions/thermal source": "0.0"})
    op0 = oracle_lib.OracleProblem(cfg0)
    r0, _ = op0.assemble_jacres(np.ones(op0.num_dofs))
    assert np.max(np.abs(r0)) < 1e-14


def _newton(op, cfg, maxiter=10, tol=1e-6):
    """nonlinearSolver (solverManager_solvers.hpp:290-640): relative residual test against the first residual, then a
    full Newton update; defaults nonlinear TOL 1e-6, max nonlinear iters 10 (solverManager_construct.hpp:58-60)."""
    import scipy.sparse.linalg as spla
    solver = cfg.get("Solver", {})
    maxiter = int(solver.get("max nonlinear iters", maxiter))
    tol = float(solver.get("nonlinear TOL", tol))
    u = np.zeros(op.num_dofs)
    fixed = op.is_fixed.astype(bool)
    # the multi-field decks pinned here all carry homogeneous Dirichlet data (helpers.dirichlet_values is nodal, one field)
    for spec in cfg.get("Physics", {}).get("Dirichlet conditions", {}).values():
        assert all(float(v) == 0.0 for v in spec.values())
    u[fixed] = 0.0
    first = None
    for it in range(maxiter + 1):
        res, jac = op.assemble_jacres(u)
        nrm = np.max(np.abs(res))
        first = nrm if first is None else first
        if first == 0.0 or nrm / first < tol or it == maxiter:
            break
        u = u + spla.spsolve(op.csr(jac).tocsc(), res)
    return u


@pytest.mark.parametrize("case", ["le/3D_manufactured", "le/2D_manufactured"])
def test_linear_elasticity_gold(oracle_lib, case):
    """regression/le/{2D,3D}_manufactured: nested `Functions:` sources, lambda = mu = 1, hex/quad Q1, strong Dirichlet."""
    cfg, errs = _errs(case)
    cfg["Physics"]["Dirichlet conditions"].pop("scalar data", None)
    op = oracle_lib.OracleProblem(cfg)
    u = _newton(op, cfg)
    assert len(errs) == op.dim
    for e in errs:
        got = op.l2_error([e["field"]], u)
        assert abs(got - e["value"]) <= _tol(e["value"]), (case, e, got)


def test_navier_stokes_channel_gold(oracle_lib):
    """regression/navierstokes/channel: 2-D PSPG-stabilised equal-order Q1, natural in/outflow, Newton to 1e-6."""
    cfg, errs = _errs("navierstokes/channel")
    cfg["Physics"]["Dirichlet conditions"].pop("scalar data", None)
    op = oracle_lib.OracleProblem(cfg)
    u = _newton(op, cfg)
    assert {e["field"] for e in errs} == {"ux", "pr", "uy"}
    for e in errs:
        got = op.l2_error([e["field"]], u)
        assert abs(got - e["value"]) <= _tol(e["value"]), (e, got)


def test_maxwell_planewave_gold(oracle_lib):
    """regression/maxwell/PlaneWave: 2x2x100 hex, x/y periodic, lowest-order HCURL E + HDIV B, DIRK-1,2 with BDF1,
    ten steps of 1e-15 s with a Gaussian-modulated current sheet.  Pins the HCURL/HDIV tabulation and push-forward
    (Intrepid2 is not visible under /root/reference), the Maxwell module's 3-D form, the stage
    combination of computeSolnTransientSeeded and the comparison operators of the function manager."""
    import scipy.sparse.linalg as spla
    cfg, errs = _errs("maxwell/PlaneWave")
    cfg["Physics"].pop("Initial conditions", None)
    op = oracle_lib.OracleProblem(cfg)
    assert op.num_elems == 400 and op.ndof_elem == 18 and op.num_dofs == 1208 + 1204   # periodic edge / face counts
    nsteps = int(cfg["Solver"]["number of steps"])
    dt = float(cfg["Solver"]["final time"]) / nsteps
    u = np.zeros(op.num_dofs)
    gold = {(e["field"], n): e["value"] for e in errs for n in range(nsteps + 1) if abs(e["time"] - n * dt) < 1e-3 * dt}
    assert len(gold) == 2 * (nsteps + 1)
    for n in range(nsteps):
        op.set_time(True, time=n * dt, dt=dt, stage=0, A=((0.5,),), b=(1.0,), c=(0.5,), bdf=(1.0, -1.0))
        us = u.copy()
        res, jac = op.assemble_jacres(us, sol_prev=[u], sol_stage=[us])   # max nonlinear iters: 1 (the problem is linear)
        u = us + spla.spsolve(op.csr(jac).tocsc(), res)
        for f in ("E", "B"):
            got = op.l2_error([f + "[x]", f + "[y]", f + "[z]"], u)
            want = gold[(f, n + 1)]
            assert abs(got - want) <= _tol(want), (f, n + 1, got, want)


def test_maxwell_nonzero_ic_gold(oracle_lib):
    """regression/maxwell/NonzeroIC: 8^3 hex, E (HCURL) and B (HDIV) start from sin(pi x) sin(pi y) sin(pi z) in every component
    through the default `initial type: L2-projection` (solverManager_util.hpp:252-267): pins setInitial -- the projection right-hand
    side (assemblyManager_initial.hpp:260-311) and getMass -- for vector-valued bases at time 0, and one DIRK-1,2 step from a
    non-zero state at time 0.01."""
    import scipy.sparse.linalg as spla
    cfg, errs = _errs("maxwell/NonzeroIC")
    gold = {(e["field"], e["time"]): e["value"] for e in errs}
    assert set(gold) == {("E", 0.0), ("B", 0.0), ("E", 0.01), ("B", 0.01)}
    op = oracle_lib.OracleProblem(cfg)
    M, _ = op.weighted_mass([1.0, 1.0])
    u = spla.spsolve(op.csr(M).tocsc(), op.project_initial())
    for f in ("E", "B"):
        got = op.l2_error([f + "[x]", f + "[y]", f + "[z]"], u)
        assert abs(got - gold[(f, 0.0)]) <= _tol(gold[(f, 0.0)]), (f, got)
    dt = float(cfg["Solver"]["final time"]) / int(cfg["Solver"]["number of steps"])
    op.set_time(True, time=0.0, dt=dt, stage=0, A=((0.5,),), b=(1.0,), c=(0.5,), bdf=(1.0, -1.0))
    us = u.copy()
    res, jac = op.assemble_jacres(us, sol_prev=[u], sol_stage=[us])
    u1 = us + spla.spsolve(op.csr(jac).tocsc(), res)
    for f in ("E", "B"):
        got = op.l2_error([f + "[x]", f + "[y]", f + "[z]"], u1)
        assert abs(got - gold[(f, 0.01)]) <= _tol(gold[(f, 0.01)]), (f, got)
    op.set_time(False)
