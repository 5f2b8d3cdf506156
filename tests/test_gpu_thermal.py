"""GPU parity tests of the thermal path: CUDA kernels through the C ABI vs the CPU oracle on the
same seeded inputs.  Tolerance (north_star): 1e-12 relative on residual entries (to the vector
max-norm) and on matrix values (to the row max-norm); identical CSR pattern by construction (the
graph is an input)."""
import numpy as np
import pytest

import configs
import helpers

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(autouse=True, params=["true", "false"], ids=["jit", "aot"])
def kernel_build(request):
    """Every test runs twice: with the plan-specialised NVRTC kernel (jit=true: a failure to specialise is an error)
    and with the ahead-of-time kernel + bytecode expression interpreter."""
    helpers.DEFAULT_OPTIONS["jit"] = request.param
    yield request.param
    helpers.DEFAULT_OPTIONS.pop("jit", None)


def _device_arrays(op, u):
    import torch
    dev = torch.device("cuda:0")
    return (torch.from_numpy(np.ascontiguousarray(u)).to(dev), torch.zeros(op.num_dofs, dtype=torch.float64, device=dev),
            torch.zeros(op.nnz, dtype=torch.float64, device=dev))


def _check(op, plan, u, time=None, oracle_kw=None, tol=TOL):
    import torch
    d_u, d_res, d_jac = _device_arrays(op, u)
    res_ref, jac_ref = op.assemble_jacres(u, **(oracle_kw or {}))
    plan.assemble_jacres(d_u, d_res, d_jac, time=time)
    torch.cuda.synchronize()
    e_res = helpers.rel_err_vec(d_res.cpu().numpy(), res_ref)
    e_jac = helpers.rel_err_rows(d_jac.cpu().numpy(), jac_ref, op.rowptr)
    assert e_res < tol, "residual rel err %.3e" % e_res
    assert e_jac < tol, "Jacobian rel err %.3e" % e_jac
    return d_res.cpu().numpy(), d_jac.cpu().numpy(), res_ref, jac_ref


@pytest.mark.parametrize("name,cfg", [("2d", configs.THERMAL_2D), ("3d", configs.THERMAL_3D)])
def test_regression_decks_match_oracle(oracle_lib, product_lib, name, cfg):
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    _check(op, plan, np.zeros(op.num_dofs))            # the state the reference's first Newton step sees
    _check(op, plan, helpers.manufactured_state(op))   # non-trivial state


def test_gold_l2_error_through_cuda_path(oracle_lib, product_lib):
    """Solve the 2-D regression problem with the CUDA-assembled system and reproduce mrhyde.gold."""
    import scipy.sparse.linalg as spla
    import torch
    cfg = configs.THERMAL_2D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    d_u, d_res, d_jac = _device_arrays(op, np.zeros(op.num_dofs))
    plan.assemble_jacres(d_u, d_res, d_jac)
    torch.cuda.synchronize()
    u = spla.spsolve(op.csr(d_jac.cpu().numpy()).tocsc(), d_res.cpu().numpy())
    assert abs(op.l2_error(["T"], u) - configs.THERMAL_2D_GOLD["T"]) < 5e-9
    assert abs(op.l2_error(["grad(T)[x]", "grad(T)[y]"], u) - configs.THERMAL_2D_GOLD["grad(T)"]) < 5e-7


@pytest.mark.parametrize("dim", [2, 3])
def test_perturbed_mesh_variable_coefficients(oracle_lib, product_lib, dim):
    """Non-affine cells + x-dependent diffusion/source exercise the general per-point path."""
    base = configs.THERMAL_2D if dim == 2 else configs.THERMAL_3D
    upd = {"Mesh/perturb": 0.02, "Functions/thermal diffusion": "1.0+0.5*x*x+exp(-y)", "Mesh/NX": 9, "Mesh/NY": 7}
    if dim == 3:
        upd["Mesh/NZ"] = 5
    cfg = configs.variant(base, **upd)
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    assert plan.stat("n_affine") < op.num_elems
    _check(op, plan, helpers.manufactured_state(op))


def test_affine_sheared_mesh_constant_coefficients(oracle_lib, product_lib):
    """Stretched (non-cubic) brick keeps the table path honest about the metric tensor."""
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/xmax": 2.0, "Mesh/ymax": 0.5, "Mesh/zmin": -1.0, "Mesh/NX": 7, "Mesh/NY": 5, "Mesh/NZ": 6,
                                                  "Functions/thermal diffusion": "2.5"})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    assert plan.stat("n_affine") == op.num_elems
    _check(op, plan, helpers.manufactured_state(op))


def test_overwrite_mode_equals_accumulate_on_zeroed(oracle_lib, product_lib):
    import torch
    cfg = configs.THERMAL_3D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    u = helpers.manufactured_state(op)
    r1, j1, _, _ = _check(op, plan, u)
    plan.set_option("accumulate", "false")
    d_u, d_res, d_jac = _device_arrays(op, u)
    d_res.fill_(7.0)
    d_jac.fill_(-3.0)   # garbage the overwrite mode must replace
    plan.assemble_jacres(d_u, d_res, d_jac)
    torch.cuda.synchronize()
    assert np.array_equal(d_res.cpu().numpy(), r1)
    assert np.array_equal(d_jac.cpu().numpy(), j1)


def test_residual_only_and_jacobian_only(oracle_lib, product_lib):
    import torch
    cfg = configs.THERMAL_3D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    u = helpers.manufactured_state(op)
    r1, j1, _, _ = _check(op, plan, u)
    d_u, d_res, d_jac = _device_arrays(op, u)
    plan.assemble_res(d_u, d_res)
    torch.cuda.synchronize()
    res_scalar = op.assemble_res(u)                        # the ScalarT workset path of the oracle
    assert helpers.rel_err_vec(d_res.cpu().numpy(), res_scalar) < TOL
    assert np.array_equal(d_res.cpu().numpy(), r1)
    d_res.zero_()
    plan.assemble_jacres(d_u, None, d_jac, compute_residual=False)   # autotune flow, SURVEY 3.2
    torch.cuda.synchronize()
    assert np.array_equal(d_jac.cpu().numpy(), j1)


def test_run_to_run_reproducible(oracle_lib, product_lib):
    import torch
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 16, "Mesh/NY": 16, "Mesh/NZ": 16})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    u = helpers.manufactured_state(op)
    outs = []
    for _ in range(3):
        d_u, d_res, d_jac = _device_arrays(op, u)
        plan.assemble_jacres(d_u, d_res, d_jac)
        torch.cuda.synchronize()
        outs.append((d_res.cpu().numpy(), d_jac.cpu().numpy()))
    for r, j in outs[1:]:
        assert np.array_equal(r, outs[0][0]) and np.array_equal(j, outs[0][1])


def test_transient_bdf1_and_dirk(oracle_lib, product_lib):
    """Mass term + stage/BDF combination (computeSolnTransientSeeded)."""
    import torch
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/density": "2.0", "Functions/specific heat": "1.5",
                                                  "Functions/thermal source": "sin(t)*x+y*z"})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    rng = np.random.default_rng(3)
    u, up = rng.standard_normal(op.num_dofs), rng.standard_normal(op.num_dofs)
    dev = torch.device("cuda:0")
    d_up = torch.from_numpy(up).to(dev)
    for (A, b, c) in (([[1.0]], [1.0], [1.0]), ([[0.5]], [1.0], [0.5])):   # BWE, DIRK-1,2 (solverManager_setup.hpp:181-250)
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
        ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[d_up], sol_stage=[d_up])
        _check(op, plan, u, time=ts, oracle_kw=dict(sol_prev=[up], sol_stage=[u]))
    op.set_time(False)


def test_weak_dirichlet_and_neumann_boundaries(oracle_lib, product_lib):
    for dim, base in ((2, configs.THERMAL_2D), (3, configs.THERMAL_3D)):
        upd = {"Solver/use strong DBCs": False, "Physics/assemble boundary terms": True, "Mesh/NX": 6, "Mesh/NY": 5, "Mesh/perturb": 0.01,
               "Physics/Dirichlet conditions/T": {"left": "1.0+y", "top": "x*x"}, "Physics/Neumann conditions/T": {"right": "2.0*y-0.3"}}
        if dim == 3:
            upd["Mesh/NZ"] = 4
        cfg = configs.variant(base, **upd)
        op = oracle_lib.OracleProblem(cfg)
        plan = helpers.plan_from_oracle(op, cfg)
        _check(op, plan, helpers.manufactured_state(op))


def test_host_buffer_entry_point(oracle_lib, product_lib):
    cfg = configs.THERMAL_3D
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, indexed=True)
    u = helpers.manufactured_state(op)
    res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
    plan.assemble_jacres_host(u, res, jac)
    res_ref, jac_ref = op.assemble_jacres(u)
    assert helpers.rel_err_vec(res, res_ref) < TOL and helpers.rel_err_rows(jac, jac_ref, op.rowptr) < TOL


def test_device_expression_evaluator(oracle_lib, product_lib):
    cfg = configs.variant(configs.THERMAL_3D, **{"Functions/f1": "exp(-x*x)*cos(3*y)+z^2/(1.0+x)", "Functions/f2": "max(x,y)*(x<0.5)+sqrt(z)-abs(y-0.5)+f1*2",
                                                  "Mesh/NX": 4, "Mesh/NY": 4, "Mesh/NZ": 4})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg)
    for name in ("f1", "f2", "thermal source"):
        for grp in range(op.num_groups):
            wts, ip = op.group_info(grp)
            ref = op.eval_function(name, grp)
            xyz = np.stack([ip[0].ravel(), ip[1].ravel(), ip[2].ravel()], axis=1)
            got = plan.eval_function(name, xyz).reshape(ref.shape)
            assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("perturb", [0.0, 0.02])
def test_many_chains_and_segments(oracle_lib, product_lib, kernel_build, perturb):
    """Small columns and short sweep segments: every kind of halo (in-plane ring, level below a segment) is exercised."""
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 13, "Mesh/NY": 11, "Mesh/NZ": 9, "Mesh/perturb": perturb})
    op = oracle_lib.OracleProblem(cfg)
    for axis in (-1, 0):
        plan = helpers.plan_from_oracle(op, cfg, options={"column elements": 6, "min segment levels": 2, "sweep axis": axis})
        assert plan.stat("n_chains") > 8 and plan.stat("n_segments") > 1
        assert plan.stat("jit") == (1 if kernel_build == "true" else 0)
        _check(op, plan, helpers.manufactured_state(op))


@pytest.mark.parametrize("mesh", ["box", "sheared", "quad"])
def test_metric_ring_variants(oracle_lib, product_lib, kernel_build, mesh):
    """Parallelepiped cells + constant coefficients: the specialised build stages element metrics instead of local systems
    (volume_kernel.cuh, METRIC ring).  Generated pull code, the descriptor-driven pull (pull patterns = 0) and the full ring
    (ring = full) must all match the oracle; natural sides leave rows of every shape in the plan."""
    if kernel_build == "false":
        pytest.skip("the metric ring exists in the plan-specialised build only")
    upd = {"Mesh/NX": 13, "Mesh/NY": 11, "Mesh/NZ": 9, "Functions/thermal diffusion": "1.7", "Physics/assemble boundary terms": False,
           "Physics/Dirichlet conditions/T": {"left": "0.0", "top": "0.0"}}
    if mesh == "sheared":
        upd["Mesh/shear"] = 0.3
    cfg = configs.variant(configs.THERMAL_2D if mesh == "quad" else configs.THERMAL_3D, **upd)
    op = oracle_lib.OracleProblem(cfg)
    u = helpers.manufactured_state(op)
    for options in ({"ring": "metric"}, {"ring": "metric", "pull patterns": 0}, {"ring": "full"}):
        options.update({"column elements": 6, "min segment levels": 2})
        plan = helpers.plan_from_oracle(op, cfg, options=options)
        want = 0 if options["ring"] == "full" else (op.dim if mesh != "sheared" else op.dim * (op.dim + 1) // 2)
        assert plan.stat("metric_ring") == want and plan.stat("jit") == 1
        _check(op, plan, u)
        import torch
        for kw in (dict(compute_jacobian=False), dict(compute_residual=False)):   # residual-only / Jacobian-only builds
            d_u, d_res, d_jac = _device_arrays(op, u)
            res_ref, jac_ref = op.assemble_jacres(u)
            plan.assemble_jacres(d_u, d_res if kw.get("compute_residual", True) else None, d_jac if kw.get("compute_jacobian", True) else None, **kw)
            torch.cuda.synchronize()
            if kw.get("compute_residual", True):
                assert helpers.rel_err_vec(d_res.cpu().numpy(), res_ref) < TOL and float(d_jac.abs().max()) == 0.0
            else:
                assert helpers.rel_err_rows(d_jac.cpu().numpy(), jac_ref, op.rowptr) < TOL and float(d_res.abs().max()) == 0.0


@pytest.mark.parametrize("options", [{"flush": "flat"}, {"stage1": "early"}, {"pull group": 28}, {"flush": "flat", "ring": "metric"},
                                     {"stage1": "early", "prefetch": "lean", "prefetch records": True}, {"stage1": "late", "prefetch": True, "prefetch records": True}, {"column cache": False}, {"column cache": True, "sweep axis": 0}, {"column cache": "registers"}, {"threads": 32}, {"threads": 64, "column elements": 40}],
                         ids=["flush-flat", "stage1-early", "group-28", "metric-flat", "early-lean-records", "late-records", "no-column-cache", "x-sweep-cache", "column-cache-registers", "shared-values-1-warp", "shared-values-2-warps"])
def test_measured_build_alternatives_match_oracle(oracle_lib, product_lib, kernel_build, options):
    """Alternatives of the specialised build that were measured on the B200 and kept as options stay correct."""
    if kernel_build == "false":
        pytest.skip("options of the plan-specialised build")
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 13, "Mesh/NY": 11, "Mesh/NZ": 9})
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, options=dict({"column elements": 6, "min segment levels": 2}, **options))
    _check(op, plan, helpers.manufactured_state(op))
    if "threads" in options:   # the last warp holds elements: it evaluates the z-only sub-expressions of the source for the whole CTA
        assert plan.stat("step_shared_axes") == 4 and plan.stat("threads_per_block") == options["threads"]


@pytest.mark.parametrize("shear", [0.0, 0.3], ids=["box", "sheared"])
def test_metric_ring_transient(oracle_lib, product_lib, kernel_build, shear):
    """Transient builds of the metric ring (mass entry + time derivative staged, J = alpha_u K + alpha_t M in the pull): the layout a
    sheared mesh gets by default (first GPU run: profiles/r02_s1_first_gpu_runs.log)."""
    import torch
    if kernel_build == "false":
        pytest.skip("the metric ring exists in the plan-specialised build only")
    upd = {"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/density": "2.0", "Functions/specific heat": "1.5", "Functions/thermal source": "sin(t)*x+y*z"}
    if shear:
        upd["Mesh/shear"] = shear
    cfg = configs.variant(configs.THERMAL_3D, **upd)
    op = oracle_lib.OracleProblem(cfg)
    plan = helpers.plan_from_oracle(op, cfg, options={"ring": "metric"})
    rng = np.random.default_rng(3)
    u, up = rng.standard_normal(op.num_dofs), rng.standard_normal(op.num_dofs)
    d_up = torch.from_numpy(up).to(torch.device("cuda:0"))
    for (A, b, c) in (([[1.0]], [1.0], [1.0]), ([[0.5]], [1.0], [0.5])):
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
        ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[d_up], sol_stage=[d_up])
        _check(op, plan, u, time=ts, oracle_kw=dict(sol_prev=[up], sol_stage=[u]))
    op.set_time(False)


def test_full_size_properties(product_lib, kernel_build):
    """BASELINE configs[1] (128^3 hex-Q1 thermal) is too large for the oracle; check size-independent properties of the
    assembled system instead: K 1 = 0 on free rows, symmetry of the free-free block, res(u) = res(0) - J u (the problem is
    linear), the analytic interior stencil, identity rows on Dirichlet dofs, and equality of the two output modes."""
    import torch
    from mrhyde_b200.problems import ThermalBrick
    if kernel_build == "false":
        pytest.skip("full-size run uses the default (specialised) kernel")
    n = 128
    prob = ThermalBrick(3, [n, n, n], device=0, options={"accumulate": "false"})
    plan = prob.plan
    assert plan.stat("jit") == 1 and plan.stat("n_box") == prob.n_elem
    dev = torch.device("cuda:0")
    u = torch.from_numpy(prob.state()).to(dev)
    res = torch.full((prob.n_rows,), 5.0, dtype=torch.float64, device=dev)
    jac = torch.full((prob.nnz,), 5.0, dtype=torch.float64, device=dev)
    plan.assemble_jacres(u, res, jac)
    res0 = torch.empty_like(res)
    plan.assemble_jacres(torch.zeros_like(u), res0, None, compute_jacobian=False)
    torch.cuda.synchronize()
    J = torch.sparse_csr_tensor(torch.from_numpy(prob.rowptr).to(dev), torch.from_numpy(prob.colind.astype(np.int64)).to(dev), jac, size=(prob.n_rows, prob.n_rows))
    free = torch.from_numpy(prob.is_fixed == 0).to(dev)
    scale = float(jac.abs().max())
    ones = torch.ones(prob.n_rows, 1, dtype=torch.float64, device=dev)
    rowsum = (J @ ones).squeeze(1)
    assert float(rowsum[free].abs().max()) < 1e-12 * scale                 # K 1 = 0
    assert float((rowsum[~free] - 1.0).abs().max()) == 0.0                # identity rows on Dirichlet dofs
    assert float(res[~free].abs().max()) == 0.0
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(prob.n_rows, 1, dtype=torch.float64, generator=g).to(dev) * free.unsqueeze(1)
    y = torch.randn(prob.n_rows, 1, dtype=torch.float64, generator=g).to(dev) * free.unsqueeze(1)
    a, b = float((x * (J @ y)).sum()), float((y * (J @ x)).sum())
    assert abs(a - b) < 1e-11 * max(abs(a), abs(b), 1.0)                  # symmetric free-free block
    lin = res0 - (J @ u.unsqueeze(1)).squeeze(1)
    assert float((res - lin)[free].abs().max()) < 1e-12 * float(res.abs().max())   # res(u) = res(0) - J u
    h = 1.0 / n
    nn = n + 1
    c = 5 + 7 * nn + 9 * nn * nn
    row = jac[prob.rowptr[c]:prob.rowptr[c + 1]].cpu().numpy()
    cols = prob.colind[prob.rowptr[c]:prob.rowptr[c + 1]]
    for v, cc in zip(row, cols):
        d = cc - c
        dk = int(round(d / (nn * nn)))
        dj = int(round((d - dk * nn * nn) / nn))
        di = d - dk * nn * nn - dj * nn
        want = {0: 8 * h / 3, 1: 0.0, 2: -h / 6, 3: -h / 12}[abs(di) + abs(dj) + abs(dk)]
        assert abs(v - want) < 1e-15
    # reference contract (sum into caller-zeroed arrays) gives the same values
    plan.set_option("accumulate", "true")
    res2, jac2 = torch.zeros_like(res), torch.zeros_like(jac)
    plan.assemble_jacres(u, res2, jac2)
    torch.cuda.synchronize()
    assert torch.equal(res2, res) and torch.equal(jac2, jac)


def test_overwrite_mode_without_strong_dbcs_leaves_fixed_rows_zero(oracle_lib, product_lib, kernel_build):
    """`use strong DBCs: false` with is_fixed rows present: the reference's scatter still skips those rows but never calls
    setJacobianConstraints (assemblyManager_constraints.hpp:250), so they stay zero rows -- in overwrite mode as in accumulate mode
    (ADVICE round 1).  Also: plan.warmup builds the variants up front, so these calls do not compile."""
    import torch
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4})
    op = oracle_lib.OracleProblem(cfg)
    u = helpers.manufactured_state(op)
    outs = {}
    for kernel in ("sweep", "general"):
        for acc in ("true", "false"):
            plan = helpers.plan_from_oracle(op, cfg, options={"use strong DBCs": "false", "accumulate": acc, "kernel": kernel})
            plan.warmup(transient=False, compute_jacobian=True, compute_residual=True)
            d_u, d_res, d_jac = _device_arrays(op, u)
            if acc == "false":
                d_res.fill_(7.0); d_jac.fill_(7.0)
            plan.assemble_jacres(d_u, d_res, d_jac)
            torch.cuda.synchronize()
            outs[(kernel, acc)] = (d_res.cpu().numpy(), d_jac.cpu().numpy())
    fixed = op.is_fixed.astype(bool)
    rows = np.repeat(np.arange(op.num_dofs), np.diff(op.rowptr))
    for key, (res, jac) in outs.items():
        assert np.all(jac[fixed[rows]] == 0.0) and np.all(res[fixed] == 0.0), key
        assert np.array_equal(res, outs[("sweep", "true")][0]) or helpers.rel_err_vec(res, outs[("sweep", "true")][0]) < TOL
        assert helpers.rel_err_rows(jac, outs[("sweep", "true")][1], op.rowptr) < TOL


@pytest.mark.parametrize("kernel", ["sweep", "general"])
def test_point_constraints(oracle_lib, product_lib, kernel):
    """disc->point_dofs: identity Jacobian rows after the assembly, residual untouched (assemblyManager_constraints.hpp:97-116, 261-266),
    in overwrite and in accumulate mode, both kernels."""
    import torch
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 7, "Mesh/NY": 6, "Mesh/NZ": 5})
    op = oracle_lib.OracleProblem(cfg)
    pts = np.array([3, 40, op.num_dofs - 1], dtype=np.int32)
    op.set_point_dofs(pts)
    for accumulate in (False, True):
        plan = helpers.plan_from_oracle(op, cfg, options={"kernel": kernel, "accumulate": "true" if accumulate else "false"})
        plan.set_point_dofs(pts)
        _check(op, plan, helpers.manufactured_state(op))
        assert plan.stat("kernel_launches_per_assemble") >= 2
        plan.set_point_dofs(np.zeros(0, dtype=np.int32))
    op.set_point_dofs(np.zeros(0, dtype=np.int32))
    _check(op, plan, helpers.manufactured_state(op))


def test_fix_zero_rows_after_the_sweep_kernel(oracle_lib, product_lib):
    """Solver: fix zero rows (assemblyManager_jacres.hpp:609-626) as a post-pass of the sweep kernel: with a vanishing diffusion coefficient
    every Jacobian row is empty and gets a unit diagonal; with the regular deck nothing changes."""
    for diff in ("0.0", "1.0"):
        cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 7, "Mesh/NY": 6, "Mesh/NZ": 5, "Solver/fix zero rows": True, "Functions/thermal diffusion": diff})
        op = oracle_lib.OracleProblem(cfg)
        plan = helpers.plan_from_oracle(op, cfg)
        assert plan.stat("general") == 0
        _, jac, _, _ = _check(op, plan, helpers.manufactured_state(op))
        if diff == "0.0":
            assert abs(op.csr(jac) - __import__("scipy.sparse").sparse.identity(op.num_dofs)).max() == 0.0


def test_cpp_host_example_matches_oracle(oracle_lib, product_lib, tmp_path):
    """examples/host_assemble.cpp (C++ against the C ABI, host buffers) on the GPU: residual norm and Jacobian trace of the reference's
    2D_verification deck on an 8 x 8 mesh agree with the oracle (both are invariant under the different dof numbering)."""
    import subprocess
    from test_abi_cpu import _build_example
    exe = _build_example(tmp_path)
    out = subprocess.run([exe, "8", "0"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("|res|_2")][0].replace("=", " ").split()
    rnorm, trace = float(line[1]), float(line[3])
    cfg = configs.variant(configs.THERMAL_2D, **{"Mesh/NX": 8, "Mesh/NY": 8})
    op = oracle_lib.OracleProblem(cfg)
    res, jac = op.assemble_jacres(np.zeros(op.num_dofs))
    assert abs(rnorm - np.linalg.norm(res)) <= 1e-12 * np.linalg.norm(res)
    assert abs(trace - op.csr(jac).diagonal().sum()) <= 1e-12 * abs(trace)
