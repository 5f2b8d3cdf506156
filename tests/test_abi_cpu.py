"""CPU-side checks of the C-ABI library (no GPU, no compute calls on a device): the shared object loads,
exports every symbol include/mrhyde_b200.h declares, validates its inputs like the header says, and its
host-side plan analysis (expression compiler, patches, scatter programs) is consistent."""
import math
import os
import re

import numpy as np
import pytest

import configs
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mrhyde_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrhyde_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(product_lib):
    L = product_lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(L, s), "libmrhyde_b200.so does not export %s" % s
    assert sorted(product_lib.SYMBOLS) == declared, "capi.SYMBOLS and the header disagree"
    assert b"sm_100a" in L.mrhyde_b200_version()


def test_no_cpu_fallback_without_a_device(product_lib):
    """On a box without CUDA plan_create(device >= 0) must fail loudly; with CUDA this is covered by the gpu tests."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mrhyde_b200 import inline_mesh as im
    pts, wts, val, grad = im.q1_reference(3)
    with pytest.raises(product_lib.MrhydeB200Error) as ei:
        product_lib.AssemblyPlan("thermal", 3, ["T"], [0], [dict(type="HGRAD", order=1, card=8, val=val, grad=grad)], 8,
                                 np.arange(8, dtype=np.int32).reshape(1, 8), pts, wts, device=0)
    assert ei.value.code == product_lib.ERR_CUDA and "no CPU path" in ei.value.message


def _host_plan(oracle_lib, cfg, **kw):
    op = oracle_lib.OracleProblem(cfg)
    return op, helpers.plan_from_oracle(op, cfg, device=-1, **kw)


def test_host_only_plan_refuses_to_assemble(oracle_lib, product_lib):
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 5, "Mesh/NY": 4, "Mesh/NZ": 3})
    op, plan = _host_plan(oracle_lib, cfg)
    assert plan.stat("n_elem") == op.num_elems and plan.stat("nnz") == op.nnz and plan.stat("n_affine") == op.num_elems
    u, r, j = np.zeros(op.num_dofs), np.zeros(op.num_dofs), np.zeros(op.nnz)
    with pytest.raises(product_lib.MrhydeB200Error) as ei:
        plan.assemble_jacres_host(u, r, j)
    assert ei.value.code == product_lib.ERR_STATE


def test_argument_validation(oracle_lib, product_lib):
    cfg = configs.variant(configs.THERMAL_2D, **{"Mesh/NX": 4, "Mesh/NY": 4})
    op = oracle_lib.OracleProblem(cfg)
    rb = op.ref_basis(0)
    mk = lambda: product_lib.AssemblyPlan("thermal", 2, ["T"], [0], [dict(type="HGRAD", order=1, card=4, val=rb["val"], grad=rb["grad"])], 4,
                                          op.offsets, op.qpts, op.qwts, device=-1)
    p = mk()
    with pytest.raises(product_lib.MrhydeB200Error):
        p.set_option("no such option", "1")                      # unknown keys are an error, never ignored
    with pytest.raises(product_lib.MrhydeB200Error):
        p.finalize()                                             # before set_mesh / set_graph
    p.set_mesh(op.elem_nodes, op.lids)
    bad = op.colind.copy()
    bad[[0, 1]] = bad[[1, 0]]
    with pytest.raises(product_lib.MrhydeB200Error):
        p.set_graph(op.rowptr, bad, op.is_fixed)                 # columns must ascend (fillComplete'd graph)
    p.set_function("thermal source", "sin(x")                    # reference: unclosed parenthesis
    p.set_graph(op.rowptr, op.colind, op.is_fixed)
    with pytest.raises(product_lib.MrhydeB200Error) as ei:
        p.finalize()
    assert ei.value.code == product_lib.ERR_PARSE
    q = mk()
    q.set_mesh(op.elem_nodes, op.lids)
    short = op.rowptr[:-3].copy()                                # graph with fewer rows than the LIDs reference
    q.set_graph(short, op.colind[: short[-1]], op.is_fixed[: len(short) - 1])
    with pytest.raises(product_lib.MrhydeB200Error):
        q.finalize()
    with pytest.raises(product_lib.MrhydeB200Error) as ei:
        # (arrays sized for the 3-D descriptor they are passed with: the library reads nqp * dim coordinates)
        product_lib.AssemblyPlan("maxwell", 3, ["E"], [0], [dict(type="HGRAD", order=1, card=4, val=rb["val"], grad=np.zeros((4, len(op.qwts), 3)))], 4,
                                 op.offsets, np.zeros((len(op.qwts), 3)), op.qwts, device=-1).finalize()
    assert ei.value.code in (product_lib.ERR_STATE, product_lib.ERR_UNSUPPORTED)


EXPRS = {
    "a": "1.0", "b": "2.0", "Ha": "1.0", "gtst": "a+b",
    "f0": "sin(x+y+t)", "f1": "x+exp(y)", "f2": "8*(pi^2)*sin(2*pi*x+1)*sin(2*pi*y+1)", "f3": "-exp(x)", "f4": "(a-sin(x))^(2+b)",
    "f5": "(a+2.0)*(b-pi)", "f6": "(a+b) + ((x+y)*a - 2.0)", "f7": "exp(-(a+b)^2)", "f8": "sin(gtst)", "f9": "8*pi^2", "f10": "min(a,b)",
    "f11": "a <= b", "f13": "(1+exp(-2.0*Ha))/(2.0*exp(-1.0*Ha))",
    "g0": "12*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)", "g1": "1.0+0.5*x*x+exp(-y)", "g2": "max(x,y)-min(y,z)+abs(x-0.5)",
    "g3": "sqrt(x*x+y*y)/(1.0+z)", "g4": "(x<0.5)*2.0+(x>=0.5)*3.0", "g5": "cosh(x)-sinh(y)+tan(0.3*z)+log(1.0+x)", "g6": "n[x]*x+n[y]*y-n[z]",
    "g7": "2.0^x^2", "g8": "x-y-z", "g9": "x/y/2.0",
}


def _py_eval(expr, fn, v):
    """Independent evaluation with Python semantics arranged to match the reference's left-to-right rules."""
    x, y, z, t, nx, ny, nz = v
    env = {"sin": math.sin, "cos": math.cos, "exp": math.exp, "log": math.log, "tan": math.tan, "abs": abs, "max": max, "min": min,
           "sqrt": math.sqrt, "sinh": math.sinh, "cosh": math.cosh, "pi": math.pi, "x": x, "y": y, "z": z, "t": t, "nx": nx, "ny": ny, "nz": nz}
    for k in ("a", "b", "Ha"):
        env[k] = float(fn[k])
    env["gtst"] = env["a"] + env["b"]
    e = expr.replace("n[x]", "nx").replace("n[y]", "ny").replace("n[z]", "nz")
    if expr == "2.0^x^2":
        return (2.0 ** x) ** 2                                   # the reference applies ^ left to right
    return float(eval(e.replace("^", "**"), {"__builtins__": {}}, env))


def test_expression_compiler_matches_reference_semantics(product_lib):
    rng = np.random.default_rng(7)
    pts = rng.uniform(0.05, 0.95, size=(40, 7))
    for name, expr in EXPRS.items():
        got = product_lib.expr_eval_host(EXPRS, name, pts)
        want = np.array([_py_eval(expr, EXPRS, v) for v in pts])
        assert np.allclose(got, want, rtol=1e-14, atol=1e-14), (name, expr, np.max(np.abs(got - want)))
    # constant trees are folded at set-up like the reference does (functionManager_create.hpp:519-534)
    assert "const" in product_lib.expr_disassemble(EXPRS, "f13").lower()
    for bad in ("sin(x", "x)+1", "foo(x)", "1.0+undefined_name"):
        with pytest.raises(product_lib.MrhydeB200Error):
            product_lib.expr_eval_host({"bad": bad}, "bad", pts)


@pytest.mark.parametrize("dim,perturb", [(2, 0.0), (3, 0.0), (3, 0.03)])
def test_scatter_programs_equal_serial_element_loop(oracle_lib, product_lib, dim, perturb):
    """The plan's pull-scatter programs (owner-computes patches + templates) must reproduce the reference's serial
    element loop (scatter.hpp:162-278) for arbitrary staged element matrices, including the fixed-row rules."""
    base = configs.THERMAL_2D if dim == 2 else configs.THERMAL_3D
    upd = {"Mesh/NX": 11, "Mesh/NY": 9, "Mesh/perturb": perturb}
    if dim == 3:
        upd["Mesh/NZ"] = 7
    cfg = configs.variant(base, **upd)
    op, plan = _host_plan(oracle_lib, cfg, options={"column elements": 4, "min segment levels": 2})
    nd = op.ndof_elem
    # staged layout: upper triangle + residual, or -- box meshes with constant coefficients, the class ring -- one value per
    # class of entries + residual; either way local entry (i, j) is stage[e, kmap[i, j]]
    kmap, rmap = plan.debug_stage_map(nd)
    assert plan.stat("stage_len") == rmap.max() + 1 and (plan.stat("class_ring") > 0) == (perturb == 0.0)
    assert np.array_equal(kmap, kmap.T)
    rng = np.random.default_rng(3)
    stage = rng.standard_normal((op.num_elems, plan.stat("stage_len")))
    # reference loop
    res_ref = np.zeros(op.num_dofs)
    jac_ref = np.zeros(op.nnz)
    fixed = op.is_fixed.astype(bool)
    for e in range(op.num_elems):
        l = op.lids[e]
        for i in range(nd):
            r = l[i]
            if fixed[r]:
                continue
            res_ref[r] -= stage[e, rmap[i]]
            cols = op.colind[op.rowptr[r]:op.rowptr[r + 1]]
            for j in range(nd):
                jac_ref[op.rowptr[r] + np.searchsorted(cols, l[j])] += stage[e, kmap[i, j]]
    for accumulate in (1, 0):
        res = np.full(op.num_dofs, 0.0 if accumulate else 7.0)
        jac = np.full(op.nnz, 0.0 if accumulate else 7.0)
        plan.debug_scatter_host(stage, accumulate, res, jac)
        rr, jr = res_ref.copy(), jac_ref.copy()
        if not accumulate:   # overwrite mode also applies dofConstraints: identity rows, zero residual
            for r in np.nonzero(fixed)[0]:
                jr[op.rowptr[r]:op.rowptr[r + 1]] = (op.colind[op.rowptr[r]:op.rowptr[r + 1]] == r)
        assert np.allclose(res, rr, rtol=0, atol=1e-13)
        assert np.allclose(jac, jr, rtol=0, atol=1e-13)
    assert plan.stat("n_chains") > 1 and plan.stat("n_patterns") < plan.stat("n_rows")


METRIC_CASES = {
    "2d_box": (configs.THERMAL_2D, {"Mesh/NX": 11, "Mesh/NY": 9}, {"column elements": 4, "min segment levels": 2}),
    "3d_box": (configs.THERMAL_3D, {"Mesh/NX": 9, "Mesh/NY": 7, "Mesh/NZ": 6}, {"column elements": 6, "min segment levels": 2}),
    "3d_stretched": (configs.THERMAL_3D, {"Mesh/xmax": 2.0, "Mesh/ymax": 0.5, "Mesh/zmin": -1.0, "Mesh/NX": 7, "Mesh/NY": 5, "Mesh/NZ": 6,
                                          "Functions/thermal diffusion": "2.5"}, {}),
    "3d_sheared": (configs.THERMAL_3D, {"Mesh/shear": 0.3, "Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/thermal diffusion": "0.7"},
                   {"column elements": 5, "min segment levels": 2}),
    "2d_sheared": (configs.THERMAL_2D, {"Mesh/shear": 0.25, "Mesh/NX": 8, "Mesh/NY": 6}, {}),
    "3d_natural_sides": (configs.THERMAL_3D, {"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Physics/assemble boundary terms": False,
                                              "Physics/Dirichlet conditions/T": {"left": "0.0", "top": "0.0"}}, {"column elements": 4}),
}


@pytest.mark.parametrize("case", sorted(METRIC_CASES))
def test_metric_ring_host_replay_matches_oracle(oracle_lib, product_lib, case):
    """Plans whose cells are all parallelepipeds with constant coefficients stage the element metric instead of the local
    system (kernel_abi.h, METRIC ring).  The host replay of that plan -- same source words and formulas as the kernel --
    must reproduce the oracle's residual and Jacobian to 1e-12."""
    base, upd, opts = METRIC_CASES[case]
    cfg = configs.variant(base, **upd)
    op, plan = _host_plan(oracle_lib, cfg, options=dict(opts, ring="metric"))   # ring=auto picks the class ring on box meshes
    assert plan.stat("metric_ring") == (op.dim if "sheared" not in case else op.dim * (op.dim + 1) // 2)
    u = helpers.manufactured_state(op)
    res_ref, jac_ref = op.assemble_jacres(u)
    for accumulate in (1, 0):
        res = np.full(op.num_dofs, 0.0 if accumulate else 7.0)
        jac = np.full(op.nnz, 0.0 if accumulate else 7.0)
        plan.debug_metric_host(u, accumulate, res, jac)
        jr = jac_ref.copy()
        if accumulate:   # dofConstraints (J(d,d) = 1 on fixed rows) is a separate kernel in accumulate mode
            for r in np.nonzero(op.is_fixed)[0]:
                jr[op.rowptr[r]:op.rowptr[r + 1]] = 0.0
        assert helpers.rel_err_vec(res, res_ref) < 1e-12
        assert helpers.rel_err_rows(jac, jr, op.rowptr) < 1e-12


def test_metric_ring_host_replay_transient(oracle_lib, product_lib):
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/density": "2.0", "Functions/specific heat": "1.5",
                                                  "Functions/thermal source": "sin(t)*x+y*z"})
    op, plan = _host_plan(oracle_lib, cfg, options={"ring": "metric"})
    rng = np.random.default_rng(3)
    u, up = rng.standard_normal(op.num_dofs), rng.standard_normal(op.num_dofs)
    for (A, b, c) in (([[1.0]], [1.0], [1.0]), ([[0.5]], [1.0], [0.5])):
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
        ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[up], sol_stage=[up])
        res_ref, jac_ref = op.assemble_jacres(u, sol_prev=[up], sol_stage=[u])
        res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
        plan.debug_metric_host(u, 0, res, jac, time=ts)
        assert helpers.rel_err_vec(res, res_ref) < 1e-12
        assert helpers.rel_err_rows(jac, jac_ref, op.rowptr) < 1e-12
    op.set_time(False)


CLASS_CASES = {k: v for k, v in METRIC_CASES.items() if "sheared" not in k}


@pytest.mark.parametrize("case", sorted(CLASS_CASES))
def test_class_ring_host_replay_matches_oracle(oracle_lib, product_lib, case):
    """The layout the headline workload runs on (boxes + constant coefficients, ring=auto): one local-matrix value per class of
    entries + the residual per element.  Host replay of that plan -- the kernel's formulas, the plan's class map and scatter
    programs -- against the oracle to 1e-12, steady and transient."""
    base, upd, opts = CLASS_CASES[case]
    cfg = configs.variant(base, **upd)
    op, plan = _host_plan(oracle_lib, cfg, options=opts)
    assert plan.stat("class_ring") == (8 if op.dim == 3 else 4) and plan.stat("metric_ring") == 0
    u = helpers.manufactured_state(op)
    res_ref, jac_ref = op.assemble_jacres(u)
    for accumulate in (1, 0):
        res = np.full(op.num_dofs, 0.0 if accumulate else 7.0)
        jac = np.full(op.nnz, 0.0 if accumulate else 7.0)
        plan.debug_class_host(u, accumulate, res, jac)
        jr = jac_ref.copy()
        if accumulate:   # dofConstraints (J(d,d) = 1 on fixed rows) is a separate kernel in accumulate mode
            for r in np.nonzero(op.is_fixed)[0]:
                jr[op.rowptr[r]:op.rowptr[r + 1]] = 0.0
        assert helpers.rel_err_vec(res, res_ref) < 1e-12
        assert helpers.rel_err_rows(jac, jr, op.rowptr) < 1e-12


def test_class_ring_host_replay_transient(oracle_lib, product_lib):
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/density": "2.0", "Functions/specific heat": "1.5",
                                                  "Functions/thermal source": "sin(t)*x+y*z"})
    op, plan = _host_plan(oracle_lib, cfg)
    assert plan.stat("class_ring") == 8
    rng = np.random.default_rng(3)
    u, up = rng.standard_normal(op.num_dofs), rng.standard_normal(op.num_dofs)
    for (A, b, c) in (([[1.0]], [1.0], [1.0]), ([[0.5]], [1.0], [0.5])):
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0))
        ts = helpers.TimeSpec(time=0.3, deltat=0.01, stage=0, A=A, b=b, c=c, bdf=(1.0, -1.0), sol_prev=[up], sol_stage=[up])
        res_ref, jac_ref = op.assemble_jacres(u, sol_prev=[up], sol_stage=[u])
        res, jac = np.zeros(op.num_dofs), np.zeros(op.nnz)
        plan.debug_class_host(u, 0, res, jac, time=ts)
        assert helpers.rel_err_vec(res, res_ref) < 1e-12
        assert helpers.rel_err_rows(jac, jac_ref, op.rowptr) < 1e-12
    op.set_time(False)


def test_ring_layout_selection(oracle_lib, product_lib):
    """ring=auto: class ring on axis-aligned boxes, metric ring on sheared parallelepipeds (both need constant coefficients),
    full local systems otherwise; asking for a layout the plan cannot have is an error."""
    small = {"Mesh/NX": 5, "Mesh/NY": 4, "Mesh/NZ": 3}
    stat = lambda plan: (plan.stat("class_ring"), plan.stat("metric_ring"), plan.stat("stage_len"))
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **small))
    assert stat(plan) == (8, 0, 16)
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_2D, **small))
    assert stat(plan) == (4, 0, 8)
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **dict(small, **{"Mesh/shear": 0.2})))
    assert stat(plan) == (0, 6, 44)
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **dict(small, **{"Mesh/perturb": 0.02})))
    assert stat(plan) == (0, 0, 44)
    cfg = configs.variant(configs.THERMAL_3D, **dict(small, **{"Functions/thermal diffusion": "1.0+x"}))
    op, plan = _host_plan(oracle_lib, cfg)
    assert stat(plan) == (0, 0, 44)
    for ring in ("metric", "class"):
        with pytest.raises(product_lib.MrhydeB200Error):
            _host_plan(oracle_lib, cfg, options={"ring": ring})
    with pytest.raises(product_lib.MrhydeB200Error):
        _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **dict(small, **{"Mesh/shear": 0.2})), options={"ring": "class"})
    for ring, want in (("full", (0, 0, 44)), ("metric", (0, 3, 44)), ("class", (8, 0, 16))):
        op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **small), options={"ring": ring})
        assert stat(plan) == want
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **small), options={"jit": "false"})   # no specialised build: full ring
    assert stat(plan) == (0, 0, 44)


def test_compressed_ring_kernels_compile(oracle_lib, product_lib, tmp_path):
    """NVRTC compiles the class-ring and metric-ring builds for sm_100a without a device (steady variant; generated pull code
    for the frequent patterns)."""
    for base, upd, ring, marks in ((configs.THERMAL_3D, {"Mesh/NX": 9, "Mesh/NY": 7, "Mesh/NZ": 6}, "metric", ("#define MRH_JIT_METRIC 1", "mrh_pull_metric_special")),
                                   (configs.THERMAL_3D, {"Mesh/NX": 9, "Mesh/NY": 7, "Mesh/NZ": 6}, "class", ("#define MRH_JIT_CLASS_NC 8", "mrh_pull_special")),
                                   (configs.THERMAL_2D, {"Mesh/NX": 9, "Mesh/NY": 7}, "metric", ("#define MRH_JIT_METRIC 1",)),
                                   (configs.THERMAL_2D, {"Mesh/NX": 9, "Mesh/NY": 7}, "class", ("#define MRH_JIT_CLASS_NC 4",)),
                                   (configs.THERMAL_3D, {"Mesh/NX": 6, "Mesh/NY": 5, "Mesh/NZ": 4, "Mesh/shear": 0.3}, "auto", ("#define MRH_JIT_METRIC_NG 6",))):
        cfg = configs.variant(base, **upd)
        op, plan = _host_plan(oracle_lib, cfg, options={"ring": ring})
        src = tmp_path / "k.cu"
        plan.debug_jit(source_path=str(src))
        text = src.read_text()
        for m in marks:
            assert m in text


def test_build_option_variants_compile(oracle_lib, product_lib, tmp_path):
    """Every tuning option of the specialised build still yields a translation unit NVRTC accepts (the options are measured
    alternatives kept for experiments); unknown values are rejected at finalize."""
    cfg = configs.variant(configs.THERMAL_3D, **{"Mesh/NX": 9, "Mesh/NY": 7, "Mesh/NZ": 6})
    variants = [{"flush": "flat"}, {"flush": "row", "flush unroll": 4}, {"pull group": 28}, {"pull patterns": 0}, {"stage1": "early"},
                {"stage1": "early", "stage2": "early"}, {"stage1": "early", "prefetch": "lean", "prefetch records": True}, {"tables": "literal"}, {"stagger ns": 3000}, {"max registers": 96},
                {"ring": "metric", "pull group": 4}, {"ring": "full", "flush": "flat"}]
    for options in variants:
        op, plan = _host_plan(oracle_lib, cfg, options=options)
        log = plan.debug_jit(source_path=str(tmp_path / "k.cu"))
        assert "error" not in log.lower()
    with pytest.raises(product_lib.MrhydeB200Error):
        _host_plan(oracle_lib, cfg, options={"ring": "sparse"})
    with pytest.raises(product_lib.MrhydeB200Error):
        _host_plan(oracle_lib, cfg, options={"no such option": 1})


@pytest.mark.parametrize("ring,shear", [("class", 0.0), ("metric", 0.0), ("metric", 0.3), ("full", 0.0)])
def test_every_build_variant_compiles(oracle_lib, product_lib, ring, shear):
    """The specialised kernel is built per (steady | transient, output mode): a sample of the twelve builds of every ring layout compiles for sm_100a."""
    upd = {"Mesh/NX": 7, "Mesh/NY": 5, "Mesh/NZ": 4, "Functions/density": "2.0", "Functions/specific heat": "1.5", "Functions/thermal source": "sin(t)*x+y*z"}
    if shear:
        upd["Mesh/shear"] = shear
    op, plan = _host_plan(oracle_lib, configs.variant(configs.THERMAL_3D, **upd), options={"ring": ring})
    import os
    from mrhyde_b200.capi import MrhydeB200Error
    os.environ.pop("MRHYDE_B200_DEBUG_OPTIONS", None)
    with pytest.raises(MrhydeB200Error):        # kernel-debugging keys are refused unless the environment allows them
        plan.set_option("debug transient", 1)
    os.environ["MRHYDE_B200_DEBUG_OPTIONS"] = "1"
    for transient in (0, 1):
        for mode in (1, 3, 6):   # residual only, residual + Jacobian overwrite, Jacobian only accumulate
            plan.set_option("debug transient", transient)
            plan.set_option("debug mode", mode)
            log = plan.debug_jit()
            assert "error" not in log.lower(), (transient, mode, log[-400:])


def test_ghost_row_chains_come_first(product_lib):
    """Multi-rank plans number the chains that complete ghost rows first, so that option "overlap halo" can launch them, start
    the exchange and launch the rest: chains [0, n_early) write every ghost row, and every row is written by exactly one chain."""
    from mrhyde_b200.problems import ThermalBrick
    for rank, world in ((0, 2), (1, 3), (2, 3), (0, 1)):
        prob = ThermalBrick(3, (12, 10, 8), device=-1, rank=rank, nranks=world, options={"column elements": 16, "min segment levels": 2})
        plan = prob.plan
        n_early, n_chains = plan.stat("n_early_chains"), plan.stat("n_chains")
        ghosts = np.arange(prob.n_rows) >= prob.n_owned
        if not ghosts.any():
            assert n_early == 0
            continue
        assert 0 < n_early < n_chains
        early = plan.debug_chain_rows(0, n_early, prob.n_rows)
        late = plan.debug_chain_rows(n_early, n_chains, prob.n_rows)
        assert early[ghosts].all() and not late[ghosts].any()
        assert not (early & late).any() and (early | late).all()


def test_column_cache_flags(product_lib):
    """Chains of an extruded brick keep their x / y intervals from step to step (bitwise): the plan marks them and the generated source
    function caches the one-coordinate sub-expressions of those axes; a perturbed mesh has no such chain."""
    from mrhyde_b200.problems import ThermalBrick
    brick = ThermalBrick(3, [12, 10, 9], device=-1, options={"column elements": 16, "min segment levels": 2})
    assert brick.plan.stat("column_cache_axes") == 3 and brick.plan.stat("n_invariant_chains") > 0
    assert brick.plan.stat("step_shared_axes") == 4   # all elements of a sweep level share their z interval: one warp evaluates sin(2 pi z) for the CTA
    assert ThermalBrick(3, [12, 10, 9], device=-1, options={"column elements": 16, "min segment levels": 2, "column cache": "registers"}).plan.stat("step_shared_axes") == 0
    assert ThermalBrick(3, [12, 10, 9], device=-1, options={"column elements": 16, "min segment levels": 2, "column cache": False}).plan.stat("column_cache_axes") == 0
    bent = ThermalBrick(3, [12, 10, 9], device=-1, perturb=0.1, options={"column elements": 16, "min segment levels": 2})
    assert bent.plan.stat("column_cache_axes") == 0 and bent.plan.stat("n_invariant_chains") == 0
    x_sweep = ThermalBrick(3, [12, 10, 9], device=-1, options={"column elements": 16, "min segment levels": 2, "sweep axis": 0})
    assert x_sweep.plan.stat("column_cache_axes") == 6 and x_sweep.plan.stat("step_shared_axes") == 1


HERE = os.path.dirname(os.path.abspath(__file__))


def _build_example(tmp_path):
    import subprocess
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "host_assemble")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "host_assemble.cpp"),
                           "-L", os.path.join(root, "mrhyde_b200"), "-lmrhyde_b200", "-Wl,-rpath," + os.path.join(root, "mrhyde_b200"), "-o", exe])
    return exe


def test_header_is_valid_c_and_cpp(tmp_path, product_lib):
    """include/mrhyde_b200.h is the boundary a C++ (or C) host binds to: it must compile on its own in both languages."""
    import subprocess
    root = os.path.dirname(HERE)
    src = tmp_path / "inc.c"
    src.write_text('#include "mrhyde_b200.h"\nint main(void) { return mrhyde_b200_version() == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", os.path.join(root, "include"), str(src)])


def test_cpp_host_example_runs_against_the_library(tmp_path, product_lib):
    """examples/host_assemble.cpp drives the C ABI from C++ with no Python in between; without a GPU it builds the host-only plan and a
    device plan fails loudly (no CPU path)."""
    import subprocess
    exe = _build_example(tmp_path)
    out = subprocess.run([exe, "8", "-1"], capture_output=True, text=True)
    assert out.returncode == 0 and "64 elements, 81 rows, 625 non-zeros" in out.stdout and "host-only plan" in out.stdout
    import torch
    if not torch.cuda.is_available():
        bad = subprocess.run([exe, "8", "0"], capture_output=True, text=True)
        assert bad.returncode != 0 and "no CPU path" in bad.stderr


def test_plan_construction_is_deterministic(product_lib):
    """The sweep plan is built by several host threads (patterns are de-duplicated under a lock); the schedule must not depend on their
    timing: every array of the plan hashes the same over repeated builds, on a mesh with many gather patterns too."""
    from mrhyde_b200.problems import ThermalBrick
    for n, perturb in (([33, 17, 29], 0.05), ([20, 20, 20], 0.0)):
        seen = {ThermalBrick(3, n, device=-1, perturb=perturb).plan.stat("plan_hash") for _ in range(3)}
        assert len(seen) == 1
