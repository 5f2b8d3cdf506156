"""Test helpers: stand in for the reference's host objects by handing the C ABI the arrays the
oracle built (mesh, LIDs, graph, reference tables, boundary groups) -- the same data a MrHyDE host
would pass -- and compare results with the tolerances north_star states."""
import numpy as np

from mrhyde_b200.capi import AssemblyPlan, TimeSpec

SIDE_NAMES_2D = ["left", "right", "bottom", "top"]
SIDE_NAMES_3D = SIDE_NAMES_2D + ["back", "front"]
BC_NAMES = {0: "none", 1: "Dirichlet", 2: "weak Dirichlet", 3: "Neumann"}
DEFAULT_OPTIONS = {}   # plan options every plan_from_oracle call applies first (the gpu tests switch "jit" through it)


def plan_from_oracle(op, cfg, device=0, options=None, indexed=False):
    """AssemblyPlan for the block the OracleProblem `op` describes (cfg = its input deck)."""
    bases = []
    rb_vdim = []
    for b in range(op.nbases):
        rb = op.ref_basis(b)
        rb_vdim.append(rb["vdim"])
        bases.append(dict(type=op.basis_type(b), order=op.basis_order(b), card=rb["card"], val=rb["val"], grad=rb["grad"], curl=rb["curl"], div=rb["div"]))
    names = op.var_names()
    plan = AssemblyPlan(op.modules(), op.dim, names, op.usebasis, bases, op.ndof_elem, op.offsets, op.qpts, op.qwts, device=device)
    for k, v in cfg.get("Functions", {}).items():
        plan.set_function(k, v)
    phys = cfg.get("Physics", {})
    solver = cfg.get("Solver", {})
    for k, v in phys.get("Initial conditions", {}).items():   # "initial <var>" / "initial <var>[x]" (physicsInterface_functions.hpp:154-226)
        plan.set_function("initial " + k, str(v))
    for key in ("form_param", "include advection", "useSUPG", "usePSPG", "assemble boundary terms", "assemble volume terms", "penalty",
                "incplanestress", "ns3d_uz_rows", "use leap frog"):
        if key in phys:
            plan.set_option(key, phys[key])
    if "use strong DBCs" in solver:
        plan.set_option("use strong DBCs", solver["use strong DBCs"])
    if "lump mass" in solver:
        plan.set_option("lump mass", solver["lump mass"])
    if "fix zero rows" in solver:
        plan.set_option("fix zero rows", solver["fix zero rows"])
    for k, v in list(DEFAULT_OPTIONS.items()) + list((options or {}).items()):
        plan.set_option(k, v)
    if indexed:
        plan.set_mesh_indexed(op.nodes, op.conn, op.lids)
    else:
        plan.set_mesh(op.elem_nodes, op.lids)
    plan.set_graph(op.rowptr, op.colind, op.is_fixed)
    sides = SIDE_NAMES_2D if op.dim == 2 else SIDE_NAMES_3D
    plan.set_sidesets(sides)
    exprs = op.bc_exprs()
    for v, vn in enumerate(names):
        for s, sn in enumerate(sides):
            code = int(op.bc_codes[v, s])
            if code:
                plan.set_bc(vn, sn, BC_NAMES[code], exprs[v][s])
    for g in range(op.num_bgroups):
        bg = op.bgroup(g)
        pts, wts, tu, tv = op.side_rule(bg["local_side"])
        side_bases = []
        for b in range(op.nbases):
            card = bases[b]["card"]
            val, grad = op.ref_basis_side(bg["local_side"], b, card, has_grad=bases[b]["grad"] is not None, vdim=rb_vdim[b])
            side_bases.append(dict(type=bases[b]["type"], order=bases[b]["order"], card=card, val=val, grad=grad))
        plan.add_boundary_group(bg["sideset"], bg["local_side"], bg["elem_ids"], pts, wts, tu, tv, side_bases)
    plan.finalize()
    return plan


def rel_err_rows(jac, jac_ref, rowptr):
    """max over entries of |J - Jref| / (max-norm of that row of Jref): the north_star metric
    (1e-12 relative, measured against the row max for analytically-zero couplings)."""
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    rowmax = np.zeros(len(rowptr) - 1)
    np.maximum.at(rowmax, rows, np.abs(jac_ref))
    rowmax[rowmax == 0.0] = 1.0
    return float(np.max(np.abs(jac - jac_ref) / rowmax[rows])) if len(jac) else 0.0


def rel_err_vec(res, res_ref):
    scale = float(np.max(np.abs(res_ref)))
    return float(np.max(np.abs(res - res_ref))) / (scale if scale > 0 else 1.0)


def manufactured_state(op, seed=20261017):
    """u = smooth field + 1e-3 U(-1,1) noise; fixed dofs keep the noise too -- the kernels
    must treat them like any other column."""
    rng = np.random.default_rng(seed)
    return 0.3 * np.sin(1.0 + 0.7 * np.arange(op.num_dofs) / max(1, op.num_dofs)) + 1e-3 * rng.uniform(-1.0, 1.0, op.num_dofs)


# ---- a minimal stand-in for the reference's solver loop (solverManager_solvers.hpp:13-31, 40-282, 290-640) ------
# so that assembled systems can be checked against the L2 errors the reference's mrhyde.gold files print.
def dirichlet_values(op, cfg, time=0.0):
    """Strong-Dirichlet data at the fixed dofs (HGRAD Q1: dof == node), evaluated like setDirichlet does."""
    import math
    vals = np.zeros(op.num_dofs)
    dc = cfg.get("Physics", {}).get("Dirichlet conditions", {})
    sides = SIDE_NAMES_2D if op.dim == 2 else SIDE_NAMES_3D
    x = op.nodes
    lo, hi = x.min(axis=0), x.max(axis=0)
    on = {"left": np.isclose(x[:, 0], lo[0]), "right": np.isclose(x[:, 0], hi[0]), "bottom": np.isclose(x[:, 1], lo[1]), "top": np.isclose(x[:, 1], hi[1])}
    if op.dim == 3:
        on["back"] = np.isclose(x[:, 2], lo[2])
        on["front"] = np.isclose(x[:, 2], hi[2])
    env = {"sin": np.sin, "cos": np.cos, "exp": np.exp, "pi": math.pi, "x": x[:, 0], "y": x[:, 1], "z": x[:, 2] if op.dim == 3 else 0.0, "t": time}
    for var, spec in dc.items():
        if not isinstance(spec, dict):
            continue
        for side, expr in spec.items():
            names = sides if side == "all boundaries" else [side]
            v = eval(str(expr).replace("^", "**"), {"__builtins__": {}}, env) * np.ones(op.num_dofs)
            for s in names:
                vals[on[s]] = v[on[s]]
    return vals


def newton_steady(op, cfg, assemble, iters=2):
    """assemble(u) -> (res = -F(u), J values); u <- u + J^-1 res, strong-Dirichlet rows carry res = 0, J(d,d) = 1."""
    import scipy.sparse.linalg as spla
    u = np.zeros(op.num_dofs)
    fixed = op.is_fixed.astype(bool)
    u[fixed] = dirichlet_values(op, cfg)[fixed]
    for _ in range(iters):
        res, jac = assemble(u)
        u = u + spla.spsolve(op.csr(jac).tocsc(), res)
    return u


def gold_errors(case):
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regression_gold.json")))
    return g[case]["deck"], g[case]["errors"]


def deck_to_cfg(deck):
    """mrhyde input deck (golden fixture) -> the flat-block form the oracle reads (block sublists dropped)."""
    import copy
    cfg = copy.deepcopy(deck)
    for sec in ("Physics", "Discretization", "Postprocess"):
        blk = cfg.get(sec, {})
        keys = [k for k in blk if k.startswith("eblock")]
        for k in keys:
            blk.update(blk.pop(k))
    return cfg
