"""Test helpers: stand in for the reference's host objects by handing the C ABI the arrays the
oracle built (mesh, LIDs, graph, reference tables, boundary groups) -- the same data a MrHyDE host
would pass -- and compare results with the tolerances north_star states."""
import numpy as np

from mrhyde_b200.capi import AssemblyPlan, TimeSpec

SIDE_NAMES_2D = ["left", "right", "bottom", "top"]
SIDE_NAMES_3D = SIDE_NAMES_2D + ["back", "front"]
BC_NAMES = {0: "none", 1: "Dirichlet", 2: "weak Dirichlet", 3: "Neumann"}


def plan_from_oracle(op, cfg, device=0, options=None, indexed=False):
    """AssemblyPlan for the block the OracleProblem `op` describes (cfg = its input deck)."""
    bases = []
    for b in range(op.nbases):
        rb = op.ref_basis(b)
        bases.append(dict(type=op.basis_type(b), order=op.basis_order(b), card=rb["card"], val=rb["val"], grad=rb["grad"], curl=rb["curl"], div=rb["div"]))
    names = op.var_names()
    plan = AssemblyPlan(op.modules(), op.dim, names, op.usebasis, bases, op.ndof_elem, op.offsets, op.qpts, op.qwts, device=device)
    for k, v in cfg.get("Functions", {}).items():
        plan.set_function(k, v)
    phys = cfg.get("Physics", {})
    solver = cfg.get("Solver", {})
    for key in ("form_param", "include advection", "useSUPG", "usePSPG", "assemble boundary terms", "assemble volume terms"):
        if key in phys:
            plan.set_option(key, phys[key])
    if "use strong DBCs" in solver:
        plan.set_option("use strong DBCs", solver["use strong DBCs"])
    for k, v in (options or {}).items():
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
            val, grad = op.ref_basis_side(bg["local_side"], b, card, has_grad=bases[b]["grad"] is not None)
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
