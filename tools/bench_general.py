#!/usr/bin/env python
"""Times the general path (element kernel + pull) on synthetic bricks: python tools/bench_general.py PHYSICS N [key=value ...]
PHYSICS = le | ns | thermal | leq2 (numpy-built bricks) | maxwell | thq2 (set up through the oracle's mesh / DOF tables, so
keep N moderate); prints one JSON line (device time per assemble call via CUDA events, elements/s, the
algorithmic-bytes rate of SURVEY 8(d) and its fraction of the measured HBM peak)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrhyde_b200.problems import ElasticityQ2Brick, SystemBrick, ThermalBrick

phys, n = sys.argv[1], int(sys.argv[2])
opts = dict(a.split("=", 1) for a in sys.argv[3:])
steps = int(opts.pop("steps", 10))
opts.setdefault("accumulate", "false")
class _OracleBacked:
    """Problem arrays from the oracle's set-up code (test infrastructure): used only to BUILD inputs for configurations that
    have no numpy builder; the timed path is the CUDA library."""

    def __init__(self, cfg, options):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import helpers
        from oracle import pyoracle
        self.op = pyoracle.OracleProblem(cfg)
        self.plan = helpers.plan_from_oracle(self.op, cfg, device=0, options=options)
        self.n_elem, self.n_rows, self.nnz = self.op.num_elems, self.op.num_dofs, self.op.nnz
        self._state = helpers.manufactured_state(self.op)
        self._A = 4.0 * self.op.ndof_elem * self.n_elem + 24.0 * self.op.num_nodes + 16.0 * self.n_rows + 8.0 * self.nnz

    def state(self):
        return self._state

    def algorithmic_bytes(self):
        return self._A


if phys == "thermal":
    opts.setdefault("kernel", "general")
    prob = ThermalBrick(3, [n, n, n], device=0, options=opts)
elif phys == "leq2":
    prob = ElasticityQ2Brick(n, device=0, options=opts)
elif phys in ("le", "ns"):
    prob = SystemBrick({"le": "linearelasticity", "ns": "navier stokes"}[phys], 3, [n, n, n], device=0, options=opts)
else:
    mesh = {"dimension": 3, "NX": n, "NY": n, "NZ": n}
    if phys == "leq2":
        cfg = {"Mesh": mesh, "Physics": {"modules": "linearelasticity", "Dirichlet conditions": {v: {"all boundaries": "0.0"} for v in ("dx", "dy", "dz")}},
               "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)"},
               "Discretization": {"order": {"dx": 2, "dy": 2, "dz": 2}, "quadrature": 4}, "Solver": {"workset size": 100}}
    elif phys == "thq2":
        cfg = {"Mesh": mesh, "Physics": {"modules": "thermal", "Dirichlet conditions": {"T": {"all boundaries": "0.0"}}},
               "Functions": {"thermal source": "12*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)"},
               "Discretization": {"order": {"T": 2}, "quadrature": 4}, "Solver": {"workset size": 100}}
        opts.setdefault("kernel", "general")
    else:
        cfg = {"Mesh": mesh, "Physics": {"modules": "maxwell", "Dirichlet conditions": {"E": {"all boundaries": "0.0"}}},
               "Functions": {"current x": "sin(2*pi*z)"}, "Discretization": {"order": {"E": 1, "B": 1}, "quadrature": 2}, "Solver": {"workset size": 100}}
    prob = _OracleBacked(cfg, opts)
plan = prob.plan
dev = torch.device("cuda:0")
u = torch.from_numpy(prob.state()).to(dev)
res = torch.empty(prob.n_rows, dtype=torch.float64, device=dev)
jac = torch.empty(prob.nnz, dtype=torch.float64, device=dev)
for _ in range(3):
    plan.assemble_jacres(u, res, jac)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.current_stream().cuda_stream
e0.record()
for _ in range(steps):
    plan.assemble_jacres(u, res, jac, stream=st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
A = prob.algorithmic_bytes()
peak = 6454.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"physics": phys, "n": n, "elements": prob.n_elem, "rows": prob.n_rows, "nnz": prob.nnz, "ms_per_assemble": ms,
                  "elements_per_s": prob.n_elem / (ms * 1e-3), "algorithmic_GBps": A / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": A / (ms * 1e-3) / 1e9 / peak,
                  "batches": plan.stat("general_batches"), "launches": plan.stat("kernel_launches_per_assemble"),
                  "scratch_MB": plan.stat("general_scratch_bytes") / 1e6, "options": opts, "jac_checksum": float(jac.sum()), "res_checksum": float(res.abs().sum())}))
