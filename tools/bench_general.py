#!/usr/bin/env python
"""Times the general path (element kernel + pull) on synthetic bricks: python tools/bench_general.py PHYSICS N [key=value ...]
PHYSICS = le | ns | thermal | leq2 | thq2 | maxwell (inputs built by mrhyde_b200/problems.py); prints one JSON line (device time per assemble call via CUDA events, elements/s, the
algorithmic-bytes rate of SURVEY 8(d) and its fraction of the measured HBM peak)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrhyde_b200.problems import ElasticityQ2Brick, MaxwellBrick, SystemBrick, ThermalBrick

phys, n = sys.argv[1], int(sys.argv[2])
opts = dict(a.split("=", 1) for a in sys.argv[3:])
steps = int(opts.pop("steps", 10))
opts.setdefault("accumulate", "false")
if phys == "thermal":
    opts.setdefault("kernel", "general")
    prob = ThermalBrick(3, [n, n, n], device=0, options=opts)
elif phys == "leq2":
    prob = ElasticityQ2Brick(n, device=0, options=opts)
elif phys in ("le", "ns"):
    prob = SystemBrick({"le": "linearelasticity", "ns": "navier stokes"}[phys], 3, [n, n, n], device=0, options=opts)
elif phys == "thq2":
    opts.setdefault("kernel", "general")
    prob = ElasticityQ2Brick(n, device=0, options=opts, physics="thermal")
elif phys == "maxwell":
    prob = MaxwellBrick(n, device=0, options=opts)
else:
    raise SystemExit("unknown PHYSICS " + phys)
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
