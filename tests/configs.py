"""Input decks used by the tests, written as the nested dicts of the reference's YAML
(ANONYMOUS root dropped).  The first two are the reference's own regression inputs verbatim."""
import copy

# regression/thermal/2D_verification/input.yaml  (BASELINE.json configs[0])
THERMAL_2D = {
    "Mesh": {"dimension": 2, "element type": "quad", "xmin": 0.0, "xmax": 1.0, "ymin": 0.0, "ymax": 1.0, "NX": 40, "NY": 40},
    "Functions": {"thermal source": "8*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)"},
    "Physics": {"modules": "thermal", "assemble boundary terms": False, "build face terms": True,
                "Dirichlet conditions": {"T": {"all boundaries": "0.0"}}, "Initial conditions": {"T": "0.0"}},
    "Discretization": {"order": {"T": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state", "workset size": 100, "nonlinear TOL": 1.0e-07, "max nonlinear iters": 2, "use strong DBCs": True},
    "Postprocess": {"compute errors": True, "True solutions": {
        "T": "sin(2*pi*x)*sin(2*pi*y)", "grad(T)[x]": "2*pi*cos(2*pi*x)*sin(2*pi*y)", "grad(T)[y]": "2*pi*sin(2*pi*x)*cos(2*pi*y)"}},
}
THERMAL_2D_GOLD = {"T": 0.00102776, "grad(T)": 0.201394}  # regression/thermal/2D_verification/mrhyde.gold

# regression/thermal/3D_verification/input.yaml
THERMAL_3D = {
    "Mesh": {"dimension": 3, "element type": "hex", "xmin": 0.0, "xmax": 1.0, "ymin": 0.0, "ymax": 1.0, "zmin": 0.0, "zmax": 1.0,
             "NX": 10, "NY": 10, "NZ": 10},
    "Physics": {"modules": "thermal", "Dirichlet conditions": {"T": {"all boundaries": "0.0"}}, "Initial conditions": {"T": "0.0"}},
    "Discretization": {"order": {"T": 1}, "quadrature": 2},
    "Functions": {"thermal source": "12*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)"},
    "Solver": {"solver": "steady-state"},
    "Postprocess": {"compute errors": True, "True solutions": {"T": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)"}},
}
THERMAL_3D_GOLD = {"T": 0.0116656}  # regression/thermal/3D_verification/mrhyde.gold


def variant(base, **updates):
    """Deep copy with 'A/B/key' = value overrides."""
    cfg = copy.deepcopy(base)
    for path, val in updates.items():
        node = cfg
        parts = path.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        if val is None:
            node.pop(parts[-1], None)
        else:
            node[parts[-1]] = val
    return cfg
