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


# ---- decks for the other modules on the path (synthetic; coefficient choices exercise every term) ------------------
LE_3D = {
    "Mesh": {"dimension": 3, "NX": 4, "NY": 3, "NZ": 3, "perturb": 0.03},
    "Physics": {"modules": "linearelasticity",
                "Dirichlet conditions": {"dx": {"all boundaries": "0.0"}, "dy": {"all boundaries": "0.0"}, "dz": {"all boundaries": "0.0"}}},
    "Functions": {"lambda": "1.0+0.3*x", "mu": "1.0", "A": "1.0", "source dx": "sin(A*pi*x)*y", "source dy": "lambda*z", "source dz": "1.0-x*y"},
    "Discretization": {"order": {"dx": 1, "dy": 1, "dz": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
LE_2D = {
    "Mesh": {"dimension": 2, "NX": 6, "NY": 5, "perturb": 0.03},
    "Physics": {"modules": "linearelasticity", "Dirichlet conditions": {"dx": {"all boundaries": "0.0"}, "dy": {"all boundaries": "0.0"}}},
    "Functions": {"lambda": "2.0", "mu": "0.7", "source dx": "x*y", "source dy": "1.0"},
    "Discretization": {"order": {"dx": 1, "dy": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
LE_3D_WEAK = {"Solver/use strong DBCs": False,
              "Physics/Dirichlet conditions": {"dx": {"left": "0.1*y", "top": "0.0"}, "dy": {"left": "0.0", "top": "x"}, "dz": {"left": "0.0", "top": "0.2"}},
              "Physics/Neumann conditions": {"dx": {"right": "1.0"}, "dy": {"right": "y"}, "dz": {"right": "0.0"}}}
LE_2D_WEAK = {"Solver/use strong DBCs": False,
              "Physics/Dirichlet conditions": {"dx": {"left": "0.1*y", "top": "0.0"}, "dy": {"left": "0.0", "top": "x"}},
              "Physics/Neumann conditions": {"dx": {"right": "1.0"}, "dy": {"right": "y"}}}
NS_2D = {
    "Mesh": {"dimension": 2, "NX": 6, "NY": 5, "perturb": 0.02},
    "Physics": {"modules": "navier stokes", "useSUPG": True, "usePSPG": True,
                "Dirichlet conditions": {"ux": {"left": "1.0", "top": "0.0", "bottom": "0.0"}, "uy": {"left": "0.0", "top": "0.0", "bottom": "0.0"},
                                         "pr": {"right": "0.0"}}},
    "Functions": {"source ux": "1.0", "viscosity": "0.5", "density": "1.3"},
    "Discretization": {"order": {"ux": 1, "pr": 1, "uy": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
NS_3D = {
    "Mesh": {"dimension": 3, "NX": 4, "NY": 3, "NZ": 3, "perturb": 0.02},
    "Physics": {"modules": "navier stokes", "useSUPG": True, "usePSPG": True,
                "Dirichlet conditions": {"ux": {"all boundaries": "0.0"}, "uy": {"all boundaries": "0.0"}, "uz": {"all boundaries": "0.0"}}},
    "Functions": {"source ux": "1.0", "source uz": "x"},
    "Discretization": {"order": {"ux": 1, "pr": 1, "uy": 1, "uz": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
NS_3D_NEUMANN = {"Physics/Dirichlet conditions": {"ux": {"left": "1.0"}, "uy": {"left": "0.0"}, "uz": {"left": "0.0"}},
                 "Physics/Neumann conditions": {"ux": {"right": "0.3"}, "uy": {"right": "y"}, "uz": {"top": "1.0"}}}
MAXWELL_3D = {
    "Mesh": {"dimension": 3, "NX": 3, "NY": 3, "NZ": 4, "perturb": 0.02},
    "Physics": {"modules": "maxwell", "Dirichlet conditions": {"E": {"all boundaries": "0.0"}}},
    "Functions": {"current x": "sin(t)*y", "permittivity": "1.5", "permeability": "0.8", "conductivity": "0.2", "refractive index": "1.1"},
    "Discretization": {"order": {"E": 1, "B": 1}, "quadrature": 2},
    "Solver": {"solver": "transient"},
}
MAXWELL_ABC = {"Physics/Dirichlet conditions": {"E": {"left": "0.0"}}, "Physics/Neumann conditions": {"B": {"right": "0.0", "top": "0.0", "front": "0.0"}}}
# two-module blocks (the reference's regression/thermoelastic/2D_transient couples "thermal, linearelasticity"; "navier stokes, thermal"
# switches thermal's have_nsvel advection on, thermal.cpp:117-149)
THERMOELASTIC_2D = {
    "Mesh": {"dimension": 2, "NX": 6, "NY": 5, "perturb": 0.03},
    "Physics": {"modules": "thermal, linearelasticity",
                "Dirichlet conditions": {"T": {"all boundaries": "0.0"}, "dx": {"all boundaries": "0.0"}, "dy": {"all boundaries": "0.0"}}},
    "Functions": {"thermal source": "2*pi*pi*sin(pi*x)*sin(pi*y)", "thermal diffusion": "1.0+0.5*x", "density": "1.2", "lambda": "2.0", "mu": "0.7",
                  "source dx": "x*y", "source dy": "1.0"},
    "Discretization": {"order": {"T": 1, "dx": 1, "dy": 1}, "quadrature": 2},
    "Solver": {"solver": "transient"},
}
THERMOELASTIC_2D_WEAK = {"Solver/use strong DBCs": False,
                         "Physics/Dirichlet conditions": {"T": {"left": "1.0+y", "top": "x*x"}, "dx": {"left": "0.1*y", "top": "0.0"}, "dy": {"left": "0.0", "top": "x"}},
                         "Physics/Neumann conditions": {"T": {"right": "2.0*y-0.3"}, "dx": {"right": "1.0"}, "dy": {"right": "y"}}}
THERMOELASTIC_3D = {
    "Mesh": {"dimension": 3, "NX": 4, "NY": 3, "NZ": 3, "perturb": 0.03},
    "Physics": {"modules": "thermal, linearelasticity",
                "Dirichlet conditions": {"T": {"all boundaries": "0.0"}, "dx": {"all boundaries": "0.0"}, "dy": {"all boundaries": "0.0"}, "dz": {"all boundaries": "0.0"}}},
    "Functions": {"thermal source": "sin(pi*x)*y+z", "lambda": "1.0+0.3*x", "mu": "1.0", "source dx": "sin(pi*x)*y", "source dy": "lambda*z", "source dz": "1.0-x*y"},
    "Discretization": {"order": {"T": 1, "dx": 1, "dy": 1, "dz": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
NS_THERMAL_2D = {
    "Mesh": {"dimension": 2, "NX": 6, "NY": 5, "perturb": 0.02},
    "Physics": {"modules": "navier stokes, thermal", "useSUPG": True, "usePSPG": True,
                "Dirichlet conditions": {"ux": {"left": "1.0", "top": "0.0", "bottom": "0.0"}, "uy": {"left": "0.0", "top": "0.0", "bottom": "0.0"},
                                         "pr": {"right": "0.0"}, "T": {"left": "1.0", "bottom": "0.0"}}},
    "Functions": {"source ux": "1.0", "viscosity": "0.5", "density": "1.3", "thermal source": "x+y", "thermal diffusion": "0.2", "specific heat": "1.1"},
    "Discretization": {"order": {"ux": 1, "pr": 1, "uy": 1, "T": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
NS_THERMAL_3D = {
    "Mesh": {"dimension": 3, "NX": 4, "NY": 3, "NZ": 3, "perturb": 0.02},
    "Physics": {"modules": "navier stokes, thermal", "useSUPG": True, "usePSPG": True,
                "Dirichlet conditions": {"ux": {"all boundaries": "0.0"}, "uy": {"all boundaries": "0.0"}, "uz": {"all boundaries": "0.0"}, "T": {"left": "1.0", "right": "0.0"}}},
    "Functions": {"source ux": "1.0", "source uz": "x", "thermal source": "x+y*z", "thermal diffusion": "0.2", "specific heat": "1.1", "density": "1.3"},
    "Discretization": {"order": {"ux": 1, "pr": 1, "uy": 1, "uz": 1, "T": 1}, "quadrature": 2},
    "Solver": {"solver": "steady-state"},
}
NS_THERMAL_2D_WEAK = {"Solver/use strong DBCs": False,
                      "Physics/Dirichlet conditions": {"ux": {"left": "1.0"}, "uy": {"left": "0.0"}, "T": {"left": "1.0+y", "top": "x*x"}},
                      "Physics/Neumann conditions": {"ux": {"right": "0.3"}, "uy": {"right": "y"}, "T": {"right": "2.0*y-0.3"}}}
THERMAL_WEAK = {"Solver/use strong DBCs": False, "Physics/assemble boundary terms": True, "Mesh/NX": 6, "Mesh/NY": 5, "Mesh/perturb": 0.01,
                "Physics/Dirichlet conditions/T": {"left": "1.0+y", "top": "x*x"}, "Physics/Neumann conditions/T": {"right": "2.0*y-0.3"}}
BWE = ([[1.0]], [1.0], [1.0], 0)
DIRK12 = ([[0.5]], [1.0], [0.5], 0)


def general_cases():
    """(name, deck, plan options, transient tableau or None, zero state?) for the general path's parity tests."""
    t3 = variant(THERMAL_3D, **{"Mesh/NX": 5, "Mesh/NY": 4, "Mesh/NZ": 3, "Mesh/perturb": 0.03})
    return [
        ("thermal3d", t3, {}, None, False),
        ("thermal3d-dirk", t3, {}, DIRK12, False),
        ("thermal2d-weak-neumann", variant(THERMAL_2D, **THERMAL_WEAK), {}, None, False),
        ("thermal3d-weak-neumann", variant(THERMAL_3D, **dict(THERMAL_WEAK, **{"Mesh/NZ": 4})), {}, None, False),
        ("thermal3d-advection", variant(t3, **{"Physics/include advection": True, "Functions/advection x": "1.0+y", "Functions/advection y": "x",
                                               "Functions/advection z": "0.5"}), {}, None, False),
        ("thermal3d-q2", variant(t3, **{"Discretization/order/T": 2, "Discretization/quadrature": 4}), {}, None, False),
        ("le3d", LE_3D, {}, None, False),
        ("le2d", LE_2D, {}, None, False),
        ("le2d-planestress", variant(LE_2D, **{"Physics/incplanestress": True}), {}, None, False),
        ("le3d-q2", variant(LE_3D, **{"Discretization/order": {"dx": 2, "dy": 2, "dz": 2}, "Discretization/quadrature": 4, "Mesh/NX": 2, "Mesh/NY": 2, "Mesh/NZ": 2}), {}, None, False),
        ("le3d-weak-neumann", variant(LE_3D, **LE_3D_WEAK), {}, None, False),
        ("le2d-weak-neumann", variant(LE_2D, **LE_2D_WEAK), {}, None, False),
        ("ns2d-supg-pspg", NS_2D, {}, None, False),
        ("ns2d-galerkin", variant(NS_2D, **{"Physics/useSUPG": False, "Physics/usePSPG": False}), {}, None, False),
        ("ns2d-bwe", NS_2D, {}, BWE, False),
        ("ns3d-reference-uz-rows", NS_3D, {}, None, False),
        ("ns3d-corrected-uz-rows", variant(NS_3D, **{"Physics/ns3d_uz_rows": "corrected"}), {}, None, False),
        ("ns3d-stagnation-tau-branch", NS_3D, {}, None, True),
        ("ns3d-neumann", variant(NS_3D, **NS_3D_NEUMANN), {}, None, False),
        ("maxwell-steady", MAXWELL_3D, {}, None, False),
        ("maxwell-dirk", MAXWELL_3D, {}, DIRK12, False),
        ("maxwell-abc-bwe", variant(MAXWELL_3D, **MAXWELL_ABC), {}, BWE, False),
        ("le3d-batched", variant(LE_3D, **{"Mesh/NX": 5, "Mesh/NY": 4, "Mesh/NZ": 4}), {"batch elems": 16}, None, False),
        # coefficients that read the solution: the reference carries its AD type through FunctionManager::evaluate
        # (functionManager_evaluate.hpp:59-229); here the bytecode is differentiated in the Jacobian stages
        ("thermal3d-state-diffusion", variant(t3, **{"Functions/thermal diffusion": "1.0+T*T", "Functions/thermal source": "sin(x)*T+y"}), {}, None, False),
        ("thermal3d-state-dirk", variant(t3, **{"Functions/thermal diffusion": "exp(0.3*T)+0.1*grad(T)[x]*grad(T)[x]", "Functions/specific heat": "1.0+0.2*T",
                                                "Functions/density": "2.0"}), {}, DIRK12, False),
        ("thermal3d-q2-state", variant(t3, **{"Discretization/order/T": 2, "Discretization/quadrature": 4, "Functions/thermal diffusion": "max(0.5,1.0+T)"}), {}, None, False),
        ("le3d-state-mu", variant(LE_3D, **{"Functions/mu": "1.0+0.2*dx*dx+0.1*grad(dy)[z]"}), {}, None, False),
        ("ns2d-state-viscosity", variant(NS_2D, **{"Functions/viscosity": "0.5+0.1*ux*ux+0.05*sqrt(1.0+uy*uy)"}), {}, None, False),
        ("maxwell-state-sigma", variant(MAXWELL_3D, **{"Functions/conductivity": "0.2+E[x]*E[x]+0.1*B[z]"}), {}, DIRK12, False),
        # element reductions inside an expression (functionManager_evaluate.hpp:413-460, literal semantics incl. emean's double count of point 0)
        ("thermal3d-emax", variant(t3, **{"Functions/thermal diffusion": "1.0+emax(x*y)", "Functions/thermal source": "emean(x)+emin(y*z)*x"}), {}, None, False),
        ("le3d-emean", variant(LE_3D, **{"Functions/lambda": "1.0+emean(x*z)", "Functions/source dy": "emax(y*z)+x"}), {}, None, False),
        ("thermal2d-weak-emin", variant(THERMAL_2D, **dict(THERMAL_WEAK, **{"Functions/thermal diffusion": "1.0+emin(x+y)"})), {}, None, False),
        # Solver: lump mass -- the fused scatter adds every entry of a row to its diagonal (assemblyManager_scatter.hpp:263-268)
        ("thermal3d-lump-dirk", variant(t3, **{"Solver/lump mass": True, "Functions/density": "2.0"}), {}, DIRK12, False),
        ("thermal2d-lump-dirk", variant(THERMAL_2D, **dict(THERMAL_WEAK, **{"Solver/lump mass": True})), {}, DIRK12, False),
        # Solver: fix zero rows (assemblyManager_jacres.hpp:609-626): rows whose entries sum to < 1e-14 in absolute value get a unit diagonal
        ("thermal3d-fix-zero-rows", variant(t3, **{"Solver/fix zero rows": True, "Functions/thermal diffusion": "0.0"}), {}, None, False),   # every row is empty
        ("maxwell-fix-zero-rows", variant(MAXWELL_3D, **{"Solver/fix zero rows": True}), {}, DIRK12, True),
        # two-module blocks
        ("thermoelastic2d", THERMOELASTIC_2D, {}, None, False),
        ("thermoelastic2d-bwe", THERMOELASTIC_2D, {}, BWE, False),
        ("thermoelastic2d-weak", variant(THERMOELASTIC_2D, **THERMOELASTIC_2D_WEAK), {}, None, False),
        ("thermoelastic3d-dirk", THERMOELASTIC_3D, {}, DIRK12, False),
        ("ns-thermal2d", NS_THERMAL_2D, {}, None, False),
        ("ns-thermal2d-bwe", NS_THERMAL_2D, {}, BWE, False),
        ("ns-thermal2d-weak", variant(NS_THERMAL_2D, **NS_THERMAL_2D_WEAK), {}, None, False),
        ("ns-thermal3d", NS_THERMAL_3D, {}, None, False),
        ("ns-thermal3d-dirk", NS_THERMAL_3D, {}, DIRK12, False),
        ("ns-thermal3d-neumann", variant(NS_THERMAL_3D, **dict(NS_3D_NEUMANN, **{"Physics/Neumann conditions": {"ux": {"right": "0.3"}, "uy": {"right": "y"}, "uz": {"top": "1.0"}, "T": {"top": "0.5*x"}}})), {}, None, False),
        ("ns-thermal2d-state", variant(NS_THERMAL_2D, **{"Functions/thermal diffusion": "0.2+0.1*T*T", "Functions/viscosity": "0.5+0.1*T"}), {}, None, False),
        ("thermal2d-weak-state", variant(THERMAL_2D, **dict(THERMAL_WEAK, **{"Functions/thermal diffusion": "1.0+0.5*T*T"})), {}, None, False),
    ]
