#!/usr/bin/env python
"""Generates the committed golden fixtures from the reference's OWN regression tests.

Run in the development container (where /root/reference exists):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the JSON files written here.

  regression_gold.json   input deck (the YAML under ANONYMOUS) + the L2-error lines of mrhyde.gold for the
                         thermal, linear elasticity and Navier-Stokes regression cases on the hot path
  functions_valid.json   the `Functions:` block of regression/functions/Valid and the decomposition
                         forest its mrhyde.gold prints (tree -> branch expressions, in order)
"""
import json
import os
import re

import yaml

REF = "/root/reference/regression"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = ["thermal/2D_verification", "thermal/2D_verification_mpi", "thermal/3D_verification", "thermal/2D_verification_transient",
         "thermal/2D_mixed_bcs", "le/3D_manufactured", "le/2D_manufactured", "navierstokes/channel", "maxwell/PlaneWave", "maxwell/NonzeroIC",
         "thermoelastic/2D_transient"]   # a two-module block: "modules: thermal, linearelasticity"
# (thermal/2D_verification_nonzeroDBC is not usable as a pin: its Dirichlet data come from the solver's boundary
#  L2 projection, solverManager_util.hpp:24-48, which is outside the path)


def load_deck(path):
    """The ANONYMOUS list of an input deck with its `<Section> input file:` includes merged in (the reference's
    UserInterface does the same, src/interfaces/user/userInterface.cpp)."""
    with open(path) as f:
        deck = yaml.safe_load(f)["ANONYMOUS"]
    for key in [k for k in deck if k.endswith(" input file")]:
        with open(os.path.join(os.path.dirname(path), deck.pop(key))) as f:
            deck.update(yaml.safe_load(f)["ANONYMOUS"])
    return deck


def errors(path):
    out = []
    pat = re.compile(r"\*+ (L2(?:-face)? norm) of the error for (.+?) = (\S+)\s+\(time = (\S+)\)")
    for line in open(path):
        m = pat.search(line)
        if m:
            out.append({"norm": m.group(1), "field": m.group(2), "value": float(m.group(3)), "time": float(m.group(4))})
    return out


def forests(path):
    res, forest, tree = {}, None, None
    for line in open(path):
        s = line.rstrip("\n")
        if s.startswith("Forest Name:"):
            forest = s[len("Forest Name:"):]
            res[forest] = {}
        elif forest is not None and s.startswith("    Tree: "):
            tree = s[len("    Tree: "):]
            res[forest][tree] = []
        elif forest is not None and tree is not None and s.startswith("        "):
            res[forest][tree].append(s[8:])
        elif s.startswith("====") and forest is not None and res[forest]:
            forest = None
    return res


def main():
    gold = {}
    for c in CASES:
        d = os.path.join(REF, c)
        gold[c] = {"deck": load_deck(os.path.join(d, "input.yaml")), "errors": errors(os.path.join(d, "mrhyde.gold")),
                   "source": "regression/%s/{input.yaml,mrhyde.gold}" % c}
    with open(os.path.join(HERE, "regression_gold.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    d = os.path.join(REF, "functions/Valid")
    fv = {"functions": load_deck(os.path.join(d, "input.yaml"))["Functions"], "forests": forests(os.path.join(d, "mrhyde.gold")),
          "source": "regression/functions/Valid/{input.yaml,mrhyde.gold}"}
    with open(os.path.join(HERE, "functions_valid.json"), "w") as f:
        json.dump(fv, f, indent=1, sort_keys=True)
    print("wrote", len(gold), "regression cases and", {k: len(v) for k, v in fv["forests"].items()}, "trees")


if __name__ == "__main__":
    main()
