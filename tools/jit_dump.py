#!/usr/bin/env python
"""Builds a host-only thermal brick plan (no GPU needed), writes the plan-specialised translation unit and its sm_100a cubin, and prints
registers / stack / shared memory and SASS instruction counts.  usage: jit_dump.py N OUTDIR [key=value ...] [--transient] [--mode M]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrhyde_b200.problems import ThermalBrick

def main():
    os.environ["MRHYDE_B200_DEBUG_OPTIONS"] = "1"   # the tool forces build variants through the kernel-debugging keys
    n, out = int(sys.argv[1]), sys.argv[2]
    opts, transient, mode = {}, 0, None
    args = sys.argv[3:]
    i = 0
    while i < len(args):
        a = args[i]
        if a == "--transient": transient = 1
        elif a == "--mode": mode = int(args[i + 1]); i += 1
        else:
            k, v = a.split("=", 1); opts[k] = v
        i += 1
    os.makedirs(out, exist_ok=True)
    prob = ThermalBrick(3, [n, n, n], device=-1, options=opts)
    plan = prob.plan
    if transient: plan.set_option("debug transient", 1)
    if mode is not None: plan.set_option("debug mode", mode)
    src, cub = os.path.join(out, "k.cu"), os.path.join(out, "k.cubin")
    log = plan.debug_jit(source_path=src, cubin_path=cub)
    if "error" in log.lower(): print(log[-3000:])
    ru = subprocess.run(["cuobjdump", "--dump-resource-usage", cub], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ru)
    sass = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
    open(os.path.join(out, "k.sass"), "w").write(sass)
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass, re.M)
    from collections import Counter
    c = Counter(x.split(".")[0] for x in ins)
    print("options", opts, "n_chains", plan.stat("n_chains"), "smem", plan.stat("smem_bytes"), "REG/STACK/SHARED", m.groups() if m else ru[-300:])
    print("SASS", len(ins), "instr;", ", ".join("%s %d" % kv for kv in c.most_common(24)))
    for key in ("UBLKCP", "UTMA", "LDL", "STL", "SYNCS", "FENCE"):
        print("  ", key, sum(v for k, v in c.items() if k.startswith(key)))

if __name__ == "__main__":
    main()
