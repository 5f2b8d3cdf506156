#!/usr/bin/env python
"""Summarises an `nvcc -Xptxas -v` log: registers, stack and spill bytes per kernel.  usage: ptxas_report.py LOG"""
import re, subprocess, sys
log = open(sys.argv[1]).read()
ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log)
for name, stack, ss, sl, regs in ents:
    dn = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dn = re.sub(r'mrhyde_b200::', '', dn)[:120]
    print("%3s regs  stack %5s  spill st/ld %5s/%5s  %s" % (regs, stack, ss, sl, dn))
