#!/usr/bin/env python
"""Registers / spills per kernel from an `nvcc -Xptxas -v` log.  usage: python tools/ptxas_regs.py build.log [name filter]"""
import re, sys
t = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", t):
    if flt in m.group(1):
        print(m.group(1)[12:64], "regs", m.group(5), "spill", m.group(3), m.group(4))
