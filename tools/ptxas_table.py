#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (demangled): registers, stack, spills, smem."""
import re, subprocess, sys
txt = sys.stdin.read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for '(\S+)'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
    rows.append(m.groups())
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines() if rows else []
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for r, nm in zip(rows, names):
    nm = re.sub(r"\(.*", "", nm).replace("ssm::", "").replace("void ", "")
    if flt and not re.search(flt, nm): continue
    print(f"{int(r[5]):4d} regs  stack {int(r[2]):4d}  spill {int(r[3]):3d}/{int(r[4]):3d}  {nm}")
