"""Opcode histogram of one kernel's SASS (cuobjdump -sass of the built library), whole function: tools/sass_hist.py <name substring> [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "semiseparablematrices.jl_b200/lib/libssmb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(2).split('.')[0]] += 1
for name, h in hist.items():
    if sys.argv[1] in name:
        tot = sum(h.values())
        fp = h['DFMA'] + h['DMUL'] + h['DADD']
        print(name, 'total', tot, 'fp64', fp)
        print('   ', ', '.join(f'{k} {v}' for k, v in h.most_common(16)))
