"""Aggregate the per-instruction stall samples of an `ncu --page source --csv` dump by opcode and by stall reason."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); byop = collections.Counter(); cnt = collections.Counter()
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ts = 0
for r in data:
    if len(r) < len(hdr):
        continue
    toks = r[ix['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    s = int(r[ix['# Samples']]); ts += s
    byop[op] += s; cnt[op] += int(r[ix['Instructions Executed']])
    for h in stalls:
        tot[h] += int(r[ix[h]])
print('total samples', ts)
for h, v in tot.most_common(12):
    print(f'{h:24s} {v:8d} {v / ts:.3f}')
ti = sum(cnt.values())
print()
for op, v in byop.most_common(25):
    print(f'{op:10s} samples {v:8d} {v / ts:.3f}   insts {cnt[op]:12d} {cnt[op] / ti:.3f}')
