#!/bin/bash
# GroupRing knobs at small batches of the config-2 shape, and K1 alternatives at the headline size (tuning aid)
run() { echo -n "$*: "; python tools/bench_configs.py c2 --steps 20 "$@" 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('qr', round(d['qr_ms'],4), 'ldiv', round(d['ldiv_ms'],4), 'fused', round(d['fused_solve_ms'],4), 'frac', round(d['qr_plus_ldiv_frac'],4))"; }
run --nb 65536 --opt carveout_all=1
run --nb 65536 --variant 2
for nb in 16384 32768; do
  run --nb $nb
  for kb in 32 64 96; do run --nb $nb --opt group_smem_kb=$kb; done
  for gb in 8192 16384 65536; do run --nb $nb --opt group_bytes=$gb; done
done
run --nb 8192 --n 8192
run --nb 8192 --n 8192 --opt group_smem_kb=96
run --nb 8192 --n 8192 --opt group_bytes=65536
