#!/bin/bash
# where the grouped-ring kernels (variant 4) should hand over to the big-batch kernels (K1 direct loads, K2/K3 per-step ring), config-2 shape
run() { echo -n "$*: "; python tools/bench_configs.py c2 --steps 20 "$@" 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('qr', round(d['qr_ms'],4), 'ldiv', round(d['ldiv_ms'],4), 'fused', round(d['fused_solve_ms'],4), 'frac', round(d['qr_plus_ldiv_frac'],4))"; }
for nb in 24576 32768 40960 49152; do
  run --nb $nb --opt group_max_per_sm=1
  run --nb $nb --opt group_max_per_sm=20
done
