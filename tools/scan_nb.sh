#!/bin/bash
# qr / ldiv! / fused at several batch sizes of the config-2 shape (tuning aid for the stage / variant selection rules)
for nb in ${NBS:-16384 32768 45056 49152 65536 131072}; do
  echo -n "nb $nb: "
  python tools/bench_configs.py c2 --steps 20 --nb $nb $@ 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('qr', round(d['qr_ms'],4), 'ldiv', round(d['ldiv_ms'],4), 'fused', round(d['fused_solve_ms'],4), 'frac', round(d['qr_plus_ldiv_frac'],4))"
done
