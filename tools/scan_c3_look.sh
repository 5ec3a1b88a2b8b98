#!/bin/bash
# scan the look-ahead depth of the one-warp BPS QR sweep at config 3 (tuning aid)
for s in 1 2 4 6; do
  echo -n "look $s: "
  python tools/bench_configs.py c3 --stages $s --steps 10 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('qr_ms', round(d['qr_ms'],4), 'frac', round(d['qr_frac_of_measured_peak'],4), 'relres', d['relres_max'])"
done
