#!/bin/bash
# ring depth of K2 (qt_stages) and K3 (bs_stages) at config 2 (tuning aid)
for q in 0 2 3 4 6 8; do for b in 0 2 3 4; do
  echo -n "qt_stages $q bs_stages $b: "
  python tools/bench_configs.py c2 --steps 20 --opt qt_stages=$q --opt bs_stages=$b 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ldiv', round(d['ldiv_ms'],4))"
done; done
