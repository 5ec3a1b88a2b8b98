#!/bin/bash
# chunk count of the chunked single-system path at config 4 (stage 1 runs in waves of 148 SMs x 48 chunks = 7,104)
for P in 7104 14208 21312 28416 35520 42624 56832 71040; do
  python tools/bench_configs.py c4 --steps 10 --chunks $P 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks', d['chunks'], 'ms', round(d['ms'],4), 'relres', d['relres'])"
done
