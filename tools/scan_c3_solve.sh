#!/bin/bash
# BPS solve (K7 + K8) at config 3 over the GroupRing knobs (tuning aid)
run() { echo -n "$*: "; python tools/bench_configs.py c3 --steps 10 "$@" 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('solve', round(d['solve_ms'],4), 'qr', round(d['qr_ms'],4))"; }
run
for gb in 8192 16384 65536; do run --opt group_bytes=$gb; done
for kb in 48 96 140; do run --opt group_smem_kb=$kb; done
run --variant 2
run --variant 1
