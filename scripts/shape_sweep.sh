#!/bin/bash
# CTA-shape sweep of the step kernel: envs per CTA (MAPDN_EPB) x helper warps (MAPDN_HELPERS).
run() { # scenario epb helpers
  MAPDN_EPB=$2 MAPDN_HELPERS=$3 python bench.py --no-cpu --e2e-steps 3 --scenario $1 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 epb=$2 helpers=$3', round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e6,2), 'M env-steps/s')"
}
for cfg in "8 1" "28 1" "28 2" "28 4" "4 1" "16 2"; do run case33 $cfg; done
for cfg in "4 1" "7 1" "7 2" "4 2"; do run case141 $cfg; done
for cfg in "3 1" "3 2" "3 4"; do run case322 $cfg; done
