#!/bin/bash
# Step-kernel time of the three BASELINE.json single-GPU configs (run after every kernel change: a change that is
# neutral for case33 / 8 lanes per env can cost the wide groups of case141 / case322 a lot, and vice versa).
for c in case33 case141 case322; do
  python bench.py --no-cpu --e2e-steps 3 --scenario $c 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e6,2), 'M env-steps/s')"
done
